#!/bin/bash
# ncu --set full of the fused pair kernels and the single-conv kernels they replace (4K shapes), raw metrics to CSV.
mkdir -p gpurun_out
for cfg in "enc 2160 3840 16 32 32 1" "dec 2160 3840 64 32 16 0"; do
  set -- $cfg; tag=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -c 3 -f -o gpurun_out/ncu_pair_$tag tools/bin/probe_pair "$@" > gpurun_out/ncu_pair_$tag.log 2>&1
  tail -2 gpurun_out/ncu_pair_$tag.log
done
ls -la gpurun_out/*.ncu-rep
