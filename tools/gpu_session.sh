#!/bin/bash
# One GPU-box session = a list of named steps (replaces the per-session scripts of round 1). Every step writes its
# log under gpurun_out/<tag>/ and prints a short tail. usage: tools/gpu_session.sh <tag> step [step ...]
#   probe_mma   tcgen05.mma issue-cost microbenchmark           witness   reference CUDA device vs this backend vs oracle
#   refbench    the reference's own CUDA device (oidnBenchmark)  ops       op-level module under the unmodified reference core
#   pytest      the -m gpu suite   smoke   __graft_entry__.smoke   bench   python bench.py   configs   tools/bench_configs.py
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
clk() { nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader; }
for step in "$@"; do
  echo "=== $step"; 
  case $step in
    probe_mma) timeout 200 bash tools/run_probe_mma.sh > /dev/null 2>&1; cp gpurun_out/probe_mma_issue.log $OUT/; cat $OUT/probe_mma_issue.log;;
    witness)   timeout 200 python tools/reference_cuda_witness.py > $OUT/reference_cuda_witness.log 2>&1; tail -5 $OUT/reference_cuda_witness.log;;
    refbench)  ( export LD_LIBRARY_PATH=$PWD/baseline/_ref/lib; B=baseline/_ref/bin/oidnBenchmark
                 ( while true; do clk; sleep 0.2; done ) > $OUT/refbench_clocks.log & CP=$!
                 timeout 60 $B -d cuda -r "RT\.hdr_alb_nrm\.3840x2160" -q high
                 timeout 60 $B -d cuda -r "RT\.hdr_calb_cnrm\.3840x2160" -q high
                 kill $CP ) > $OUT/refbench.log 2>&1; cat $OUT/refbench.log; sort $OUT/refbench_clocks.log | uniq -c | sort -rn | head -3;;
    ops)       bash tools/integration_run.sh 90 ops > /dev/null 2>&1; cp gpurun_out/integration_run_ops.log $OUT/; tail -30 $OUT/integration_run_ops.log;;
    pytest)    timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log;;
    smoke)     timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log;;
    bench)     timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err; cat $OUT/bench.json;;
    configs)   timeout 900 python tools/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; tail -3 $OUT/configs.err; cat $OUT/configs.jsonl;;
    *)         echo "unknown step $step";;
  esac
done
