#!/bin/bash
# Ring-size sweep of the fused conv pairs on the 4K shapes (env overrides of the pair planner).
P=tools/bin/probe_pair
mkdir -p gpurun_out
one() { echo "--- [$1] $2"; env $1 timeout 60 $P $2 2>&1 | grep -E "^cfg|RESULT|TIME|not supp" | cut -c1-150; }
{
for shape in "2160 3840 16 32 32 1 10" "2160 3840 64 32 16 0 10"; do
  one "X=0" "$shape"
  one "OIDN_B200_PAIR_STREAMS=1" "$shape"
  one "OIDN_B200_PAIR_STREAMS=2 OIDN_B200_PAIR_RA=4 OIDN_B200_PAIR_RB=4" "$shape"
  one "OIDN_B200_PAIR_STREAMS=2 OIDN_B200_PAIR_RA=3 OIDN_B200_PAIR_RB=5" "$shape"
  one "OIDN_B200_PAIR_STREAMS=2 OIDN_B200_PAIR_RA=3 OIDN_B200_PAIR_RB=5 OIDN_B200_PAIR_NM=4" "$shape"
  one "OIDN_B200_PAIR_STREAMS=1 OIDN_B200_PAIR_RA=6 OIDN_B200_PAIR_RB=10 OIDN_B200_PAIR_NM=8" "$shape"
  one "OIDN_B200_PAIR_STREAMS=1 OIDN_B200_PAIR_RA=8 OIDN_B200_PAIR_RB=8 OIDN_B200_PAIR_NM=4" "$shape"
  one "OIDN_B200_PAIR_STREAMS=1 OIDN_B200_PAIR_RA=4 OIDN_B200_PAIR_RB=6 OIDN_B200_PAIR_NM=3 OIDN_B200_PAIR_NA=6" "$shape"
done
} > gpurun_out/probe_pair_sweep.log 2>&1
cat gpurun_out/probe_pair_sweep.log
