#!/bin/bash
# Wait-time trace of the fused conv pairs on the 4K shapes (OIDN_B200_TRACE build of the kernels).
P=tools/bin/probe_pair_trace
mkdir -p gpurun_out
export PROBE_TRACE=1
{
for cfg in "2160 3840 16 32 32 1 5" "2160 3840 64 32 16 0 5"; do echo "--- $cfg"; timeout 60 $P $cfg 2>&1 | tail -16; done
} > gpurun_out/probe_pair_trace.log 2>&1
cat gpurun_out/probe_pair_trace.log
