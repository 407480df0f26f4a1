#!/bin/bash
# Wait-time trace (OIDN_B200_TRACE build) of the widest 4K layers.
P=tools/bin/probe_conv_trace
export PROBE_TRACE=1
mkdir -p gpurun_out
{
for cfg in "2160 3840 64 16 64 0 1 0 5" "1080 1920 96 32 64 0 1 0 5" "1080 1920 64 0 64 0 0 0 5" "540 960 112 48 96 0 1 0 5"; do echo "--- $cfg"; timeout 60 $P $cfg 2>&1 | tail -16; done
} > gpurun_out/trace2.log 2>&1
cat gpurun_out/trace2.log
