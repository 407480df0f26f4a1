#!/bin/bash
# Wait-time traces of the conv kernel on the 4K layer shapes (H W C1 C2 Cout post up mode iters)
P=tools/bin/probe_conv_trace
mkdir -p gpurun_out
run() { echo "--- $*"; PROBE_TRACE=1 timeout 120 $P "$@"; echo "exit=$?"; }
{
run 2160 3840 16 0 32 0 0 0 10
run 2160 3840 32 0 32 1 0 0 10
run 2160 3840 32 0 16 0 0 0 10
run 2160 3840 64 0 32 0 0 0 10
run 2160 3840 64 16 64 0 1 0 10
run 1080 1920 96 32 64 0 1 0 10
run 1080 1920 64 0 64 0 0 0 10
run 540 960 112 48 96 0 1 0 10
run 135 240 80 0 96 0 0 0 10
} 2>&1 | tee gpurun_out/trace.log
