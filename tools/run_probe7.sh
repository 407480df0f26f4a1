#!/bin/bash
# Bias through the constant bank (default) vs through shared memory (OIDN_B200_SMEM_BIAS=1) on the 4K layers.
P=tools/bin/probe_conv
mkdir -p gpurun_out
run() { echo "--- [$MODE] $*"; timeout 120 $P "$@" 2>&1 | grep -E "RESULT|TIME|error|timeout" | cut -c1-200; }
runp() { echo "--- [$MODE] pair $*"; timeout 120 tools/bin/probe_pair "$@" 2>&1 | grep -E "RESULT|TIME|error|timeout" | cut -c1-200; }
shapes() {
run 2160 3840 64 16 64 0 1 0 20
run 1080 1920 96 32 64 0 1 0 20
run 1080 1920 64 0 64 0 0 0 20
run 540 960 112 48 96 0 1 0 20
run 540 960 96 0 96 0 0 0 20
run 1080 1920 32 0 48 1 0 0 20
run 540 960 48 0 64 1 0 0 20
runp 2160 3840 16 32 32 1 20
runp 2160 3840 64 32 16 0 20
}
{
MODE=const; shapes
MODE=smem; export OIDN_B200_SMEM_BIAS=1; shapes
} > gpurun_out/probe7.log 2>&1
cat gpurun_out/probe7.log
