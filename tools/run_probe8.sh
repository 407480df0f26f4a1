#!/bin/bash
# Row folding of upsampled sources (ConvKernelParams::up_fold) against the SIMT witness, and against the unfolded path.
P=tools/bin/probe_conv
mkdir -p gpurun_out
run() { echo "--- [$MODE] $*"; timeout 120 $P "$@" 2>&1 | grep -E "^cfg|RESULT|TIME|error|timeout|trap|CUDA" | cut -c1-220; }
small() {
run 16 96 96 64 112 0 1 0
run 32 136 112 48 96 0 1 0
run 32 260 96 32 64 0 1 0
run 32 260 64 16 64 0 1 0
run 18 70 256 128 192 0 1 0
run 6 130 64 0 32 0 1 0
run 2 2 32 0 32 0 1 0
run 34 300 64 16 64 0 1 0
}
big() {
run 2160 3840 64 16 64 0 1 0 20
run 1080 1920 96 32 64 0 1 0 20
run 540 960 112 48 96 0 1 0 20
run 270 480 96 64 112 0 1 0 20
}
{
MODE=planner; small; big
MODE=forcefold; export OIDN_B200_FORCE_UPFOLD=1; small; big; unset OIDN_B200_FORCE_UPFOLD
MODE=nofold; export OIDN_B200_NO_UPFOLD=1; big
} > gpurun_out/probe8.log 2>&1
cat gpurun_out/probe8.log
