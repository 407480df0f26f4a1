#!/bin/bash
# Four streams (CoutG <= 32) vs two, and the planner's new defaults on the concat layers (4K base-UNet shapes).
P=tools/bin/probe_conv
mkdir -p gpurun_out
run() { echo "--- [$MODE] $*"; timeout 120 $P "$@" 2>&1 | grep -E "^cfg|RESULT|TIME|error|timeout"; }
narrow() {
run 2160 3840 16 0 32 0 0 0 20
run 2160 3840 32 0 32 1 0 0 20
run 2160 3840 64 0 32 0 0 0 20
run 2160 3840 32 0 16 0 0 0 20
}
wide() {
run 270 480 96 64 112 0 1 0 20
run 270 480 112 0 112 0 0 0 20
run 540 960 112 48 96 0 1 0 20
run 1080 1920 96 32 64 0 1 0 20
run 2160 3840 64 16 64 0 1 0 20
}
{
MODE=default; narrow; wide
MODE=streams2; export OIDN_B200_STREAMS=2; narrow
MODE=streams1; export OIDN_B200_STREAMS=1; wide
} > gpurun_out/probe5.log 2>&1
grep -c PASS gpurun_out/probe5.log; grep -c FAIL gpurun_out/probe5.log
