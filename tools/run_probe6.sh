#!/bin/bash
# Direct stores vs staged TMA stores on the wide 4K layers (shared-memory pipe pressure), and stream counts.
P=tools/bin/probe_conv
mkdir -p gpurun_out
run() { echo "--- [$MODE] $*"; timeout 120 $P "$@" 2>&1 | grep -E "^cfg|RESULT|TIME|error|timeout" | cut -c1-200; }
shapes() {
run 2160 3840 64 16 64 0 1 0 20
run 2160 3840 64 0 32 0 0 0 20
run 1080 1920 96 32 64 0 1 0 20
run 1080 1920 64 0 64 0 0 0 20
run 540 960 112 48 96 0 1 0 20
run 540 960 96 0 96 0 0 0 20
run 1080 1920 32 0 48 1 0 0 20
run 2160 3840 16 0 32 0 0 0 20
run 2160 3840 32 0 32 1 0 0 20
}
{
MODE=default; shapes
MODE=direct; export OIDN_B200_DIRECT_STORE=1; shapes
} > gpurun_out/probe6.log 2>&1
cat gpurun_out/probe6.log
