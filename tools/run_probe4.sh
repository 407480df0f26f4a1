#!/bin/bash
# Per-layer A/B of the 4K base-UNet conv shapes: staging + TMA store vs direct-store epilogue, 1 vs 2 streams.
P=tools/bin/probe_conv
mkdir -p gpurun_out
run() { echo "--- [$MODE] $*"; timeout 120 $P "$@" 2>&1 | grep -E "^cfg|RESULT|TIME|error|timeout"; }
layers() {
run 2160 3840 16 0 32 0 0 0 20
run 2160 3840 32 0 32 1 0 0 20
run 1080 1920 32 0 48 1 0 0 20
run 540 960 48 0 64 1 0 0 20
run 270 480 64 0 80 1 0 0 20
run 136 240 80 0 96 0 0 0 20
run 136 240 96 0 96 0 0 0 20
run 270 480 96 64 112 0 1 0 20
run 270 480 112 0 112 0 0 0 20
run 540 960 112 48 96 0 1 0 20
run 540 960 96 0 96 0 0 0 20
run 1080 1920 96 32 64 0 1 0 20
run 1080 1920 64 0 64 0 0 0 20
run 2160 3840 64 16 64 0 1 0 20
run 2160 3840 64 0 32 0 0 0 20
run 2160 3840 32 0 16 0 0 0 20
}
{
MODE=default; layers
MODE=direct; export OIDN_B200_DIRECT_STORE=1; layers
MODE=direct_s1; export OIDN_B200_STREAMS=1; layers
MODE=staged_s1; export OIDN_B200_DIRECT_STORE=0; layers
MODE=staged_s2; export OIDN_B200_STREAMS=2; layers
MODE=direct_s2; export OIDN_B200_DIRECT_STORE=1; layers
} > gpurun_out/probe4.log 2>&1
grep -c PASS gpurun_out/probe4.log; grep -c FAIL gpurun_out/probe4.log
