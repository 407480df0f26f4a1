#!/bin/bash
# Fused conv pairs against the two-launch path (bit-identical) + timing. One process per configuration.
P=tools/bin/probe_pair
mkdir -p gpurun_out
run() { echo "--- $*"; timeout 60 $P "$@" 2>&1 | tail -6; echo "exit=$?"; }
{
run 40 200 16 32 32 1
run 40 200 64 32 16 0
run 37 131 16 32 32 0
run 300 1000 16 32 32 1 5
run 300 1000 64 32 16 0 5
run 300 1000 32 32 16 0 5
run 128 400 16 64 64 1 5
run 128 400 64 64 16 0 5
run 2160 3840 16 32 32 1 20
run 2160 3840 64 32 16 0 20
run 2160 3840 32 32 16 0 20
} > gpurun_out/probe_pair.log 2>&1
cat gpurun_out/probe_pair.log
