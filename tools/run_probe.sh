#!/bin/bash
# Runs the conv probe configurations one process each (a trap poisons the CUDA context).
P=tools/bin/probe_conv
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
run() { echo "--- $*"; timeout 90 $P "$@"; echo "exit=$?"; }
{
run 8 128 64 0 64 0 0 2
run 8 128 64 0 64 0 0 0
run 8 128 64 0 64 0 0 1
run 8 128 32 0 32 0 0 2
run 8 128 32 0 32 0 0 0
run 8 128 32 0 32 0 0 1
run 8 128 16 0 32 0 0 2
run 8 128 16 0 32 0 0 0
run 8 128 16 0 32 0 0 1
run 24 300 48 0 48 0 0 0
run 16 256 64 0 64 1 0 0
run 16 256 64 16 64 0 0 0
run 16 256 64 16 64 0 1 0
run 16 256 64 0 64 2 0 0
run 16 256 96 0 96 0 0 0
run 16 128 160 0 112 0 0 0
run 16 128 32 0 16 0 0 0
run 70 200 112 48 96 0 1 0
run 1088 1920 64 16 64 0 0 0 20
run 1088 1920 64 16 64 0 1 0 20
run 1088 1920 32 0 32 1 0 0 20
run 1088 1920 64 0 32 0 0 0 20
run 1088 1920 16 0 32 0 0 0 20
run 1088 1920 32 0 16 0 0 0 20
run 544 960 128 0 64 0 0 0 20
run 544 960 64 0 64 0 0 0 20
run 272 480 160 0 96 0 0 0 20
} 2>&1 | tee gpurun_out/probe1.log
