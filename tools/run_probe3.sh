#!/bin/bash
P=tools/bin/probe_conv
mkdir -p gpurun_out
run() { echo "--- $*"; timeout 120 $P "$@"; echo "exit=$?"; }
{
run 8 128 64 0 64 0 0 0
run 8 128 32 0 32 0 0 0
run 8 128 16 0 32 0 0 0
run 24 300 48 0 48 0 0 0
run 16 256 64 0 64 1 0 0
run 38 300 32 0 48 1 0 0
run 16 256 64 16 64 0 1 0
run 16 256 96 0 96 0 0 0
run 16 128 160 0 112 0 0 0
run 16 128 32 0 112 0 0 0
run 16 128 32 0 16 0 0 0
run 70 200 112 48 96 0 1 0
run 34 100 256 128 192 0 1 0
run 1088 1920 64 16 64 0 1 0 20
run 1088 1920 32 0 32 1 0 0 20
run 1088 1920 64 0 32 0 0 0 20
run 1088 1920 16 0 32 0 0 0 20
run 1088 1920 32 0 16 0 0 0 20
run 544 960 96 32 64 0 1 0 20
run 544 960 64 0 64 0 0 0 20
run 544 960 32 0 48 1 0 0 20
run 272 480 112 48 96 0 1 0 20
run 272 480 96 0 96 0 0 0 20
run 136 240 96 64 112 0 1 0 20
run 68 120 80 0 96 0 0 0 20
} 2>&1 | tee gpurun_out/probe3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 3 -c 1 -f -o gpurun_out/prof_conv_1a $P 1088 1920 64 16 64 0 1 0 5 > gpurun_out/ncu_1a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 3 -c 1 -f -o gpurun_out/prof_conv_1b $P 1088 1920 64 0 32 0 0 0 5 > gpurun_out/ncu_1b.log 2>&1
