#!/bin/bash
P=tools/bin/probe_conv
mkdir -p gpurun_out
run() { echo "--- $*"; timeout 120 $P "$@"; echo "exit=$?"; }
{
run 24 300 48 0 48 0 0 0
run 16 256 64 0 64 1 0 0
run 70 200 112 48 96 0 1 0
run 16 256 96 0 96 0 0 0
run 1088 1920 64 16 64 0 1 0 20
run 1088 1920 32 0 32 1 0 0 20
run 1088 1920 64 0 32 0 0 0 20
run 1088 1920 16 0 32 0 0 0 20
run 1088 1920 32 0 16 0 0 0 20
run 544 960 128 0 64 0 0 0 20
run 544 960 64 0 64 0 0 0 20
run 272 480 160 0 96 0 0 0 20
} 2>&1 | tee gpurun_out/probe2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 3 -c 1 -f -o gpurun_out/prof_conv_1a $P 1088 1920 64 16 64 0 1 0 5 > gpurun_out/ncu_1a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 3 -c 1 -f -o gpurun_out/prof_conv_e0 $P 1088 1920 16 0 32 0 0 0 5 > gpurun_out/ncu_e0.log 2>&1
ls -la gpurun_out
