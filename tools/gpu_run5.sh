#!/bin/bash
# Fine-grained wait traces of the HBM-bound layer shapes + one ncu --set full pass over the 16 conv launches of a 4K frame
P=tools/bin/probe_conv_trace
mkdir -p gpurun_out
run() { echo "--- $*"; PROBE_TRACE=1 timeout 120 $P "$@"; echo "exit=$?"; }
{
run 2160 3840 16 0 32 0 0 0 10
run 2160 3840 32 0 32 1 0 0 10
run 2160 3840 32 0 16 0 0 0 10
run 2160 3840 64 0 32 0 0 0 10
run 1080 1920 96 32 64 0 1 0 10
} 2>&1 | tee gpurun_out/trace2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 64 -c 16 -o gpurun_out/prof_conv16 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu16.log 2>&1
tail -3 gpurun_out/ncu16.log
ncu -i gpurun_out/prof_conv16.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_uniform.sum > gpurun_out/conv16_raw.csv 2>&1
head -c 3000 gpurun_out/conv16_raw.csv
