#!/bin/bash
# GPU session (1 GPU): full bench line, ncu launch list of the bench command, ncu --set full of the 16 conv
# launches of one 4K frame (DRAM traffic per launch -> profiles/conv_traffic.json), all BASELINE configs.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 48 -c 16 -o gpurun_out/conv_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/conv_ncu.log 2>&1; tail -2 gpurun_out/conv_ncu.log
ncu -i gpurun_out/conv_full.ncu-rep --page raw --csv > gpurun_out/conv_full_raw.csv 2>/dev/null; wc -c gpurun_out/conv_full_raw.csv
rm -f gpurun_out/conv_full.ncu-rep
timeout 900 python tools/bench_configs.py > gpurun_out/configs_eager.jsonl 2> gpurun_out/configs_eager.err; tail -3 gpurun_out/configs_eager.err; cut -c1-400 gpurun_out/configs_eager.jsonl
