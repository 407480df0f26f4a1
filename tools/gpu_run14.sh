#!/bin/bash
# GPU session (1 GPU): full -m gpu suite, smoke, bench, ncu launch list of the bench command, ncu --set full of
# the bandwidth-bound passes (input process, output process, autoexposure) of one 4K frame.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"input_process|output_process|autoexposure" -s 12 -c 4 -o gpurun_out/elementwise_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/elementwise_ncu.log 2>&1; tail -2 gpurun_out/elementwise_ncu.log
ncu -i gpurun_out/elementwise_full.ncu-rep --page raw --csv > gpurun_out/elementwise_full_raw.csv 2>/dev/null; wc -c gpurun_out/elementwise_full_raw.csv
