#!/bin/bash
# GPU session (1 GPU): full gpu test tier, 4K bench, 8K single-tile bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --steps 10 --warmup 3 --width 7680 --height 4320 --no-cpu-baseline > gpurun_out/bench_8k_n1.json 2> gpurun_out/bench_8k_n1.err; tail -3 gpurun_out/bench_8k_n1.err; cat gpurun_out/bench_8k_n1.json
