#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_filter_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
