#!/bin/bash
# GPU session (1 GPU): graph-replay test, all BASELINE configs (eager and graph replay), parity at full size.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_filter_gpu.py -m gpu -q -x -k "graph or async or tiled" > gpurun_out/pytest_graph.log 2>&1; tail -5 gpurun_out/pytest_graph.log
timeout 900 python tools/bench_configs.py --parity --parity-budget 30 > gpurun_out/configs_eager.jsonl 2> gpurun_out/configs_eager.err; tail -3 gpurun_out/configs_eager.err; cat gpurun_out/configs_eager.jsonl
timeout 600 python tools/bench_configs.py --graph --only 1,2,5 > gpurun_out/configs_graph.jsonl 2> gpurun_out/configs_graph.err; tail -3 gpurun_out/configs_graph.err; cat gpurun_out/configs_graph.jsonl
