#!/bin/bash
# GPU session (1 GPU): -m gpu suite (incl. fused output-process tests), bench with the output process fused
# into dec_conv0 (default) and as a separate pass (A/B), smoke.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
OIDN_B200_FUSE_OUTPUT=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; tail -3 gpurun_out/bench_unfused.err; cat gpurun_out/bench_unfused.json
