#!/bin/bash
# GPU session (1 GPU): -m gpu suite, conv tests again under the direct-store / stream overrides, per-layer probe matrix.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for env in "OIDN_B200_DIRECT_STORE=1" "OIDN_B200_DIRECT_STORE=1 OIDN_B200_STREAMS=1" "OIDN_B200_DIRECT_STORE=0 OIDN_B200_STREAMS=1" "OIDN_B200_DIRECT_STORE=1 OIDN_B200_STREAMS=2"; do
  echo "== $env"; env $env timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_filter_gpu.py -m gpu -q -x -k "conv or golden or tiled" 2>&1 | tail -3
done > gpurun_out/pytest_variants.log 2>&1; cat gpurun_out/pytest_variants.log
bash tools/run_probe4.sh
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
