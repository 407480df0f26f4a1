#!/bin/bash
# GPU session (1 GPU), round-end validation of the committed state: -m gpu suite, smoke, the bench line,
# full-size parity of the BASELINE configs against the oracle (new SFU transfer functions + fused output).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python tools/bench_configs.py --parity --parity-budget 30 --only 1,2,4,41,5 --steps 10 > gpurun_out/configs_parity.jsonl 2> gpurun_out/configs_parity.err; tail -3 gpurun_out/configs_parity.err; cut -c1-120 gpurun_out/configs_parity.jsonl; grep -o '"parity": {[^}]*}' gpurun_out/configs_parity.jsonl
