#!/bin/bash
# tap-packed last conv: kernel-level tests + a 4K frame bench with the per-layer conv times
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "tap_packed" 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-8k --no-kernel-to-beat > gpurun_out/bench_tap.json 2> gpurun_out/bench_tap.err
tail -c 300 gpurun_out/bench_tap.err
python - <<EOF
import json
d=json.loads(open("gpurun_out/bench_tap.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["conv_ms_per_frame"], d["roofline"]["frac"], d["roofline"]["sustained"]["ms_per_step"], d["e2e"]["ms_per_step"])
print(d["passes"]["conv_layers_ms"])
EOF
