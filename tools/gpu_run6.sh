#!/bin/bash
# GPU session: parity tests, wait traces of the 4K layer shapes, bench line.
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_filter_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
P=tools/bin/probe_conv_trace
run() { echo "--- $*"; PROBE_TRACE=1 timeout 120 $P "$@"; echo "exit=$?"; }
{
run 2160 3840 16 0 32 0 0 0 10
run 2160 3840 32 0 32 1 0 0 10
run 2160 3840 32 0 16 0 0 0 10
run 2160 3840 64 0 32 0 0 0 10
run 2160 3840 64 16 64 0 1 0 10
run 1080 1920 96 32 64 0 1 0 10
run 540 960 112 48 96 0 1 0 10
} > gpurun_out/trace3.log 2>&1
grep -E "^cfg|TIME|RESULT" gpurun_out/trace3.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
