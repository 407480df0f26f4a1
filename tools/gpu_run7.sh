#!/bin/bash
# GPU session: parity tests, L2-prefetch A/B on the shallow-ring layer shapes, bench line.
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_filter_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
P=tools/bin/probe_conv
run() { echo "--- PF=$OIDN_B200_PREFETCH $*"; timeout 120 $P "$@" | grep -E "^cfg|TIME|FAIL"; }
{
for pf in 0 2 3 5; do
export OIDN_B200_PREFETCH=$pf
run 2160 3840 64 16 64 0 1 0 20
run 1080 1920 96 32 64 0 1 0 20
run 540 960 112 48 96 0 1 0 20
run 1080 1920 64 0 64 0 0 0 20
run 2160 3840 64 0 32 0 0 0 20
run 270 480 160 0 112 0 0 0 20
done
unset OIDN_B200_PREFETCH
} > gpurun_out/prefetch_ab.log 2>&1
grep -E "^---|TIME" gpurun_out/prefetch_ab.log | paste - - | awk '{print $2, $3, $4, $5, $6, $7, $8, $12, $13, $14, $15}'
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
