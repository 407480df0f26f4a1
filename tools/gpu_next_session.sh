#!/bin/bash
# First GPU session of the next round (1 GPU, ~3 minutes): the measurements round 1 prepared but could not run.
#   1. tools/reference_cuda_witness.py   this backend vs the reference's own CUDA device vs the oracle (needs baseline/_ref:
#                                        OIDN_B200_BUILD_REFERENCE=1 python __graft_entry__.py in the build container first)
#   2. tools/run_probe_mma.sh            what one thread pays per tcgen05.mma (N, issuing warps, accumulator reuse)
#   3. the regular validation            -m gpu suite, smoke, bench line
#   4. tools/integration_run.sh 40 ops   the op-level device module (unmodified reference core) under the reference's
#                                        oidnBenchmark / oidnTest -- needs baseline/_b200_ops on the box: comment its line
#                                        out of .gpurunignore (62 MB) and build it first (tools/build_integration_module.sh)
# For 8 GPUs afterwards: OIDN_B200_TILE_POLICY=2 torchrun ... bench.py --gpus 8 (strip-aware tiles, 2x4 of 3936x1232)
# against the default (4x2 of 2064x2256): profiles/README.md has the 1.41 ms to compare with.
mkdir -p gpurun_out
make -s -C oidn_b200/csrc probe_mma > /dev/null 2>&1
[ -d baseline/_ref/lib ] && timeout 120 python tools/reference_cuda_witness.py > gpurun_out/reference_cuda_witness.log 2>&1; tail -5 gpurun_out/reference_cuda_witness.log
timeout 120 bash tools/run_probe_mma.sh > /dev/null 2>&1; cat gpurun_out/probe_mma_issue.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
[ -d baseline/_b200_ops/lib ] && bash tools/integration_run.sh 40 ops
