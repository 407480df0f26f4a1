#!/bin/bash
# GPU session (N GPUs): -m gpu suite, sharded parity check (rank-0 frame, staged, distributed) + torchrun bench at world = $1
N=${1:-2}
mkdir -p gpurun_out
df -h /dev/shm | tail -1 > gpurun_out/shm.txt; cat gpurun_out/shm.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/sharded_check_$N.log 2>&1; tail -4 gpurun_out/sharded_check_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err; tail -1 gpurun_out/bench_n$N.json
