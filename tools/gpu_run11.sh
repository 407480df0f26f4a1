#!/bin/bash
# GPU session (8 GPUs): sharded parity check + torchrun bench at world 4 and 8 (+2 for the series).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus8.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tools/sharded_check.py > gpurun_out/sharded_check_$N.log 2>&1; tail -2 gpurun_out/sharded_check_$N.log
done
for N in 8 4 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err; tail -1 gpurun_out/bench_n$N.json
done
