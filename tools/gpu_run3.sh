#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/sharded_check_$N.log 2>&1; tail -5 gpurun_out/sharded_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
timeout 600 python bench.py --steps 20 --warmup 5 --width 7680 --height 4320 --no-cpu-baseline > gpurun_out/bench_8k_n1.json 2> gpurun_out/bench_8k_n1.err; tail -3 gpurun_out/bench_8k_n1.err; cat gpurun_out/bench_8k_n1.json
