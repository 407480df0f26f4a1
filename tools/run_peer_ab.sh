#!/bin/bash
# N=2: correctness of the kernel-scattered bin exchange, then e2e A/B peer vs nccl on one box
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/sharded_check.py > gpurun_out/peer2_check.log 2>&1
echo "check rc=$?"; grep -a "distributed frame\|bit-identical" gpurun_out/peer2_check.log
F="--gpus 2 --steps 20 --warmup 5 --no-8k --no-single-process --no-rank0"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py $F > gpurun_out/peer2_on.log 2>&1
OIDN_B200_EXCHANGE=nccl timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py $F > gpurun_out/peer2_nccl.log 2>&1
for n in on nccl; do tail -1 gpurun_out/peer2_$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$n', d['ms_per_step'], d['e2e']['ms_per_step'])
"; done
