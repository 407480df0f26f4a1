#!/bin/bash
# N-GPU check + bench of the collective-free sharded exchange. Usage: tools/run_peer_flags.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/peer_check_$N.log 2>&1
echo "check rc=$?" >> gpurun_out/peer_check_$N.log
tail -8 gpurun_out/peer_check_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 --no-8k --no-single-process > gpurun_out/peer_bench_$N.log 2>&1
echo "bench rc=$?"
tail -1 gpurun_out/peer_bench_$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'nccl', d.get('exchange_nccl'), 'e2e', d['e2e'] and d['e2e']['ms_per_step'], 'alt', d.get('tile_plan_alternative'), 'rank0', d.get('frame_on_rank0'))
print('roof', d['roofline']['conv_ms_per_frame'], d['roofline']['frac'])
"
