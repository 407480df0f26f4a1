#!/bin/bash
# tcgen05.mma issue cost vs N, issuing warps and accumulator reuse (tools/probe_mma_issue.cu). ~1 s per line.
P=tools/bin/probe_mma_issue
mkdir -p gpurun_out
{
for N in 32 48 64 96 128 192 256; do timeout 30 $P $N 1 1 64; done
for N in 48 96 192; do timeout 30 $P $N 2 1 64; done
for N in 48 96; do timeout 30 $P $N 4 1 64; done
for N in 48 96; do timeout 30 $P $N 1 2 64; timeout 30 $P $N 1 4 64; done
timeout 30 $P 96 2 2 64
timeout 30 $P 96 1 1 8
timeout 30 $P 96 1 1 256
} 2>&1 | tee gpurun_out/probe_mma_issue.log
