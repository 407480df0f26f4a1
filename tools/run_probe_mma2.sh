#!/bin/bash
# Does a shifted (not swizzle-atom-aligned) A operand cost more tensor-pipe time? (4 issuers = pipe-bound)
P=tools/bin/probe_mma_issue
mkdir -p gpurun_out
{
for rb in 128 64 32; do for sh in 0 1; do for N in 48 96; do timeout 30 $P $N 4 1 192 $rb $sh; done; timeout 30 $P 192 2 1 192 $rb $sh; done; done
for rb in 128 64; do for sh in 0 1; do timeout 30 $P 96 2 1 192 $rb $sh; timeout 30 $P 96 1 1 192 $rb $sh; done; done
} 2>&1 | tee gpurun_out/probe_mma_shift.log
