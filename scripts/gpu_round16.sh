#!/bin/bash
mkdir -p gpurun_out
{
python scripts/diag_group_pool.py
DIAG_BIG=0 python scripts/diag_group_pool.py
for b in nccl gloo; do for big in 1 0; do
DIAG_BACKEND=$b DIAG_BIG=$big python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 scripts/diag_group_pool.py 2>&1 | grep -E "world=|Error|error" | head -5
done; done
} > gpurun_out/diag_group_pool.txt 2>&1
cat gpurun_out/diag_group_pool.txt | cut -c1-300
