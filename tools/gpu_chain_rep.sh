#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
for S in 60 60 60 120 40 80; do
  V2GT_CHAIN_CHECK_S=$S timeout 300 $TR --master-port 2961$((RANDOM % 10)) tools/chain_nccl_check.py 2>&1 | grep '^{' | cut -c1-400
done
