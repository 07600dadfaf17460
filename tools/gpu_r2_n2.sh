#!/bin/bash
# 2-GPU pass: NCCL chain test (in-library ncclAllGather + CUDA graph), chain bench at N=1 and N=2, default bench line at N=2 (short)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_chain.py -m gpu -x -q > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_chain.log
grep -v "^$" gpurun_out/pytest_chain.log | tail -6
timeout 600 python bench.py --workload chain --steps 3 --warmup 2 > gpurun_out/chain_n1.log 2> gpurun_out/chain_n1.err; tail -1 gpurun_out/chain_n1.log | cut -c1-1200
timeout 600 $TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --workload chain --steps 3 --warmup 2 > gpurun_out/chain_n2.log 2> gpurun_out/chain_n2.err; tail -1 gpurun_out/chain_n2.log | cut -c1-2500; tail -3 gpurun_out/chain_n2.err
