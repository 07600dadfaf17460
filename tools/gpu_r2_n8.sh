#!/bin/bash
# 8-GPU pass: the default bench line at N = 8 (Monte-Carlo batch sharded by trial + the chain leg over 8 GPUs), chain alone at N = 4
mkdir -p gpurun_out
nproc > gpurun_out/nproc8.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1500 $TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err
tail -1 gpurun_out/bench_n8.log | cut -c1-400; tail -3 gpurun_out/bench_n8.err | cut -c1-300
timeout 600 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --workload chain --steps 3 --warmup 2 > gpurun_out/chain_n4.log 2> gpurun_out/chain_n4.err
tail -1 gpurun_out/chain_n4.log | cut -c1-300
