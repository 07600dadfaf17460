#!/bin/bash
# 8-GPU validation: MC bench (weak scaling, 128 trials per GPU = the 1024-trial batch) and the chain workload.
mkdir -p gpurun_out
nproc > gpurun_out/nproc8.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1200 $TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --steps 2 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/mc_n8.log 2>&1
tail -1 gpurun_out/mc_n8.log | cut -c1-900
timeout 600 $TR --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --workload chain --steps 3 --warmup 2 > gpurun_out/chain_n8.log 2>&1
tail -1 gpurun_out/chain_n8.log | cut -c1-1200
timeout 600 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --workload chain --steps 3 --warmup 2 > gpurun_out/chain_n4.log 2>&1
tail -1 gpurun_out/chain_n4.log | cut -c1-1200
cat gpurun_out/nproc8.txt
