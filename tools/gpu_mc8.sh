#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1200 $TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --steps 3 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/mc_n8.log 2>&1
tail -1 gpurun_out/mc_n8.log | cut -c1-900
