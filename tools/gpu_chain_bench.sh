#!/bin/bash
# chain workload (config 5) at 1..N GPUs.  usage: tools/gpu_chain_bench.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --workload chain --steps 3 --warmup 2 > gpurun_out/chain_n1.log 2>&1; tail -1 gpurun_out/chain_n1.log | cut -c1-1500
n=2
while [ $n -le $N ]; do
  timeout 600 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --workload chain --steps 3 --warmup 2 > gpurun_out/chain_n$n.log 2>&1
  tail -1 gpurun_out/chain_n$n.log | cut -c1-1500
  n=$((n*2))
done
