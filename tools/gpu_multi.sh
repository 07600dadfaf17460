#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): NCCL chain check, MC bench and chain bench at 1..N GPUs.  usage: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_chain.py -m gpu -x -q > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_chain.log
tail -4 gpurun_out/pytest_chain.log
timeout 600 $TR --nproc-per-node $N --master-port 29511 tools/chain_nccl_check.py > gpurun_out/chain_check_n$N.log 2>&1; grep '^{' gpurun_out/chain_check_n$N.log | cut -c1-500
timeout 600 python bench.py --workload chain --steps 2 --warmup 1 > gpurun_out/chain_n1.log 2>&1; tail -1 gpurun_out/chain_n1.log | cut -c1-1500
n=2
while [ $n -le $N ]; do
  timeout 600 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --workload chain --steps 2 --warmup 1 > gpurun_out/chain_n$n.log 2>&1
  tail -1 gpurun_out/chain_n$n.log | cut -c1-1500
  timeout 900 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 2 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/mc_n$n.log 2>&1
  tail -1 gpurun_out/mc_n$n.log | cut -c1-700
  n=$((n*2))
done
