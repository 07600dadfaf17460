#!/bin/bash
# Round 2, first GPU pass: all GPU tests, chain leg alone at N=1 (graph vs plain launches), single-trajectory latency with the graph
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -16
timeout 600 python bench.py --workload chain --steps 3 --warmup 2 > gpurun_out/chain_n1.log 2> gpurun_out/chain_n1.err; tail -1 gpurun_out/chain_n1.log | cut -c1-2500; tail -3 gpurun_out/chain_n1.err
timeout 600 python tools/diag_single_latency.py > gpurun_out/single_latency.log 2>&1; tail -4 gpurun_out/single_latency.log
