#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1200 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.log | cut -c1-300; tail -2 gpurun_out/bench_n2.err | cut -c1-200
timeout 900 python -m pytest tests/test_chain.py -m gpu -q -s 2>&1 | grep -i "passed\|failed\|FAILED" | cut -c1-200
