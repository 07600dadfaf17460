#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --trials 8 --duration 30 --steps 2 --warmup 1 --waves-trials 24 --chain-duration 300 > gpurun_out/bench_sanity.log 2> gpurun_out/bench_sanity.err; tail -1 gpurun_out/bench_sanity.log | cut -c1-200; tail -5 gpurun_out/bench_sanity.err
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-sim --no-chain --no-pipeline --waves-trials 0 --trials 16"
timeout 600 ncu --set full --clock-control none -k regex:'^k_reduced$' -s 6 -c 4 -o gpurun_out/prof_r02b_k_reduced -f $BCMD > gpurun_out/ncu2_r02b_k_reduced.log 2>&1; tail -1 gpurun_out/ncu2_r02b_k_reduced.log | cut -c1-120
ncu -i gpurun_out/prof_r02b_k_reduced.ncu-rep --page raw --csv > gpurun_out/raw_r02b_k_reduced.csv 2>/dev/null; rm -f gpurun_out/prof_r02b_k_reduced.ncu-rep
