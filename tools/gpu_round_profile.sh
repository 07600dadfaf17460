#!/bin/bash
# bench line + ncu launch list of the SAME command (short form), for profiles/.  usage: tools/gpu_round_profile.sh <tag>
TAG=$1
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-400
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency"
$BCMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu1_$TAG.log 2>&1
tail -2 gpurun_out/ncu1_$TAG.log | cut -c1-200
