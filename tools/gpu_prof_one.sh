#!/bin/bash
# ncu --set full capture of ONE kernel (2 launches) on a 16-trial batch.  usage: tools/gpu_prof_one.sh <kernel-regex> <tag> [skip] [bench args]
K=$1; TAG=$2; SKIP=${3:-3}; shift 3
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --trials 16 $*"
$BCMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -o gpurun_out/prof_$TAG -f $BCMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/ncu2_$TAG.log
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
