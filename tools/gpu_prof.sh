#!/bin/bash
# ncu pass: launch list of a short bench + full capture of one kernel.  usage: tools/gpu_prof.sh <kernel-regex> <tag> [bench args]
K=$1; TAG=$2; shift 2
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency $*"
$BCMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu1_$TAG.log 2>&1
$BCMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -o gpurun_out/prof_$TAG -f $BCMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-600; tail -3 gpurun_out/ncu2_$TAG.log
