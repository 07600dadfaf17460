#!/bin/bash
# ncu --set full capture of every hot kernel (2 launches each) on a 16-trial batch + launch list.  usage: tools/gpu_prof_all.sh <tag> [bench args]
TAG=$1; shift
mkdir -p gpurun_out
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --trials 16 $*"
$BCMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu1_$TAG.log 2>&1
for K in k_assemble k_factor_eval k_seg_backsub k_seg_eliminate k_preintegrate; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 2 -o gpurun_out/prof_${TAG}_$K -f $BCMD > gpurun_out/ncu2_${TAG}_$K.log 2>&1
  tail -1 gpurun_out/ncu2_${TAG}_$K.log
done
tail -1 gpurun_out/plain_$TAG.log | cut -c1-400
