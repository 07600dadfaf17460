#!/bin/bash
# Evidence pass for the kernels changed after tools/gpu_r2_profiles.sh ran (tag r02e): ncu --set full of the two
# preintegration kernels (16 full-size trials), raw counter pages as CSV, and the launch list of the default bench command.
mkdir -p gpurun_out
TAG=${1:-r02e}
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-sim --no-chain --waves-trials 0 --trials 16"
$BCMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
for SPEC in "k_preintegrate:1" "k_preint_cov:1"; do
  K=${SPEC%%:*}; SKIP=${SPEC##*:}; N=$(echo $K | tr -cd 'a-z0-9_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -o gpurun_out/prof_${TAG}_$N -f $BCMD > gpurun_out/ncu2_${TAG}_$N.log 2>&1
  echo "$N: $(tail -1 gpurun_out/ncu2_${TAG}_$N.log | cut -c1-120)"
  ncu -i gpurun_out/prof_${TAG}_$N.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$N.csv 2>/dev/null
done
DCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --waves-trials 0"
$DCMD > gpurun_out/plain_default_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv $DCMD > gpurun_out/ncu1_$TAG.log 2>&1
tail -1 gpurun_out/ncu1_$TAG.log | cut -c1-200
du -sh gpurun_out
