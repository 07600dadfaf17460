#!/bin/bash
# Round 2 evidence pass at HEAD: ncu --set full of every kernel of the hot path (2 launches each, 16 full-size trials),
# the launch list of the default bench command (short form), compute-sanitizer on the smoke case.
mkdir -p gpurun_out
TAG=${1:-r02}
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-sim --no-chain --waves-trials 0 --trials 16"
$BCMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
for SPEC in k_init_states:0 k_compact:0 k_preintegrate:1 "k_factor_eval:6" k_assemble:3 k_reduce_gamma:3 k_seg_eliminate:6 k_reduced_assemble:6 "k_reduced[(]:6" k_l2_prepare:3 k_sep_delta:3 k_seg_backsub:6 k_lm_decide:3; do
  K=${SPEC%%:*}; SKIP=${SPEC##*:}; N=$(echo $K | tr -cd 'a-z0-9_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 4 -o gpurun_out/prof_${TAG}_$N -f $BCMD > gpurun_out/ncu2_${TAG}_$N.log 2>&1
  echo "$N: $(tail -1 gpurun_out/ncu2_${TAG}_$N.log | cut -c1-120)"
  # the reports together exceed what a call may bring back (64 MiB): keep the raw counter page of every kernel as CSV and
  # the report itself only for the elimination (source page)
  ncu -i gpurun_out/prof_${TAG}_$N.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$N.csv 2>/dev/null
  [ "$N" = "k_seg_eliminate" ] || rm -f gpurun_out/prof_${TAG}_$N.ncu-rep
done
# launch list of the default bench command (short form: 1 + 1 steps), after the command has run clean without ncu
DCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --waves-trials 0"
$DCMD > gpurun_out/plain_default_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv $DCMD > gpurun_out/ncu1_$TAG.log 2>&1
tail -1 gpurun_out/ncu1_$TAG.log | cut -c1-200
# (compute-sanitizer is closed on this GPU pool -- the wrapper refuses to run it -- so no memcheck log can be produced here)
du -sh gpurun_out
