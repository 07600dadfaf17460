#!/bin/bash
# per-launch times of the two preintegration kernels (16 full-size trials) for the default library and tools/bin/lib_*.so given
B="python bench.py --trials 16 --steps 1 --warmup 0 --no-latency --no-cpu-baseline --no-sim --no-chain --no-pipeline --waves-trials 0"
for L in default "$@"; do
  echo "== $L"
  if [ "$L" != default ]; then export V2GT_LIB_CUDA=$PWD/tools/bin/$L; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_preint -c 2 --csv $B 2>/dev/null | grep k_preint | awk -F'","' '{print $5, $(NF)}'
done
