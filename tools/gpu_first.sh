#!/bin/bash
# First GPU pass: microbenchmarks, parity tests, bench line, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc > gpurun_out/nproc.txt
./tools/bin/fp64_micro > gpurun_out/fp64_micro.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --trials 16 --steps 2 --warmup 3 > gpurun_out/bench_t16.log 2>&1
BCMD="python bench.py --trials 4 --steps 1 --warmup 1 --duration 60 --no-cpu-baseline --no-latency"
$BCMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu1.log 2>&1
$BCMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_seg_eliminate -s 4 -c 2 -o gpurun_out/prof_elim $BCMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/fp64_micro.txt; tail -1 gpurun_out/bench_t16.log
