#!/bin/bash
# HEAD pass: whole GPU suite, smoke, reference arm (short), default bench line, ncu capture of the elimination
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | grep -i "60 s dataset\|16 full-size\|FAILED\|passed\|failed\|Error\|rc=" | tail -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
T0=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 --ref-budget-s 60 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.log | cut -c1-300; echo "reference arm took $(( $(date +%s) - T0 )) s"
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c1-300; echo "default bench took $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/bench_default.err
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-sim --no-chain --no-pipeline --waves-trials 0 --trials 16"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seg_eliminate -s 6 -c 4 -o gpurun_out/prof_r02d_k_seg_eliminate -f $BCMD > gpurun_out/ncu2_r02d.log 2>&1; tail -1 gpurun_out/ncu2_r02d.log | cut -c1-120
ncu -i gpurun_out/prof_r02d_k_seg_eliminate.ncu-rep --page raw --csv > gpurun_out/raw_r02d_k_seg_eliminate.csv 2>/dev/null
