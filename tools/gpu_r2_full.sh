#!/bin/bash
# Round 2: whole GPU suite, smoke, reference arm (short), default bench line (with the chain and waves legs)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=12 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | grep -i "60 s dataset\|16 full-size\|chain vs single\|FAILED\|passed\|failed\|Error" | tail -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c1-300; echo "default bench took $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/bench_default.err
