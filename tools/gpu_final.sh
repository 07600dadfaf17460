#!/bin/bash
# round-end style pass: GPU tests, smoke(), reference arm, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -${TAILN:-8}
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
T0=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.log | cut -c1-900; echo "reference arm took $(( $(date +%s) - T0 )) s"
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c1-600; echo "default bench took $(( $(date +%s) - T0 )) s"
