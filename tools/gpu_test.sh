#!/bin/bash
# GPU pass: parity tests + a short bench.  usage: tools/gpu_test.sh [bench args...]
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -${TAILN:-25}
if [ -n "$1" ]; then timeout 900 python bench.py "$@" > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log; fi
