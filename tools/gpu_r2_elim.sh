#!/bin/bash
# Round 2: new record layouts + elimination variants.  GPU tests with the default build, then the 128-trial step per variant.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -12
bash tools/gpu_variants.sh "$@" 2>&1 | tee gpurun_out/variants.txt
