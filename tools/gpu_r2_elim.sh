#!/bin/bash
# elimination variants: solve-related GPU tests with the default build, then the 128-trial step per variant
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_chain.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -8
bash tools/gpu_variants.sh "$@" 2>&1 | tee gpurun_out/variants.txt
