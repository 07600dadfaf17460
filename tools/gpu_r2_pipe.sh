#!/bin/bash
# software-pipelined elimination: parity tests with that build, then the 128-trial step per variant
mkdir -p gpurun_out
V2GT_LIB_CUDA=$PWD/tools/bin/lib_pipe3acc.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_chain.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pipe.log
grep -v "^$" gpurun_out/pytest_pipe.log | tail -8
bash tools/gpu_variants.sh "$@" 2>&1 | tee gpurun_out/variants_pipe.txt
