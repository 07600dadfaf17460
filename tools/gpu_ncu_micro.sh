#!/bin/bash
# calibrate the ncu FP64 pipe counters on the DMMA microbenchmark, then capture k_seg_eliminate on a 16-trial batch
mkdir -p gpurun_out
ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,smsp__cycles_active.avg,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/ncu_micro.csv ./tools/bin/fp64_micro2 > gpurun_out/ncu_micro.log 2>&1
tail -3 gpurun_out/ncu_micro.log
bash tools/gpu_prof_one.sh k_seg_eliminate r02a_elim 3 --no-sim --no-chain
