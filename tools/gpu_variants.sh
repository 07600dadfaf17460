#!/bin/bash
# compare alternative builds of the CUDA library (tools/bin/lib_*.so) on the 128-trial bench
B="python bench.py --trials 128 --steps 1 --warmup 1 --no-latency --no-cpu-baseline --no-sim --no-chain --no-pipeline --waves-trials 0"
P='import sys,json; d=json.loads(sys.stdin.read()); p=d["phases_ms_per_step"]; sb=d["per_rank"]["state_blocks_per_step"][0]/1e6; print("%.1f traj/s  build %.0f lin %.0f elim %.0f red %.0f back %.0f cost %.0f total %.0f | %.1f M state-blocks, elim %.3f backsub %.3f ns/state-block" % (d["value"], p["build"], p["linearize"], p["eliminate"], p["reduced"], p["backsub"], p["cost"], p["optimize_total"], sb, p["eliminate"]*1e3/sb, p["backsub"]*1e3/sb))'
echo "== default"; $B 2>&1 | tail -1 | python -c "$P"
for L in "$@"; do echo "== $L"; V2GT_LIB_CUDA=$PWD/tools/bin/$L $B 2>&1 | tail -1 | python -c "$P"; done
