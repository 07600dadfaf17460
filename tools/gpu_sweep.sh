#!/bin/bash
B="python bench.py --trials 128 --steps 1 --warmup 1 --no-latency --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read()); p=d["phases_ms_per_step"]; print("%.1f traj/s  elim %.0f red %.0f back %.0f total %.0f" % (d["value"], p["eliminate"], p["reduced"], p["backsub"], p["optimize_total"]))'
for tw in 3552 5328 7104 10656 14208; do echo "== target_warps=$tw"; $B --target-warps $tw 2>&1 | tail -1 | python -c "$P"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
