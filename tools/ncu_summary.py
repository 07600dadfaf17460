#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md>      per-kernel count / total / share / average
  python tools/ncu_summary.py raw <file.ncu-rep> <out.md> [regex]   key counters of each profiled launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    # the FP64 pipe of a sub-partition is shared by the scalar FP64 instructions and DMMA: "pipe_shared" is its busy fraction
    # (calibrated on tools/fp64_micro2.cu: 99 % for a saturated DMMA stream), "fp64" the scalar part, "dmma" the mma part
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6}.get(r[ui], 1.0)
        k = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} | {v[1] / v[0]:.1f} |\n")
        f.write(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches "
                "(ncu --metrics gpu__time_duration.sum --clock-control none; serialised, cold cache: compare shares)\n")


def raw(rep, out, pattern=None):
    if rep.endswith(".csv"):  # raw page already exported on the GPU box (tools/gpu_r2_profiles.sh)
        txt = open(rep, errors="replace").read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if pattern and not re.search(pattern, name):
                continue
            f.write(f"## {name.split('(')[0]}  (launch id {r[0]})\n\n| metric | value | unit |\n|---|---|---|\n")
            stalls = []
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"| {h} | {r[i]} | {units[i]} |\n")
                m = STALL.fullmatch(h)
                if m:
                    try:
                        stalls.append((float(r[i]), m.group(1)))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            f.write("\nwarp stall reasons per issue (top): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + "\n\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        raw(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
