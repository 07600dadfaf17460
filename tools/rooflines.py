#!/usr/bin/env python
"""profiles/<tag>_rooflines.md: one table over every kernel of the hot path, from the ncu --set full captures of
tools/gpu_r2_profiles.sh (gpurun_out/prof_<tag>_<kernel>.ncu-rep): algorithmic bytes and flops per unit,
measured DRAM bytes, achieved GB/s and FP64 TFLOP/s, fractions of the measured HBM peak (MEASURED_PEAKS.json) and of the
FP64 pipe.

  python tools/rooflines.py <tag> <states_in_the_profiled_batch> [fp64_peak_tflops]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel -> (unit, algorithmic bytes per unit, algorithmic flops per unit, bound, note)   [SURVEY.md 8(d), DESIGN.md 4]
ALG = {
    "k_init_states": ("raw state", 3 * 230 + 128, 3 * 1500, "latency", "three interpolations per state (t, t +- 0.25 s)"),
    "k_compact": ("raw state", 2 * 136, 0, "latency", "order-preserving compaction of times + state rows"),
    "k_preintegrate": ("interval", 5 * 56 + 128 + 512, 15e3, "latency", "U1 means and bias Jacobians x ~5 IMU steps, one thread per interval; writes the 64-double constants row"),
    "k_preint_cov": ("interval", 5 * 56 + 128 + 960, 45e3, "fp64", "U1 covariance (RK4, ~5 IMU steps) + U1' information matrix, sixteen lanes per interval; writes the packed 120-double information matrix"),
    "k_factor_eval<1>": ("state", 1900 + 1216, 8e3, "hbm", "U3 without the whitening products: reads state / constants, writes the factor pieces"),
    "k_factor_eval<0>": ("state", 1900 + 128 + 384, 3e3, "hbm", "U5: trial point (retract) + both residuals"),
    "k_assemble": ("state", 960 + 456 + 704 + 512 + 3968 + 128, 17e3, "hbm", "J^T Lambda J from the 3x3 block structure; writes the 496-double record"),
    "k_reduce_gamma": ("trial", 0, 0, "latency", "fixed-order sum of the per-CTA partials of the 9x9 global block"),
    "k_seg_eliminate": ("state block", 7920, 19.4e3, "hbm+fp64", "U4: reads D,O,B,g once, writes L^-1,W once (Thomas-equivalent counts)"),
    "k_reduced_assemble": ("separator", 3 * 3968, 600, "latency", "separator record minus the Schur terms of its two segments"),
    "k_reduced": ("separator", 7920, 19.4e3, "latency", "the same elimination over the (level-2) separator chain, one warp per trial"),
    "k_l2_prepare": ("trial", 0, 0, "latency", "Schur terms of the level-1 segments on the global block"),
    "k_sep_delta": ("separator", 256, 0, "latency", "separator steps into the state-step array"),
    "k_seg_backsub": ("state block", 3960 + 128, 1500, "hbm", "reads the factor record once, writes the step"),
    "k_lm_decide": ("trial", 0, 0, "latency", "accept / reject, lambda update"),
}


def raw_rows(rep):
    """rows of the raw counter page: from the CSV written on the GPU box (raw_<tag>_<kernel>.csv) or from the report itself"""
    if rep.endswith(".csv"):
        txt = open(rep, errors="replace").read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        for h, u in zip(hdr, units):  # durations -> microseconds, byte counters -> bytes
            if u in scale and h in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
                try:
                    d[h] = str(float(d[h].replace(",", "")) * scale[u])
                except ValueError:
                    pass
        out.append(d)
    return out


def f(d, k, default=0.0):
    try:
        return float(d.get(k, default))
    except (TypeError, ValueError):
        return default


def main():
    tag, states = sys.argv[1], float(sys.argv[2])
    fp64_peak = float(sys.argv[3]) if len(sys.argv) > 3 else 34.2
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6551.4))
    out = ["# Per-kernel rooflines at HEAD (%s)" % tag, "",
           "ncu --set full --clock-control none, 16 full-size sim.launch trials (%.0f states), one representative launch per kernel "
           "(the largest of the captured ones).  HBM peak %.1f GB/s (MEASURED_PEAKS.json), FP64 peak %.1f TFLOP/s (DFMA loop measured by "
           "bench.py on the same GPU type).  `pipe` = busy fraction of the FP64 pipe of the SM sub-partitions, which scalar FP64 "
           "instructions and DMMA share (sm__pipe_shared_cycles_active; 99 %% for a saturated DMMA stream, tools/fp64_micro2.cu); "
           "`executed` counts what the kernel issued (DFMA = 2 flop, DADD / DMUL = 1, DMMA m8n8k4 = 512)." % (states, hbm, fp64_peak), "",
           "| kernel | unit | units / launch | time (us) | DRAM read + written (MB) | DRAM B / unit | algorithmic B / unit | achieved GB/s (DRAM) | "
           "frac of HBM peak (DRAM) | frac (algorithmic bytes) | executed FP64 (TFLOP/s) | frac of FP64 peak | algorithmic flops / unit | "
           "FP64 pipe busy | warps active | regs | top stalls |", "|" + "---|" * 17]
    notes = []
    for kern, (unit, ab, af, bound, note) in ALG.items():
        base = kern.split("<")[0]
        rep = os.path.join(ROOT, "gpurun_out", "raw_%s_%s.csv" % (tag, base))
        if not os.path.exists(rep):
            rep = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.ncu-rep" % (tag, base))
        if not os.path.exists(rep):
            continue
        rows = raw_rows(rep)
        want = None
        if "<" in kern:
            inst = kern.split("<")[1].rstrip(">")
            rows = [r for r in rows if "<%s>" % inst in r.get("Kernel Name", "") or "<(bool)%s>" % inst in r.get("Kernel Name", "")]
        if not rows:
            continue
        want = max(rows, key=lambda r: f(r, "gpu__time_duration.sum"))
        t_us = f(want, "gpu__time_duration.sum")
        rd, wr = f(want, "dram__bytes_read.sum"), f(want, "dram__bytes_write.sum")
        units = {"raw state": states * 1.0035, "state": states, "interval": states, "state block": states,
                 "separator": 16 * 590.0, "trial": 16.0}[unit]
        # thread-level instruction counts: (sum over sub-partitions of instructions per elapsed cycle) x elapsed cycles; DMMA
        # warp instructions from the busy cycles of the dmma sub-pipe (16 cycles each, checked against the SASS count)
        cyc = f(want, "sm__cycles_elapsed.max")
        dfma = f(want, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cyc
        dmul = f(want, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cyc
        dadd = f(want, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cyc
        n_smsp = 4.0 * f(want, "device__attribute_multiprocessor_count", 148.0)
        dmma = f(want, "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg") * n_smsp / 16.0
        flops = 2 * dfma + dmul + dadd + 512.0 * dmma
        tf = flops / (t_us * 1e-6) / 1e12 if t_us > 0 else 0.0
        gbs = (rd + wr) / (t_us * 1e-6) / 1e9 if t_us > 0 else 0.0
        alg_gbs = ab * units / (t_us * 1e-6) / 1e9 if t_us > 0 else 0.0
        stalls = sorted(((f(want, k), k.split("issue_stalled_")[1].split("_per")[0]) for k in want
                         if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)[:3]
        out.append("| `%s` | %s | %.3g | %.1f | %.1f | %.0f | %s | %.0f | %.2f | %s | %.2f | %.2f | %s | %.0f %% | %.0f %% | %d | %s |" % (
            kern, unit, units, t_us, (rd + wr) / 1e6, (rd + wr) / units, ("%.0f" % ab) if ab else "-", gbs, gbs / hbm,
            ("%.2f" % (alg_gbs / hbm)) if ab else "-", tf, tf / fp64_peak, ("%.3g" % af) if af else "-",
            f(want, "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"),
            f(want, "sm__warps_active.avg.pct_of_peak_sustained_active"), int(f(want, "launch__registers_per_thread")),
            ", ".join("%s %.1f" % (n, v) for v, n in stalls)))
        notes.append("* `%s` (%s-bound): %s." % (kern, bound, note))
    out += ["", "What the algorithmic figures count:", ""] + notes
    path = os.path.join(ROOT, "profiles", "%s_rooflines.md" % tag)
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
