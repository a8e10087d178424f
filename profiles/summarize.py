"""Summarise ncu captures brought back in gpurun_out/ into tracked files under profiles/.

    python profiles/summarize.py <tag>     # e.g. r1b

Inputs (made on the GPU box by profiles/capture.sh):
  gpurun_out/launches_<tag>.csv                ncu --metrics gpu__time_duration.sum launch list
  gpurun_out/prof_<kernel>_<tag>.ncu-rep       ncu --set full captures of the hot kernels
Outputs: profiles/<tag>_launches.csv (copy), profiles/<tag>_summary.md, profiles/traffic.json
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
]


def launches(tag):
    src = os.path.join(GO, f"launches_{tag}.csv")
    shutil.copy(src, os.path.join(OUT, f"{tag}_launches.csv"))
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ddp::", "").replace("ddp::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)   # -> us
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for k, v in agg.items() if not k.startswith("peak_"))
    lines = ["| kernel | launches | total ms | share of step | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = "(microbenchmark)" if k.startswith("peak_") else f"{100 * v[1] / tot:.1f} %"
        lines.append(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {share} | {v[1] / v[0]:.1f} |")
    return lines


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    S = hdr.index("# Samples")
    cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[S]) for r in data) or 1
    agg = collections.Counter()
    for r in data:
        for i, h in cols:
            agg[h] += int(r[i])
    return ", ".join(f"{h[6:]} {100 * v / tot:.0f}%" for h, v in agg.most_common(5))


def main():
    tag = sys.argv[1]
    md = [f"# ncu summary `{tag}` (bench.py C4: quadruped n=36 m=12 N=200 B=1024, one B200)", "",
          "Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` over "
          "`python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras` (cold-cache, serialised: compare shares).", ""]
    md += launches(tag) + [""]
    traffic = {}
    names = {"backward_sym": "backward_sym_kernel", "quad_fused": "quad_fused_kernel",
             "rollout_quad8": "rollout_quad8_kernel", "linearize": "linearize_kernel",
             "rollout_arm8": "rollout_arm8_kernel", "quad_quat_fused": "quad_quat_fused_kernel",
             "backward_sym_n37": "backward_sym_kernel<QuadrupedQuat>"}
    for kern in ("backward_sym", "quad_fused", "linearize", "rollout_quad8", "rollout_arm8", "quad_quat_fused",
                 "backward_sym_n37"):
        rep = os.path.join(GO, f"prof_{kern}_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        m = raw(rep)
        md += [f"## `{kern}` kernel (`ncu --set full --clock-control none --import-source on`, one launch)", "",
               "| metric | value |", "|---|---|"]
        for k in METRICS:
            if k in m:
                md.append(f"| `{k}` | {m[k][0]} {m[k][1]} |")
        md += ["", "Top stall reasons (warp samples): " + stalls(rep), ""]

        def gb(key):
            v, u = m[key]
            v = float(v.replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
        traffic[names[kern]] = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    open(os.path.join(OUT, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
