#!/usr/bin/env python
"""Turns the ncu outputs a gpurun call brought back (gpurun_out/) into the small tracked summaries under
profiles/: launch list (per-kernel totals and shares), key raw metrics of the full captures, bench JSON lines.

usage: python tools/summarise_profiles.py r01
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return {}
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h == "Kernel Name":
            out[h] = f"{v} {u}".strip()
    return out


def launch_list(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        t = float(r[mv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu].strip(), 1e-6)
        name = r[kn].split("(")[0]
        tot[name] += t * scale
        cnt[name] += 1
    return tot, cnt


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    lines = [f"# ncu summary {tag}", ""]
    ll = os.path.join(OUT, f"launches_{tag}.csv")
    if os.path.exists(ll):
        tot, cnt = launch_list(ll)
        total = sum(tot.values())
        lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised:",
                  "compare SHARES, not absolutes) of `python bench.py --steps 2 --warmup 1 --no-cpu`", "",
                  "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for k in sorted(tot, key=lambda k: -tot[k]):
            lines.append(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.1f}% |")
        lines.append("")
        with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
            f.write(open(ll).read())
    for kname in ("k1", "k2", "gn"):
        rep = os.path.join(OUT, f"prof_{kname}_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        m = raw_metrics(rep)
        lines += [f"## `ncu --set full --clock-control none --import-source on` — {kname.upper()}: {m.get('Kernel Name', '')}", "",
                  "| metric | value |", "|---|---|"]
        for k in KEEP:
            if k in m:
                lines.append(f"| `{k}` | {m[k]} |")
        lines.append("")
        def num(key):
            v, u = m[key].split()[0], m[key].split()[1]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        if kname in ("k1", "gn") and "dram__bytes_read.sum" in m:
            traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            rec = {"dram_bytes_per_launch": traffic, "dram_bytes_read": num("dram__bytes_read.sum"),
                   "dram_bytes_write": num("dram__bytes_write.sum"), "capture": f"profiles/{tag}_ncu_summary.md",
                   "launch": m.get("launch__grid_size", ""), "gpu_time": m.get("gpu__time_duration.sum", "")}
            def pct(key):
                return float(m[key].split()[0]) if key in m else None
            if kname == "k1":   # the pipes that actually bound the matcher (integer ALU and XU / POPC)
                rec["alu_pipe_pct"] = pct("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed")
                rec["xu_pipe_pct"] = pct("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed")
                rec["issue_active_pct"] = pct("smsp__issue_active.avg.pct_of_peak_sustained_active")
            if kname == "gn":
                rec["problems"] = 1024          # bench.py --workload c5 default
                rec["algorithmic_bytes_per_launch"] = 1024 * (8000 * 32 + 2000 * 64)
            json.dump(rec, open(os.path.join(PROF, f"{kname}_traffic.json"), "w"), indent=1)
    names = [f"bench_{tag}.json", f"bench_ref_{tag}.json", f"bench_c5_{tag}.json"] + \
            [f"bench_{w}_{tag}.json" for w in ("c1", "c3")] + [f"bench_ref_{w}_{tag}.json" for w in ("c1", "c3")] + \
            [f"bench_{w}_{tag}.json" for w in ("stereo", "stereo_track")]
    for name in names:
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            txt = open(p).read().strip()
            open(os.path.join(PROF, name), "w").write(txt + "\n")
            lines += [f"## {name}", "", "```json", txt, "```", ""]
    open(os.path.join(PROF, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines))
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
