#!/usr/bin/env python
"""Turns the ncu outputs a gpurun call brought back (gpurun_out/) into the small tracked summaries under profiles/:
launch lists (per-kernel totals and shares, and the kernel sequence of one step), key raw metrics of every full capture, the
top source lines of each capture by stall samples, the traffic / pipe JSONs bench.py reads, and the bench JSON lines.

usage: python tools/summarise_profiles.py r2b
(tools/gpu_profile.sh runs it on the GPU box with PROF_OUT=gpurun_out/profiles: the .ncu-rep files of a full pass exceed what
gpurun copies back, so only the summaries and two of the reports travel; copy gpurun_out/profiles/* to profiles/ afterwards)
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("PROF_OUT") or os.path.join(ROOT, "profiles")   # PROF_OUT: summarise on the GPU box into gpurun_out/

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]
# capture name -> (title, what the bench calls the kernel)
CAPTURES = [
    ("tc_hamming", "K1b tc_hamming_kernel (C2)"), ("tc_expand", "K1a tc_expand_kernel (C2)"), ("tc_resolve", "K1c tc_resolve_kernel (C2)"),
    ("stream_prepare", "S1 stream_prepare_kernel (C2)"), ("gn_loop", "S2 gn_loop_stream_kernel, a stage-2 launch (C2: records resident in smem)"),
    ("stream_outlier", "S3 stream_outlier_kernel (C2)"), ("stream_finalize", "S4 stream_finalize_kernel (C2)"),
    ("track_solve", "K2 track_solve_kernel (C2, PLSTVO_STREAM_SOLVE=0)"),
    ("gn_loop_c5", "S2 gn_loop_stream_kernel, a stage-2 launch (C5: records streamed from HBM)"),
    ("stream_outlier_c5", "S3 stream_outlier_kernel (C5)"), ("stream_prepare_c5", "S1 stream_prepare_kernel (C5)"),
    ("tc_hamming_c5", "K1b tc_hamming_kernel (C5)"),
    ("gn_eval", "stand-alone sweep kernel gn_eval_stream_kernel (--workload c5_sweep: 1024 C5 problems per sweep)"),
]


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return {}
    out = {}
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        out[h] = (v, u)
    return out


def stall_lines(rep, top=8):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    per, cur, tot = collections.OrderedDict(), None, 0
    for r in csv.reader(txt.splitlines()):
        if len(r) < 7:
            continue
        if r[0].isdigit():
            cur = (int(r[0]), r[1].strip()[:100])
            continue
        if r[2].startswith("0x") and cur is not None:
            try:
                n = int(r[4])
            except ValueError:
                continue
            per[cur] = per.get(cur, 0) + n
            tot += n
    return tot, sorted(per.items(), key=lambda kv: -kv[1])[:top]


def num(m, key):
    v, u = m[key]
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3, "usecond": 1e-3,
                "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(u, 1)


def launch_seq(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        t = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu].strip(), 1e-6)
        name = r[kn].split("(")[0].replace("void ", "").replace("plstvo::", "").replace("<unnamed>::", "")
        seq.append((name, t))
    return seq


def launch_section(lines, path, title, step_marker):
    seq = launch_seq(path)
    tot, cnt = collections.defaultdict(float), collections.defaultdict(int)
    for k, t in seq:
        tot[k] += t
        cnt[k] += 1
    total = sum(tot.values())
    lines += [f"## Launch list: {title}", "", "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes)", "",
              "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k in sorted(tot, key=lambda k: -tot[k]):
        lines.append(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.1f}% |")
    idx = [i for i, (k, _) in enumerate(seq) if k.startswith(step_marker)]
    if len(idx) >= 2:
        a, b = idx[-2], idx[-1]
        lines += ["", "Kernel sequence of one pass over one chunk of the batch (the last complete one in the capture):", "",
                  "| # | kernel | us |", "|---|---|---|"]
        for j, (k, t) in enumerate(seq[a:b]):
            lines.append(f"| {j} | `{k}` | {t * 1e3:.1f} |")
    lines.append("")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2b"
    os.makedirs(PROF, exist_ok=True)
    lines = [f"# ncu summary {tag}", ""]
    for suffix, title, marker in (("", "`python bench.py --steps 2 --warmup 1 --no-cpu --no-hbm-run` (C2, 512 pairs)", "tc_expand"),
                                  ("_c5", "`python bench.py --workload c5 --steps 3 --warmup 2` (C5, 512 pairs)", "tc_expand"),
                                  ("_stereo_track", "`python bench.py --workload stereo_track --steps 1 --warmup 3 --no-cpu`", "mg_")):
        ll = os.path.join(OUT, f"launches{suffix}_{tag}.csv")
        if os.path.exists(ll):
            launch_section(lines, ll, title, marker)
            with open(os.path.join(PROF, f"{tag}_launches{suffix}.csv"), "w") as f:
                f.write(open(ll).read())
    traffic = {}
    for name, title in CAPTURES:
        rep = os.path.join(OUT, f"prof_{name}_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        m = raw_metrics(rep)
        if not m:
            continue
        lines += [f"## `ncu --set full --clock-control none --import-source on` — {title}", "", f"`{m.get('Kernel Name', ('', ''))[0][:150]}`", "",
                  "| metric | value |", "|---|---|"]
        for k in KEEP:
            if k in m:
                lines.append(f"| `{k}` | {m[k][0]} {m[k][1]} |")
        stalls = sorted(((h, float(v.replace(',', ''))) for h, (v, u) in m.items()
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h),
                        key=lambda kv: -kv[1])[:5]
        if stalls:
            lines.append("| top stall reasons (warps per issue) | " + ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for h, v in stalls) + " |")
        lines.append("")
        tot, top = stall_lines(rep)
        if tot:
            lines += ["Top source lines by warp-stall samples:", "", "| share | line | source |", "|---|---|---|"]
            for (ln, src), n in top:
                lines.append(f"| {100.0 * n / tot:.1f}% | {ln} | `{src.replace('|', '/')}` |")
            lines.append("")
        rec = {"dram_bytes_per_launch": num(m, "dram__bytes_read.sum") + num(m, "dram__bytes_write.sum"),
               "dram_bytes_read": num(m, "dram__bytes_read.sum"), "dram_bytes_write": num(m, "dram__bytes_write.sum"),
               "gpu_time_ms": num(m, "gpu__time_duration.sum"), "grid": m["launch__grid_size"][0], "capture": f"profiles/{tag}_ncu_summary.md"}
        for key, short in (("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu_pipe_pct"),
                           ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma_pipe_pct"),
                           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
                           ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64_pipe_pct"),
                           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                           ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct")):
            if key in m:
                rec[short] = float(m[key][0].replace(",", ""))
        traffic[name] = rec
    if traffic:
        # bench.py: roofline.traffic of the dominant kernel by its name; roofline.pipes = the matcher's pipe utilisation
        k = {}
        names = {"tc_hamming": "tc_hamming_kernel", "tc_expand": "tc_expand_kernel", "tc_resolve": "tc_resolve_kernel",
                 "stream_prepare": "stream_prepare_kernel", "gn_loop": "gn_loop_stream_kernel", "stream_outlier": "stream_outlier_kernel",
                 "track_solve": "track_solve_kernel"}
        for cap, kn in names.items():
            if cap in traffic:
                k[kn] = traffic[cap]
        if "tc_hamming" in traffic:
            t = traffic["tc_hamming"]
            k["pipes"] = {"kernel": "tc_hamming_kernel", "alu_pipe_pct": t.get("alu_pipe_pct"), "fma_pipe_pct": t.get("fma_pipe_pct"),
                          "tensor_pipe_pct": t.get("tensor_pipe_pct"), "issue_active_pct": t.get("issue_active_pct")}
        json.dump(k, open(os.path.join(PROF, "k_traffic.json"), "w"), indent=1)
        if "gn_eval" in traffic:   # bench.py c5_sweeps: roofline_hbm_kernel.traffic
            g = dict(traffic["gn_eval"])
            g.update(problems=1024, algorithmic_bytes_per_launch=1024 * (8000 * 32 + 2000 * 64))
            json.dump(g, open(os.path.join(PROF, "gn_traffic.json"), "w"), indent=1)
        if "gn_loop_c5" in traffic:
            c5 = dict(traffic["gn_loop_c5"])
            c5["note"] = "one stage-2 gn_loop_stream_kernel launch over a 256-pair chunk of the C5 batch (8000 + 2000 records per problem)"
            json.dump(c5, open(os.path.join(PROF, "c5_traffic.json"), "w"), indent=1)
    names = [f"bench_{tag}.json", f"bench_ref_{tag}.json", f"bench_c5_{tag}.json", f"bench_k2_{tag}.json"] + \
            [f"bench_{w}_{tag}.json" for w in ("c1", "c3")] + [f"bench_ref_{w}_{tag}.json" for w in ("c1", "c3")] + \
            [f"bench_{w}_{tag}.json" for w in ("stereo", "stereo_track")] + [f"bench_n{n}_{tag}.json" for n in (2, 4, 8)] + \
            [f"bench_ref_n{n}_{tag}.json" for n in (2, 4, 8)]
    for name in names:
        p = os.path.join(OUT, name)
        if os.path.exists(p) and os.path.getsize(p) > 10:
            txt = open(p).read().strip().splitlines()[-1]
            open(os.path.join(PROF, name), "w").write(txt + "\n")
            lines += [f"## {name}", "", "```json", txt, "```", ""]
    open(os.path.join(PROF, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines))
    print("\n".join(lines[:80]))


if __name__ == "__main__":
    main()
