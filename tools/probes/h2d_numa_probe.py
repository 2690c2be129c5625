"""Probe: pinned host->device copy bandwidth of cuda:0 with the allocating thread bound to each NUMA node in turn
(does host-memory placement bound the e2e figure on this box?).  Plain torch plumbing, no product code."""
import glob
import os
import subprocess

import torch


def cpus_of(node):
    out = set()
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
allcpus = os.sched_getaffinity(0)
print("nodes", nodes, "cpus allowed", len(allcpus))
N = 256 << 20
dst = torch.empty(N, dtype=torch.uint8, device="cuda:0")
dst2 = torch.empty(N, dtype=torch.uint8, device="cuda:0")
s2 = torch.cuda.Stream()
for node in nodes + [None]:
    cp = (cpus_of(node) & allcpus) if node is not None else allcpus
    if not cp:
        print("node", node, "no allowed cpus")
        continue
    os.sched_setaffinity(0, cp)
    src = torch.empty(N, dtype=torch.uint8).pin_memory()
    src.fill_(1)
    src2 = torch.empty(N, dtype=torch.uint8).pin_memory()
    src2.fill_(2)
    for mode in ("h2d", "d2h", "h2d x2 streams"):
        best = 1e9
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            if mode == "h2d":
                dst.copy_(src, non_blocking=True)
            elif mode == "d2h":
                src.copy_(dst, non_blocking=True)
            else:
                dst.copy_(src, non_blocking=True)
                with torch.cuda.stream(s2):
                    dst2.copy_(src2, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        nb = N * (2 if "x2" in mode else 1)
        print(f"node {node}: {mode:15s} {nb / best / 1e6:.1f} GB/s")
    del src, src2
