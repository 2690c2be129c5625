"""Probe: pure H2D time of one C2 batch from pinned memory, with the process bound to each NUMA node in turn
(how far is the e2e path from the PCIe bound, and does host-memory placement matter on this box?)."""
import glob, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stvo_pl_b200 import synth, types as T
from stvo_pl_b200.engine import Engine


def cpus_of(node):
    out = set()
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("numa nodes", nodes, "gpu numa:", [open(p).read().strip() for p in glob.glob("/sys/bus/pci/devices/*/numa_node")
                                          if os.path.exists(os.path.dirname(p) + "/vendor") and
                                          open(os.path.dirname(p) + "/vendor").read().strip() == "0x10de"][:8])
prev, curr, _, cam = synth.make_batch("kitti", 512, overlap=1.0)
cfg = T.kitti_config()
allcpus = os.sched_getaffinity(0)
for node in nodes + [None]:
    os.sched_setaffinity(0, cpus_of(node) & allcpus if node is not None else allcpus)
    eng = Engine(0)
    pp, pc = eng.pinned.pin_frames(prev), eng.pinned.pin_frames(curr)
    best = 1e9
    for k in range(4):
        t0 = time.perf_counter(); db = eng.upload(cam, cfg, pp, pc); best = min(best, time.perf_counter() - t0); db.free()
    nbytes = prev.input_bytes("prev") + curr.input_bytes("curr")
    print(f"bound to node {node}: upload {best*1e3:.3f} ms  {nbytes/best/1e9:.1f} GB/s")
    eng.close()
