"""Probe: pure H2D time of one C2 batch from pinned memory (how far is the e2e path from the PCIe bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stvo_pl_b200 import synth, types as T
from stvo_pl_b200.engine import Engine
eng = Engine(0)
prev, curr, _, cam = synth.make_batch("kitti", 512, overlap=1.0)
pp, pc = eng.pinned.pin_frames(prev), eng.pinned.pin_frames(curr)
cfg = T.kitti_config()
for k in range(4):
    t0 = time.perf_counter(); db = eng.upload(cam, cfg, pp, pc); dt = time.perf_counter() - t0
    nbytes = prev.input_bytes("prev") + curr.input_bytes("curr")
    print(f"upload {k}: {dt*1e3:.3f} ms  {nbytes/dt/1e9:.1f} GB/s"); db.free()
