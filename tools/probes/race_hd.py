"""racecheck probe: the TMA ring of the streamed evaluator with more tiles per problem than ring stages (slot reuse), through
the stand-alone sweep kernel (RACE_KERNEL=sweep) or the GN loop kernel (default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from stvo_pl_b200 import synth, types as T
from stvo_pl_b200.engine import Engine
eng = Engine(0)
mb_hd, T_hd, cam_hd = synth.make_matched_batch(os.environ.get("RACE_SHAPE", "hd"), 2)
if os.environ.get("RACE_KERNEL") == "sweep":
    eng.gn_eval_stream(cam_hd, T.kitti_config(), mb_hd, T_hd, iters=1)
else:
    cfg = T.kitti_config()
    if os.environ.get("RACE_ITERS"):
        cfg.max_iters = cfg.max_iters_ref = int(os.environ["RACE_ITERS"])
    eng.optimize_pose(cam_hd, cfg, mb_hd)
print("ok")
