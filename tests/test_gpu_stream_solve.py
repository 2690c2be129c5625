"""The streamed form of optimizePose (evaluation sweeps from HBM + per-problem step kernel, SURVEY 8(d) / VERDICT r1 item 5)
forced onto ordinary frame sizes (PLSTVO_STREAM_SOLVE=1 is read when the library loads, hence the subprocess), so that every
branch is exercised where the oracle finishes in seconds: stage 1 -> gate -> removeOutliers -> stage 2, the robust fallback and
the robust main mode (both handed back to K2), too few features before / after, priors with the motion model, explicit lists."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, sys
import numpy as np
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import ref_numpy as R
from stvo_pl_b200 import synth, types as T
from stvo_pl_b200.engine import Engine
from oracle.oracle import Oracle
eng, orc = Engine(0), Oracle()
out = []
def run(name, shape, cfg, B, priors=None, **kw):
    prev, curr, Tgt, cam = synth.make_batch(shape, B, **kw)
    g = eng.track_batch(cam, cfg, prev, curr, priors=priors)
    r = orc.track_batch(cam, cfg, prev, curr, priors=priors)
    rec = dict(name=name, m12=bool((g["m12_pt"] == r["m12_pt"]).all() and (g["m12_ls"] == r["m12_ls"]).all()),
               flags=int((g["inlier_pt"] != r["inlier_pt"]).sum() + (g["inlier_ls"] != r["inlier_ls"]).sum()),
               n=int(prev.n_pt + prev.n_ls), status=[], ang=0.0, tr=0.0, counts=True, tfw=0.0)
    for p in range(B):
        a, b = g["results"][p], r["results"][p]
        rec["status"].append([int(a["status"]), int(b["status"]), int(a["good"]), int(b["good"])])
        ang, tr = R.pose_error(a["DT"], b["DT"])
        rec["ang"], rec["tr"] = max(rec["ang"], float(ang)), max(rec["tr"], float(tr))
        rec["tfw"] = max(rec["tfw"], float(np.abs(a["Tfw"] - b["Tfw"]).max()))
        rec["counts"] = rec["counts"] and int(a["n_matched_pt"]) == int(b["n_matched_pt"]) and int(a["n_matched_ls"]) == int(b["n_matched_ls"])
    out.append(rec)
cfg = T.kitti_config()
run("kitti", "kitti", cfg, 5)
run("kitti_small", "kitti", cfg, 4, n_pt=300, n_ls=80)
run("points_only", "kitti_points", cfg, 3)
e = T.euroc_config(); run("euroc_levels", "euroc", e, 3)
e1 = T.euroc_config(); e1.solver_mode = 1; run("robust_mode_delegated", "euroc", e1, 3)
few = T.kitti_config(); run("few_before", "kitti", few, 2, n_pt=6, n_ls=2, overlap=1.0, outlier_frac=0.0)
fa = T.kitti_config(); fa.inlier_k = 0.05; run("few_after", "kitti", fa, 2, n_pt=10, n_ls=2, overlap=1.0, outlier_frac=0.0)
mm = T.kitti_config(); mm.use_motion_model = 1
pri = T.identity_priors(3)
for p in range(3):
    pri[p]["DT"] = synth.expmap_se3(np.array([0.4, -0.3, 2.5, 0.05, -0.08, 0.03])); pri[p]["DT_cov"] = np.eye(6) * 1e-6; pri[p]["err_norm"] = 0.2
    pri[p]["Tfw"] = synth.expmap_se3(np.array([1.0, 2.0, 3.0, 0.1, 0.2, -0.1]))
run("bad_prior_robust_fallback", "kitti", mm, 3, priors=pri, n_pt=500, n_ls=120)
# explicit lists
mb, Ts, cam = synth.make_matched_batch("kitti", 3)
res, ip, il = eng.optimize_pose(cam, cfg, mb)
rc, ref, rp, rl = orc.optimize_pose(cam, cfg, mb)
rec = dict(name="explicit_lists", m12=True, flags=int((ip != rp).sum() + (il != rl).sum()), n=len(ip) + len(il), status=[], ang=0.0, tr=0.0, counts=True, tfw=0.0)
for p in range(3):
    rec["status"].append([int(res[p]["status"]), int(ref[p]["status"]), int(res[p]["good"]), int(ref[p]["good"])])
    ang, tr = R.pose_error(res[p]["DT"], ref[p]["DT"])
    rec["ang"], rec["tr"] = max(rec["ang"], float(ang)), max(rec["tr"], float(tr))
out.append(rec)
print("RESULT " + json.dumps(out))
'''


@pytest.mark.gpu
def test_streamed_solver_on_ordinary_frames():
    env = dict(os.environ, PLSTVO_STREAM_SOLVE="1")
    r = subprocess.run([sys.executable, "-c", "ROOT = %r\n" % ROOT + SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    for rec in json.loads(line[7:]):
        print(rec)
        assert rec["m12"] and rec["counts"], rec                       # matching is untouched: bit-exact
        for st_gpu, st_ref, good_gpu, good_ref in rec["status"]:
            assert st_gpu == st_ref and good_gpu == good_ref, rec      # same branch of optimizePose
        assert rec["ang"] < 1e-5 and rec["tr"] < 1e-4, rec             # north_star's tolerance
        assert rec["flags"] <= max(2, rec["n"] // 1000), rec           # borderline residuals only
