"""gn_eval_stream (fp32 records) against the numpy fp64 evaluation for one pair's matched lists, whole and by halves, to find
the features whose streamed contribution is off.   usage: python tools/probes/stream_eval_probe.py pair"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import ref_numpy as R
from stvo_pl_b200 import types as T
from stvo_pl_b200.engine import Engine

pid = int(sys.argv[1])
eng = Engine(0)
cfg = bench.workload_config()
prev, curr, Tgt, cam = bench.make_workload(1, first_pair=pid)
out = eng.track_batch(cam, cfg, prev, curr)
res = out["results"] if isinstance(out, dict) else out[0]
m12p, m12l = out["m12_pt"], out["m12_ls"]
ip, il = np.nonzero(m12p >= 0)[0], np.nonzero(m12l >= 0)[0]
s2 = prev.ls_sigma2[il].copy()
if prev.ls_level is not None:
    s2 = s2 * cfg.lsd_scale ** prev.ls_level[il]
args = dict(P=prev.pt_P[ip], obs=curr.pt_pl[m12p[ip]], s2p=prev.pt_sigma2[ip], sP=prev.ls_sP[il], eP=prev.ls_eP[il], le=curr.ls_le[m12l[il]],
            spl=prev.ls_spl[il], epl=prev.ls_epl[il], s2l=1.0 / (s2 * s2))
DT = np.linalg.inv(res["DT"][0]) if False else res["DT_opt"][0]
print("pair", pid, "np", len(ip), "nl", len(il))

def both(selp, sell, tag):
    a = {k: (v[selp] if k in ("P", "obs", "s2p") else v[sell]) for k, v in args.items()}
    pr = R.Problem(cam, cfg, a["P"], a["obs"], a["s2p"], a["sP"], a["eP"], a["le"], a["spl"], a["epl"], a["s2l"])
    H0, g0, e0 = pr.evaluate(DT)
    mb = T.MatchedBatch(pt_off=[0, len(a["P"])], ls_off=[0, len(a["sP"])], pt_P=a["P"], pt_pl_obs=a["obs"], pt_sigma2=a["s2p"], ls_sP=a["sP"],
                        ls_eP=a["eP"], ls_le_obs=a["le"], ls_spl=a["spl"], ls_epl=a["epl"], ls_sigma2=a["s2l"])
    H1, g1, e1, _ = eng.gn_eval_stream(cam, cfg, mb, DT[None], iters=1)
    n = len(a["P"]) + len(a["sP"])
    dH = np.abs(H1[0] - H0).max() / np.abs(H0).max(); dg = np.abs(g1[0] - g0).max() / (np.abs(g0).max() + 1e-300)
    print("%-28s n=%5d  rel dH %.2e  rel dg %.2e (|g| %.2e)  e %.8e vs %.8e" % (tag, n, dH, dg, np.abs(g0).max(), e1[0] / 1.0, e0 * 1.0))
    return dH, dg

allp, alll = np.arange(len(ip)), np.arange(len(il))
none_p, none_l = allp[:0], alll[:0]
both(allp, alll, "all")
both(allp, none_l, "points only")
both(none_p, alll, "lines only")
# bisect the lines, then the points
for name, full, mk in (("lines", alll, lambda s: both(none_p, s, "lines subset")), ("points", allp, lambda s: both(s, none_l, "points subset"))):
    sel = full
    while len(sel) > 1:
        h = len(sel) // 2
        da, db = mk(sel[:h]), mk(sel[h:])
        sel = sel[:h] if max(da) > max(db) else sel[h:]
    print("worst single", name, "index", sel)
