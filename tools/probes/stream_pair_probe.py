"""One pair (or a few) through K2 and the streamed solver in the same process order: is a deviation a property of the pair?
usage: python tools/probes/stream_pair_probe.py first_pair count"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def child(first, B, path):
    import bench
    from stvo_pl_b200.engine import Engine
    eng = Engine(0)
    cfg = bench.workload_config()
    if os.environ.get("PROBE_ITERS_REF"): cfg.max_iters_ref = int(os.environ["PROBE_ITERS_REF"])
    if os.environ.get("PROBE_ITERS"): cfg.max_iters = int(os.environ["PROBE_ITERS"])
    prev, curr, Tgt, cam = bench.make_workload(B, first_pair=first)
    db = eng.upload(cam, cfg, prev, curr)
    db.run(); eng.synchronize()
    out = db.download()
    np.savez(path, T=out["results"]["DT"], it1=out["results"]["iters_stage1"], it2=out["results"]["iters_stage2"], err=out["results"]["err_norm"],
             inl_pt=out["inlier_pt"], inl_ls=out["inlier_ls"])

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]); sys.exit(0)
    first, B = int(sys.argv[1]), int(sys.argv[2])
    res = {}
    for mode in ("0", "1"):
        path = f"/tmp/pair_probe_{mode}.npz"
        subprocess.run([sys.executable, __file__, "--child", str(first), str(B), path], env=dict(os.environ, PLSTVO_STREAM_SOLVE=mode), check=True)
        res[mode] = np.load(path)
    from ref_numpy import pose_error
    a, b = res["0"], res["1"]
    for i in range(B):
        rot, tr = pose_error(a["T"][i], b["T"][i])
        print("pair %d: rot %.3e trans %.3e it %d/%d vs %d/%d err %.7e %.7e" % (first + i, rot, tr, a["it1"][i], a["it2"][i], b["it1"][i], b["it2"][i], a["err"][i], b["err"][i]))
    print("flag diffs", int((a["inl_pt"] != b["inl_pt"]).sum()), int((a["inl_ls"] != b["inl_ls"]).sum()))
