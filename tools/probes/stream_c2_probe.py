"""C2 (or any workload) through K2 and through the streamed solver: stage times of both and the difference of the results.
usage: python tools/probes/stream_c2_probe.py [pairs] [workload]   (runs itself once per PLSTVO_STREAM_SOLVE mode)"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

def child(mode, B, wl, path):
    import bench
    from stvo_pl_b200.engine import Engine
    bench.ACTIVE = wl
    eng = Engine(0)
    cfg = bench.workload_config()
    prev, curr, Tgt, cam = bench.make_workload(B, first_pair=0)
    db = eng.upload(cam, cfg, prev, curr)
    for _ in range(3):
        db.run()
    eng.synchronize()
    ms = db.run_timed(20, flush_l2=False)
    st = db.stage_times(iters=10)
    out = db.download()
    print(f"mode {mode}: {ms:.4f} ms/step  " + "  ".join(f"{k}={v:.4f}" for k, v in st.items() if k.startswith("ms_")), "streamed", st["streamed_solver"], "delegated", st["delegated_to_fp64"], flush=True)
    np.savez(path, off_pt=prev.pt_off, off_ls=prev.ls_off, T=out["results"]["DT"], good=out["results"]["good"], err=out["results"]["err_norm"], eig=out["results"]["DT_cov_eig"], ninl=out["results"]["n_inliers"], status=out["results"]["status"], inl_pt=out["inlier_pt"],
             inl_ls=out["inlier_ls"], m12=out["m12_pt"], it1=out["results"]["iters_stage1"], it2=out["results"]["iters_stage2"], Tgt=Tgt)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5])
        sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    wl = sys.argv[2] if len(sys.argv) > 2 else "c2"
    res = {}
    for mode in ("0", "1"):
        path = f"/tmp/stream_probe_{mode}.npz"
        env = dict(os.environ, PLSTVO_STREAM_SOLVE=mode)
        subprocess.run([sys.executable, __file__, "--child", mode, str(B), wl, path], env=env, check=True)
        res[mode] = np.load(path)
    a, b = res["0"], res["1"]
    sys.path.insert(0, os.path.join(ROOT, "tests")); from ref_numpy import pose_error
    rot = np.zeros(B); tr = np.zeros(B)
    for i in range(B):
        rot[i], tr[i] = pose_error(a["T"][i], b["T"][i])
    print("streamed vs K2: max rot %.3e rad, max trans %.3e m; status diffs %d, good diffs %d; inlier flag diffs pt %d / %d, ls %d / %d; match diffs %d"
          % (rot.max(), tr.max(), int((a["status"] != b["status"]).sum()), int((a["good"] != b["good"]).sum()),
             int((a["inl_pt"] != b["inl_pt"]).sum()), a["inl_pt"].size, int((a["inl_ls"] != b["inl_ls"]).sum()), a["inl_ls"].size,
             int((a["m12"] != b["m12"]).sum())))
    print("iterations: K2 stage1 mean %.2f stage2 %.2f; streamed %.2f / %.2f" % (a["it1"].mean(), a["it2"].mean(), b["it1"].mean(), b["it2"].mean()))
    op, ol = a["off_pt"], a["off_ls"]
    fd = np.array([int((a["inl_pt"][op[i]:op[i+1]] != b["inl_pt"][op[i]:op[i+1]]).sum() + (a["inl_ls"][ol[i]:ol[i+1]] != b["inl_ls"][ol[i]:ol[i+1]]).sum()) for i in range(B)])
    itd = (a["it1"] != b["it1"]) | (a["it2"] != b["it2"])
    same = (fd == 0) & ~itd
    print("pairs with flag diffs %d, with iteration-count diffs %d (stage1 %d, stage2 %d); both equal %d" % (int((fd > 0).sum()), int(itd.sum()), int((a["it1"] != b["it1"]).sum()), int((a["it2"] != b["it2"]).sum()), int(same.sum())))
    for name, m in (("same flags+iters", same), ("flag diffs", fd > 0), ("iter diffs only", itd & (fd == 0))):
        if m.any():
            print("  %-18s rot max %.3e median %.3e | trans max %.3e median %.3e" % (name, rot[m].max(), np.median(rot[m]), tr[m].max(), np.median(tr[m])))
    # deviation from the ground-truth motion, for scale
    rg = np.array([pose_error(a["T"][i], np.linalg.inv(a["Tgt"][i]) if False else a["Tgt"][i])[0] for i in range(B)])
    print("K2 vs ground truth rot median %.3e (inverse convention not checked)" % np.median(rg))
    order = np.argsort(-tr)[:8]
    for i in order:
        print("  pair %3d: rot %.3e trans %.3e  it %d/%d vs %d/%d  flags diff %d  err %.6e vs %.6e  inliers %d  eig max %.3e min %.3e" % (
            i, rot[i], tr[i], a["it1"][i], a["it2"][i], b["it1"][i], b["it2"][i], fd[i], a["err"][i], b["err"][i], a["ninl"][i], a["eig"][i].max(), a["eig"][i].min()))
    print("percentiles of trans dev: 50%% %.2e 90%% %.2e 99%% %.2e; rot: 50%% %.2e 90%% %.2e 99%% %.2e" % (*np.percentile(tr, [50, 90, 99]), *np.percentile(rot, [50, 90, 99])))
