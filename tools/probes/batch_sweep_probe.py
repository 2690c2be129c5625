"""ms per resident step against batch size, K2 and streamed solver (PLSTVO_STREAM_SOLVE=0 / 1), workload c2.
usage: python tools/probes/batch_sweep_probe.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

def child():
    import bench
    from stvo_pl_b200.engine import Engine
    eng = Engine(0)
    cfg = bench.workload_config()
    out = []
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 148, 256, 296, 512, 1024):
        prev, curr, Tgt, cam = bench.make_workload(B, first_pair=0)
        db = eng.upload(cam, cfg, prev, curr)
        for _ in range(3):
            db.run()
        eng.synchronize()
        ms = db.run_timed(20) / 20
        st = db.stage_times(iters=5)
        out.append((B, ms, st["ms_solve"]))
        db.free()
    print(os.environ.get("PLSTVO_STREAM_SOLVE"), " ".join("%d:%.3f/%.3f" % o for o in out), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(); sys.exit(0)
    for mode in ("0", "1"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, PLSTVO_STREAM_SOLVE=mode), check=True)
