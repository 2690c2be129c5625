"""Host-side state machine of the handler: adaptive FAST threshold (src/stereoFrameHandler.cpp:66-86) and the
key-frame test (:1136-1218).  Three implementations are held together on synthetic sequences: the C++ host layer
(include/plstvo.hpp, compiled here with g++), the Python mirror (stvo_pl_b200/handler.py) and the C oracle."""
import os
import subprocess

import numpy as np
import pytest
from scipy.linalg import expm

from conftest import ROOT
from stvo_pl_b200 import handler as H


def _hat(x):
    t, w = x[:3], x[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = t
    return M


def make_sequence(n, seed, speed=0.4, rot=0.02, cov_scale=1e-6):
    """A forward-moving camera: per frame an increment DT, the chained pose since the last key frame is NOT known to the
    generator (the test decides), so Tfw is chained by the consumer.  Includes failed frames (DT = I, cov = 0, err = -1)."""
    rng = np.random.default_rng(seed)
    frames = []
    for i in range(n):
        if rng.random() < 0.08:
            frames.append(dict(DT=np.eye(4), DT_cov=np.zeros((6, 6)), err_norm=-1.0, n_inliers_pt=int(rng.integers(0, 40))))
            continue
        x = np.concatenate([rng.normal(0, 0.02, 2), [speed + rng.normal(0, 0.05)], rng.normal(0, rot, 3)])
        A = rng.normal(size=(6, 6))
        cov = cov_scale * (A @ A.T + 6 * np.eye(6)) * rng.uniform(0.5, 2.0)
        frames.append(dict(DT=expm(_hat(x)), DT_cov=cov, err_norm=float(rng.uniform(0.05, 0.7)),
                           n_inliers_pt=int(rng.integers(20, 260))))
    return frames


def run_python(frames):
    hc, kf, th, Tfw = H.HandlerConfig(), H.KeyframeTest(), H.HandlerConfig().orb_fast_th, np.eye(4)
    out, fed = [], []
    for fr in frames:
        Tfw = Tfw @ fr["DT"]                     # chained from the last key frame (optimizePose, :377)
        fed.append(Tfw.copy())
        new_kf = kf.needNewKF(hc, Tfw, fr["DT"], fr["DT_cov"])
        th = H.update_fast_threshold(hc, th, fr["DT"], fr["err_norm"], fr["n_inliers_pt"])
        out.append((int(new_kf), th, kf.entropy_curr, kf.entropy_ratio, kf.t, kf.r, kf.N_prevKF_currF))
        if new_kf:
            class _F:  # currFrameIsKF resets the frame's world pose
                pass
            f = _F()
            kf.currFrameIsKF(f)
            Tfw = f.Tfw
    return out, fed


def run_oracle(oracle, frames, fed):
    c, s = oracle.handler_default_config(), oracle.kf_state()
    th, out = c.orb_fast_th, []
    for fr, Tfw in zip(frames, fed):
        new_kf = oracle.need_new_kf(c, s, Tfw, fr["DT"], fr["DT_cov"])
        th = oracle.update_fast_threshold(c, th, fr["DT"], fr["err_norm"], fr["n_inliers_pt"])
        out.append((int(new_kf), th, s.entropy_curr, s.entropy_ratio, s.t, s.r, s.N_prevKF_currF))
        if new_kf:
            s = oracle.kf_state()
    return out


def run_cpp(tmp_path, frames, fed):
    exe = str(tmp_path / "kf_cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "kf_cpp.cpp"), "-o", exe], check=True)
    path = str(tmp_path / "seq.bin")
    with open(path, "wb") as f:
        for fr, Tfw in zip(frames, fed):
            rec = np.concatenate([Tfw.ravel(), fr["DT"].ravel(), fr["DT_cov"].ravel(), [fr["err_norm"], fr["n_inliers_pt"]]])
            f.write(rec.astype(np.float64).tobytes())
    lines = subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    return [tuple(float(v) for v in ln.split()) for ln in lines]


def _agree(a, b, what):
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert int(x[0]) == int(y[0]) and int(x[1]) == int(y[1]) and int(x[6]) == int(y[6]), (what, i, x, y)
        for k in (2, 3, 4, 5):
            if np.isfinite(x[k]) or np.isfinite(y[k]):
                assert x[k] == pytest.approx(y[k], rel=1e-9, abs=1e-12), (what, i, k, x, y)


@pytest.mark.parametrize("seed,speed,rot", [(0, 0.4, 0.02), (1, 1.5, 0.01), (2, 0.05, 0.12), (3, 0.3, 0.002)])
def test_keyframe_test_three_ways(tmp_path, oracle, seed, speed, rot):
    frames = make_sequence(80, seed, speed, rot)
    py, fed = run_python(frames)
    orc = run_oracle(oracle, frames, fed)
    cpp = run_cpp(tmp_path, frames, fed)
    _agree(py, orc, "python vs oracle")
    _agree(py, cpp, "python vs c++")
    kfs = sum(p[0] for p in py)
    assert 0 < kfs < len(frames)                 # both outcomes occur
    assert len({p[1] for p in py}) > 2           # the FAST threshold moves


def test_keyframe_known_answers(oracle):
    c, s = oracle.handler_default_config(), oracle.kf_state()
    cov = np.diag([1e-4, 2e-4, 3e-4, 1e-6, 2e-6, 3e-6])
    k = 3.0 * (1.0 + np.log(2.0 * np.pi))
    # first frame after a key frame with T = I: entropy_first = k + 0.5 log det(cov); accumulated cov = cov -> ratio 1
    assert oracle.need_new_kf(c, s, np.eye(4), np.eye(4), cov) is False
    assert s.entropy_first_prevKF == pytest.approx(k + 0.5 * np.log(np.linalg.det(cov)), rel=1e-13)
    assert s.entropy_ratio == pytest.approx(1.0, rel=1e-12) and s.N_prevKF_currF == 1 and s.t == 0.0 and s.r == 0.0
    # translation beyond max_kf_t_dist = 5 -> new key frame regardless of entropy (:1175)
    T = np.eye(4); T[2, 3] = 5.5
    assert oracle.need_new_kf(c, s, T, np.eye(4), cov) is True and s.t == pytest.approx(5.5)
    # a failed frame (DT = I, cov = 0) forces a key frame (:1174)
    s = oracle.kf_state()
    assert oracle.need_new_kf(c, s, np.eye(4), np.eye(4), np.zeros((6, 6))) is True
    assert s.entropy_first_prevKF == -999999999.99
    # uncTinv_se3 = Ad(T^-1) cov Ad(T^-1)^T; det is invariant under the adjoint of a rigid motion
    x = np.array([0.3, -0.2, 1.0, 0.05, -0.02, 0.1])
    Tm = expm(_hat(x))
    assert oracle.det6(oracle.unctinv_se3(Tm, cov)) == pytest.approx(np.linalg.det(cov), rel=1e-10)
    assert oracle.det6(np.eye(6) * 2.0) == 64.0


def test_fast_threshold_branches(oracle):
    c, hc = oracle.handler_default_config(), H.HandlerConfig()
    DT = np.eye(4); DT[2, 3] = 0.5
    cases = [(np.eye(4), 0.1, 200, 20, 10), (DT, 0.6, 200, 20, 10), (DT, 0.1, 40, 20, 10), (DT, 0.1, 80, 20, 15),
             (DT, 0.1, 120, 20, 20), (DT, 0.1, 160, 20, 25), (DT, 0.1, 500, 48, 50), (DT, 0.6, 10, 8, 5),
             (DT, 0.5, 120, 20, 20)]          # err == th is not "bad" (:75 is >)
    for dt, err, n, th, want in cases:
        assert oracle.update_fast_threshold(c, th, dt, err, n) == want
        assert H.update_fast_threshold(hc, th, dt, err, n) == want
    c.adaptative_fast, hc.adaptative_fast = 0, False
    assert oracle.update_fast_threshold(c, 20, np.eye(4), 0.9, 0) == 20 == H.update_fast_threshold(hc, 20, np.eye(4), 0.9, 0)


def test_chain_poses_cpp_vs_python(tmp_path):
    """plstvo::chainPoses (C++ host layer) == handler.chain_poses (Python mirror) on a synthetic sequence with failed pairs."""
    from stvo_pl_b200 import types as T
    rng = np.random.default_rng(4)
    n = 40
    res = np.zeros(n, dtype=T.POSE_RESULT_DTYPE)
    for k in range(n):
        x = np.concatenate([rng.normal([0, 0, 0.5], 0.05), rng.normal(0, 0.02, 3)])
        A = rng.normal(size=(6, 6))
        res["DT"][k], res["DT_cov"][k], res["good"][k] = expm(_hat(x)), 1e-6 * (A @ A.T + np.eye(6)), int(rng.random() > 0.15)
    exe = str(tmp_path / "chain_cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "chain_cpp.cpp"), "-o", exe], check=True)
    path = str(tmp_path / "res.bin")
    with open(path, "wb") as f:
        for k in range(n):
            f.write(np.concatenate([[float(res["good"][k])], res["DT"][k].ravel(), res["DT_cov"][k].ravel()]).astype(np.float64).tobytes())
    lines = subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    got = np.array([[float(v) for v in ln.split()] for ln in lines])
    ref = H.chain_poses(res.copy())
    assert 0 < (res["good"] == 0).sum() < n
    np.testing.assert_allclose(got[:, :16].reshape(n, 4, 4), ref["Tfw"], atol=1e-12)
    np.testing.assert_allclose(got[:, 16:].reshape(n, 6, 6), ref["Tfw_cov"], rtol=1e-10, atol=1e-15)
    k0 = int(np.nonzero(res["good"] == 0)[0][0])
    if k0 > 0:
        np.testing.assert_array_equal(ref["Tfw"][k0], ref["Tfw"][k0 - 1])      # a failed pair carries the pose over (:388-389)
