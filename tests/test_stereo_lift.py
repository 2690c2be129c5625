"""3-D lifting of the stereo matches (src/stereoFrame.cpp:149-172, :348-397).
CPU: the C oracle against an independent Python restatement (tests/ref_lift.py) and source-derived known answers.
GPU: plstvo_stereo_lift_points / _lines against the oracle, bitwise (the kernel uses the same IEEE operations)."""
import numpy as np
import pytest

from stvo_pl_b200 import types as T
import ref_lift
from stvo_pl_b200.stereo_synth import make_lift_lines as make_lines, make_lift_points as make_points

KEYS_PT = ("pl", "disp", "P", "sigma2", "level", "desc", "src_idx")
KEYS_LS = ("spl", "epl", "sdisp", "edisp", "sP", "eP", "le", "angle", "sigma2", "level", "desc", "src_idx")


def _same(a, b, key):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (key, a.shape, b.shape)
    if a.dtype.kind == "f":
        assert np.array_equal(a.view(np.uint64) if a.size else a, b.view(np.uint64) if b.size else b), key   # bitwise, NaN-safe
    else:
        assert np.array_equal(a, b), key


def _as_arrays(ref, keys):
    shapes = dict(pl=(0, 2), P=(0, 3), spl=(0, 2), epl=(0, 2), sP=(0, 3), eP=(0, 3), le=(0, 3), desc=(0, 32))
    out = {}
    for k in keys:
        dt = np.int32 if k in ("level", "src_idx") else np.uint8 if k == "desc" else np.float64
        v = np.asarray(ref[k], dt)
        out[k] = v.reshape(shapes.get(k, (0,))) if v.size == 0 else v
    return out


# ---------------------------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("n_l,n_r,seed", [(0, 0, 0), (1, 1, 1), (37, 50, 2), (600, 580, 3), (2000, 1900, 4)])
def test_oracle_points_vs_python(oracle, n_l, n_r, seed):
    cam, sc = T.kitti_camera(), T.default_stereo_config()
    args = make_points(n_l, n_r, seed)
    k, got = oracle.stereo_lift_points(cam, sc, *args)
    ref = _as_arrays(ref_lift.lift_points(cam, sc, *args), KEYS_PT)
    assert k == len(ref["disp"])
    if n_l >= 37:
        assert 0 < k < (args[4] >= 0).sum()            # both filters reject something
    for key in KEYS_PT:
        _same(got[key], ref[key], key)


@pytest.mark.parametrize("n_l,n_r,seed", [(0, 0, 0), (1, 1, 1), (40, 31, 2), (300, 320, 3), (1200, 1100, 4)])
def test_oracle_lines_vs_python(oracle, n_l, n_r, seed):
    cam, sc = T.kitti_camera(), T.default_stereo_config()
    args = make_lines(n_l, n_r, seed)
    k, got = oracle.stereo_lift_lines(cam, sc, *args)
    ref = _as_arrays(ref_lift.lift_lines(cam, sc, *args), KEYS_LS)
    assert k == len(ref["sdisp"])
    if n_l >= 40:
        assert 0 < k < (args[5] >= 0).sum()
    for key in KEYS_LS:
        _same(got[key], ref[key], key)


def test_oracle_kitti_epipolar_zero(oracle):
    """config_kitti.yaml:17 sets max_dist_epip = 0: only matches on exactly the same row survive (:156 is `<=`)."""
    cam, sc = T.kitti_camera(), T.default_stereo_config()
    sc.max_dist_epip = 0.0
    kp_l, octave, desc, kp_r, m12 = make_points(400, 400, 9)
    k, got = oracle.stereo_lift_points(cam, sc, kp_l, octave, desc, kp_r, m12)
    assert k > 0
    assert np.all(kp_l[got["src_idx"], 1] == kp_r[m12[got["src_idx"]], 1])


def test_oracle_known_answers(oracle):
    cam, sc = T.kitti_camera(), T.default_stereo_config()
    # one point at (cx + 10, cy - 5) with disparity 4: P = b/4 * (10, -5, fx)  (src/pinholeStereoCamera.cpp:221-229)
    kp_l = np.array([[cam.cx + 10, cam.cy - 5]], np.float32)
    kp_r = kp_l - np.array([[4, 0]], np.float32)
    k, got = oracle.stereo_lift_points(cam, sc, kp_l, [2], np.zeros((1, 32), np.uint8), kp_r, [0])
    assert k == 1 and got["disp"][0] == 4.0 and got["level"][0] == 2
    u, v = float(kp_l[0, 0]), float(kp_l[0, 1])
    np.testing.assert_allclose(got["P"][0], [cam.b / 4 * (u - cam.cx), cam.b / 4 * (v - cam.cy), cam.b / 4 * cam.fx], rtol=1e-15)
    np.testing.assert_allclose(got["sigma2"][0], 1.0 / 1.2 ** 4, rtol=1e-14)     # src/stereoFeatures.cpp:43-45
    # lineSegmentOverlapStereo (:473-508)
    ov = oracle.line_segment_overlap_stereo
    assert ov(sc, 10.0, 10.05, 0.0, 100.0) == 1.0            # near-horizontal: untouched initial value
    assert ov(sc, 0.0, 10.0, 20.0, 30.0) == 0.0              # disjoint
    assert ov(sc, 0.0, 10.0, -5.0, 15.0) == pytest.approx(10.0 / 15.0)   # projection covers the observation: (eln-sln)/(eln-spn)
    assert ov(sc, 0.0, 10.0, 5.0, 20.0) == 1.0               # (10-5)/(10-5)
    assert ov(sc, 10.0, 0.0, 2.0, 8.0) == pytest.approx(6.0 / 8.0)       # reversed observation
    assert ov(sc, 0.0, 10.0, 9.995, 20.0) == 0.0             # length <= 0.01f
    # a vertical segment seen with disparity 8 at both ends
    seg_l = np.array([[100, 50, 100, 150]], np.float32)
    seg_r = np.array([[92, 50, 92, 150]], np.float32)
    k, got = oracle.stereo_lift_lines(cam, sc, seg_l, [0.5], [1], np.zeros((1, 32), np.uint8), seg_r, [0])
    assert k == 1 and got["sdisp"][0] == 8.0 and got["edisp"][0] == 8.0
    np.testing.assert_allclose(got["le"][0], [-1.0, 0.0, 100.0], rtol=1e-15)
    np.testing.assert_allclose(got["sigma2"][0], 1.0 / 1.2 ** 2, rtol=1e-14)
    assert got["angle"][0] == float(np.float32(0.5))
    # exactly horizontal right segment: 0/0 -> NaN disparities -> rejected (every comparison false), no crash
    seg_r2 = np.array([[92, 50, 192, 50]], np.float32)
    k, _ = oracle.stereo_lift_lines(cam, sc, seg_l, [0.5], [1], np.zeros((1, 32), np.uint8), seg_r2, [0])
    assert k == 0


# ---------------------------------------------------------------------------------------------- GPU: CUDA vs oracle
def _batch_points(sizes, seed):
    frames = [make_points(n_l, n_r, seed + 17 * i) for i, (n_l, n_r) in enumerate(sizes)]
    l_off = np.concatenate([[0], np.cumsum([len(f[0]) for f in frames])]).astype(np.int32)
    r_off = np.concatenate([[0], np.cumsum([len(f[3]) for f in frames])]).astype(np.int32)
    cat = [np.concatenate([f[j] for f in frames]) for j in range(5)]
    return frames, l_off, r_off, cat


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [[(600, 580)], [(0, 0), (37, 50), (1, 1), (0, 5), (900, 800), (255, 256), (257, 256)],
                                   [(2000, 1900)] * 3, [(20000, 20000)]])
def test_gpu_points_vs_oracle(engine, oracle, sizes):
    cam, sc = T.kitti_camera(), T.default_stereo_config()
    frames, l_off, r_off, (kp_l, octave, desc, kp_r, m12) = _batch_points(sizes, 100)
    total, out = engine.stereo_lift_points(cam, sc, l_off, kp_l, octave, desc, r_off, kp_r, m12)
    tot_ref = 0
    for p, f in enumerate(frames):
        k, ref = oracle.stereo_lift_points(cam, sc, *f)
        assert out["counts"][p] == k
        tot_ref += k
        a = l_off[p]
        for key in KEYS_PT:
            _same(out[key][a:a + k], ref[key], key)
    assert total == tot_ref


def _batch_lines(sizes, seed):
    frames = [make_lines(n_l, n_r, seed + 17 * i) for i, (n_l, n_r) in enumerate(sizes)]
    l_off = np.concatenate([[0], np.cumsum([len(f[0]) for f in frames])]).astype(np.int32)
    r_off = np.concatenate([[0], np.cumsum([len(f[4]) for f in frames])]).astype(np.int32)
    cat = [np.concatenate([f[j] for f in frames]) for j in range(6)]
    return frames, l_off, r_off, cat


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [[(300, 320)], [(0, 0), (40, 31), (1, 1), (0, 3), (700, 650), (256, 256), (257, 300)],
                                   [(1200, 1100)] * 3, [(15000, 15000)]])
def test_gpu_lines_vs_oracle(engine, oracle, sizes):
    cam, sc = T.kitti_camera(), T.default_stereo_config()
    frames, l_off, r_off, (seg_l, angle, octave, desc, seg_r, m12) = _batch_lines(sizes, 200)
    total, out = engine.stereo_lift_lines(cam, sc, l_off, seg_l, angle, octave, desc, r_off, seg_r, m12)
    tot_ref = 0
    for p, f in enumerate(frames):
        k, ref = oracle.stereo_lift_lines(cam, sc, *f)
        assert out["counts"][p] == k
        tot_ref += k
        a = l_off[p]
        for key in KEYS_LS:
            _same(out[key][a:a + k], ref[key], key)
    assert total == tot_ref


@pytest.mark.gpu
def test_gpu_lift_kitti_config_and_errors(engine, oracle):
    cam, sc = T.kitti_camera(), T.default_stereo_config()
    sc.max_dist_epip = 0.0                                     # config_kitti.yaml:17
    kp_l, octave, desc, kp_r, m12 = make_points(500, 500, 5)
    l_off, r_off = np.array([0, 500], np.int32), np.array([0, 500], np.int32)
    total, out = engine.stereo_lift_points(cam, sc, l_off, kp_l, octave, desc, r_off, kp_r, m12)
    k, ref = oracle.stereo_lift_points(cam, sc, kp_l, octave, desc, kp_r, m12)
    assert total == k > 0
    for key in KEYS_PT:
        _same(out[key][:k], ref[key], key)
    with pytest.raises(RuntimeError):                          # offsets must start at 0
        engine.stereo_lift_points(cam, sc, np.array([1, 500], np.int32), kp_l, octave, desc, r_off, kp_r, m12)


def test_default_stereo_config_matches_library():
    """plstvo_default_stereo_config and the Python preset agree (src/config.cpp:58-69, :96, :106); no GPU needed."""
    from stvo_pl_b200.engine import load_library
    import ctypes as C
    lib = load_library()
    c = T.PlStereoConfig()
    lib.plstvo_default_stereo_config(C.byref(c))
    d = T.default_stereo_config()
    for name, _ in T.PlStereoConfig._fields_:
        assert getattr(c, name) == getattr(d, name), name
