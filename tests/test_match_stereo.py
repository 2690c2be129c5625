"""StereoFrame::matchStereoPoints / matchStereoLines in one device pass (plstvo_match_stereo_*; src/stereoFrame.cpp:120-173,
:309-398): grid cells from raw key points / key lines -> matchGrid -> 3-D lifting.  Held bitwise against the composition of the
oracle's matchGrid and lifting restatements on the grid coordinates the reference's caller forms (stereo_synth.stereo_cells_*)."""
import ctypes as C

import numpy as np
import pytest

from stvo_pl_b200 import stereo_synth as SS, types as T

KEYS_PT = ("pl", "disp", "P", "sigma2", "level", "desc", "src_idx")
KEYS_LS = ("spl", "epl", "sdisp", "edisp", "sP", "eP", "le", "angle", "sigma2", "level", "desc", "src_idx")


def _same(a, b, key):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (key, a.shape, b.shape)
    if a.dtype.kind == "f" and a.size:
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), key
    else:
        assert np.array_equal(a, b), key


def _window(mc):
    return T.PlGridWindow(mc.matching_s_ws, 0, 0, 0)


def oracle_points(oracle, cam, mc, sc, frame):
    kp_l, octave, d1, kp_r, d2 = frame
    q_cell, t_cell = SS.stereo_cells_points(kp_l, kp_r, cam.width, cam.height)
    _, m12 = oracle.match_grid_points(q_cell, d1, t_cell, d2, _window(mc), mc.min_ratio_12_p, bool(mc.best_lr_matches),
                                      mc.grid_rows, mc.grid_cols)
    k, rec = oracle.stereo_lift_points(cam, sc, kp_l, octave, d1, kp_r, m12)
    return m12, k, rec


def oracle_lines(oracle, cam, mc, sc, frame):
    seg_l, angle, octave, d1, seg_r, d2 = frame
    q_line, t_line, t_dir = SS.stereo_cells_lines(seg_l, seg_r, cam.width, cam.height)
    _, m12 = oracle.match_grid_lines(q_line, d1, t_line, t_dir, d2, _window(mc), mc.min_ratio_12_p, mc.line_sim_th,
                                     bool(mc.best_lr_matches), mc.grid_rows, mc.grid_cols)
    k, rec = oracle.stereo_lift_lines(cam, sc, seg_l, angle, octave, d1, seg_r, m12)
    return m12, k, rec


def test_oracle_fused_equals_composition(oracle):
    """orc_match_stereo_* (cells formed in C like the reference's caller) == numpy cells + orc_match_grid_* + orc_stereo_lift_*;
    the threaded batch returns the same counts."""
    cam, mc, sc = T.kitti_camera(), T.default_stereo_match_config(), T.default_stereo_config()
    pts = [SS.make_stereo_frame_points(a, b, seed=30 + i) for i, (a, b) in enumerate([(700, 650), (0, 0), (5, 0), (300, 310)])]
    lns = [SS.make_stereo_frame_lines(a, b, seed=40 + i) for i, (a, b) in enumerate([(200, 190), (0, 3), (1, 1), (90, 95)])]
    counts = []
    for fp, fl in zip(pts, lns):
        m12, k, rec = oracle.match_stereo_points(cam, mc, sc, fp[0], fp[1], fp[2], fp[3], fp[4])
        m12b, kb, recb = oracle_points(oracle, cam, mc, sc, fp)
        np.testing.assert_array_equal(m12, m12b)
        assert k == kb
        for key in KEYS_PT:
            _same(rec[key], recb[key], key)
        m12, k2, rec = oracle.match_stereo_lines(cam, mc, sc, fl[0], fl[1], fl[2], fl[3], fl[4], fl[5])
        m12b, kb, recb = oracle_lines(oracle, cam, mc, sc, fl)
        np.testing.assert_array_equal(m12, m12b)
        assert k2 == kb
        for key in KEYS_LS:
            _same(rec[key], recb[key], key)
        counts.append((k, k2))
    pl_off, pr_off, P = _cat(pts, 0, 3)
    ll_off, lr_off, L = _cat(lns, 0, 4)
    got = oracle.stereo_batch(cam, mc, sc, pl_off, P[0], P[1], P[2], pr_off, P[3], P[4], ll_off, L[0], L[1], L[2], L[3], lr_off,
                              L[4], L[5], threads=3)
    np.testing.assert_array_equal(got, np.array(counts, np.int32))


def _cat(frames, idx_l, idx_r):
    l_off = np.concatenate([[0], np.cumsum([len(f[idx_l]) for f in frames])]).astype(np.int32)
    r_off = np.concatenate([[0], np.cumsum([len(f[idx_r]) for f in frames])]).astype(np.int32)
    return l_off, r_off, [np.concatenate([f[j] for f in frames]) for j in range(len(frames[0]))]


def test_default_stereo_match_config_matches_library():
    from stvo_pl_b200.engine import load_library
    lib = load_library()
    c = T.PlStereoMatchConfig()
    lib.plstvo_default_stereo_match_config(C.byref(c))
    d = T.default_stereo_match_config()
    for name, _ in T.PlStereoMatchConfig._fields_:
        assert getattr(c, name) == getattr(d, name), name


def test_oracle_composition_finds_the_planted_matches(oracle):
    """Sanity of the generator + composition on the CPU: most planted stereo pairs survive matchGrid and the lifting."""
    cam, mc, sc = T.kitti_camera(), T.default_stereo_match_config(), T.default_stereo_config()
    m12, k, rec = oracle_points(oracle, cam, mc, sc, SS.make_stereo_frame_points(900, 850, seed=3))
    assert (m12 >= 0).sum() > 300 and 200 < k <= (m12 >= 0).sum()
    assert np.all(rec["disp"] >= sc.min_disp) and np.all(rec["P"][:, 2] > 0)
    m12, k, rec = oracle_lines(oracle, cam, mc, sc, SS.make_stereo_frame_lines(300, 280, seed=4))
    assert (m12 >= 0).sum() > 60 and 10 < k <= (m12 >= 0).sum()
    np.testing.assert_allclose(np.hypot(rec["le"][:, 0], rec["le"][:, 1]), 1.0, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("sizes,kitti", [([(900, 850)], False), ([(0, 0), (40, 60), (1, 1), (0, 7), (1500, 1400), (7, 0), (600, 610)], False),
                                          ([(2000, 2000)] * 3, True)])
def test_gpu_match_stereo_points(engine, oracle, sizes, kitti):
    cam, mc, sc = T.kitti_camera(), T.default_stereo_match_config(), T.default_stereo_config()
    if kitti:
        mc.min_ratio_12_p, sc.max_dist_epip = 0.75, 0.0              # config_kitti.yaml:17, :19
    frames = [SS.make_stereo_frame_points(a, b, seed=50 + i) for i, (a, b) in enumerate(sizes)]
    l_off, r_off, (kp_l, octave, d1, kp_r, d2) = _cat(frames, 0, 3)
    total, out = engine.match_stereo_points(cam, mc, sc, l_off, kp_l, octave, d1, r_off, kp_r, d2)
    tot_ref = 0
    for p, f in enumerate(frames):
        m12, k, rec = oracle_points(oracle, cam, mc, sc, f)
        a = l_off[p]
        np.testing.assert_array_equal(out["m12"][a:a + len(m12)], m12)
        assert out["counts"][p] == k
        for key in KEYS_PT:
            _same(out[key][a:a + k], rec[key], key)
        tot_ref += k
    assert total == tot_ref and (not kitti or total > 0)
    # the same through the two separate entry points
    q_cell, t_cell = SS.stereo_cells_points(kp_l, kp_r, cam.width, cam.height)
    _, m12b, _ = engine.match_grid_points(l_off, q_cell, d1, r_off, t_cell, d2, _window(mc), mc.min_ratio_12_p,
                                          bool(mc.best_lr_matches))
    np.testing.assert_array_equal(m12b, out["m12"])
    total2, out2 = engine.stereo_lift_points(cam, sc, l_off, kp_l, octave, d1, r_off, kp_r, m12b)
    assert total2 == total
    np.testing.assert_array_equal(out2["counts"], out["counts"])
    for p in range(len(frames)):                                     # slots beyond counts[p] are unspecified
        a, k = l_off[p], int(out["counts"][p])
        for key in KEYS_PT:
            _same(out2[key][a:a + k], out[key][a:a + k], key)


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [[(300, 280)], [(0, 0), (30, 45), (1, 1), (0, 4), (520, 500), (5, 0), (257, 256)], [(500, 500)] * 3])
def test_gpu_match_stereo_lines(engine, oracle, sizes):
    cam, mc, sc = T.kitti_camera(), T.default_stereo_match_config(), T.default_stereo_config()
    frames = [SS.make_stereo_frame_lines(a, b, seed=70 + i) for i, (a, b) in enumerate(sizes)]
    l_off, r_off, (seg_l, angle, octave, d1, seg_r, d2) = _cat(frames, 0, 4)
    total, out = engine.match_stereo_lines(cam, mc, sc, l_off, seg_l, angle, octave, d1, r_off, seg_r, d2)
    tot_ref = 0
    for p, f in enumerate(frames):
        m12, k, rec = oracle_lines(oracle, cam, mc, sc, f)
        a = l_off[p]
        np.testing.assert_array_equal(out["m12"][a:a + len(m12)], m12)
        assert out["counts"][p] == k
        for key in KEYS_LS:
            _same(out[key][a:a + k], rec[key], key)
        tot_ref += k
    assert total == tot_ref
    q_line, t_line, t_dir = SS.stereo_cells_lines(seg_l, seg_r, cam.width, cam.height)
    _, m12b, _ = engine.match_grid_lines(l_off, q_line, d1, r_off, t_line, t_dir, d2, _window(mc), mc.min_ratio_12_p,
                                         mc.line_sim_th, bool(mc.best_lr_matches))
    np.testing.assert_array_equal(m12b, out["m12"])


@pytest.mark.gpu
def test_gpu_match_stereo_feeds_the_tracker(engine, oracle):
    """The lifted records are the tracker's PlFrameBatch fields: two stereo frames of the same scene go through
    matchStereoPoints and then plstvo_track_batch, and the pose agrees with the oracle run on the same records."""
    cam, mc, sc, cfg = T.kitti_camera(), T.default_stereo_match_config(), T.default_stereo_config(), T.kitti_config()
    cfg.has_lines = 0
    frames = []
    for seed in (11, 12):
        kp_l, octave, d1, kp_r, d2 = SS.make_stereo_frame_points(1200, 1200, seed=11, overlap=0.9)   # same scene twice ...
        if seed == 12:                                                # ... second frame: everything moved 1.5 px to the right
            kp_l, kp_r = kp_l + np.float32([1.5, 0]), kp_r + np.float32([1.5, 0])
        off = np.array([0, len(kp_l)], np.int32)
        tot, out = engine.match_stereo_points(cam, mc, sc, off, kp_l, octave, d1, off, kp_r, d2)
        k = int(out["counts"][0])
        assert k > 300
        z = np.zeros((0, 3))
        frames.append(T.FrameBatch(pt_off=[0, k], ls_off=[0, 0], pdesc=out["desc"][:k], ldesc=np.zeros((0, 32), np.uint8),
                                   pt_P=out["P"][:k], pt_pl=out["pl"][:k], pt_sigma2=out["sigma2"][:k], ls_sP=z, ls_eP=z, ls_le=z,
                                   ls_spl=np.zeros((0, 2)), ls_epl=np.zeros((0, 2)), ls_sigma2=np.zeros(0),
                                   ls_level=np.zeros(0, np.int32)))
    got = engine.track_batch(cam, cfg, frames[0], frames[1])
    ref = oracle.track_batch(cam, cfg, frames[0], frames[1])
    np.testing.assert_array_equal(got["m12_pt"], ref["m12_pt"])
    assert got["results"]["status"][0] == ref["results"]["status"][0]
    np.testing.assert_allclose(got["results"]["DT"][0], ref["results"]["DT"][0], atol=1e-8)
    assert got["results"]["n_matched_pt"][0] > 250


@pytest.mark.gpu
def test_gpu_match_stereo_errors(engine):
    cam, mc, sc = T.kitti_camera(), T.default_stereo_match_config(), T.default_stereo_config()
    kp_l, octave, d1, kp_r, d2 = SS.make_stereo_frame_points(50, 50, seed=1)
    off = np.array([0, 50], np.int32)
    bad = T.default_stereo_match_config()
    bad.grid_rows = 0
    with pytest.raises(RuntimeError):
        engine.match_stereo_points(cam, bad, sc, off, kp_l, octave, d1, off, kp_r, d2)
    with pytest.raises(RuntimeError):                              # offsets must start at 0
        engine.match_stereo_points(cam, mc, sc, np.array([1, 50], np.int32), kp_l, octave, d1, off, kp_r, d2)
    # more than 128 right key points in one query window: no limit, as in the reference (the frame takes matchGrid's
    # sequential path); exact against the oracle
    kp_r2 = np.tile(np.float32([[100.0, 100.0]]), (200, 1))
    kp_l2 = np.tile(np.float32([[110.0, 100.0]]), (10, 1))
    rng = np.random.default_rng(5)
    d2_200 = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    tot, out = engine.match_stereo_points(cam, mc, sc, np.array([0, 10], np.int32), kp_l2, np.zeros(10, np.int32), d1[:10],
                                          np.array([0, 200], np.int32), kp_r2, d2_200)
    from oracle.oracle import Oracle
    m12_o, k_o, _ = Oracle().match_stereo_points(cam, mc, sc, kp_l2, np.zeros(10, np.int32), d1[:10], kp_r2, d2_200)
    np.testing.assert_array_equal(out["m12"], m12_o)
    assert tot == k_o


# ---------------------------------------------------------------------------- raw stereo features -> pose (records stay in HBM)
def _oracle_frames(oracle, cam, mc, sc, d):
    """Oracle stereo step for every frame of a PlStereoFeatures dict -> FrameBatch of the lifted records + counts."""
    B = len(d["pl_off"]) - 1
    recs_p, recs_l, counts = [], [], []
    for p in range(B):
        a, b, ar, br = d["pl_off"][p], d["pl_off"][p + 1], d["pr_off"][p], d["pr_off"][p + 1]
        _, k, rp = oracle.match_stereo_points(cam, mc, sc, d["kp_l"][a:b], d["poct_l"][a:b], d["pdesc_l"][a:b], d["kp_r"][ar:br],
                                              d["pdesc_r"][ar:br])
        c, e, cr, er = d["ll_off"][p], d["ll_off"][p + 1], d["lr_off"][p], d["lr_off"][p + 1]
        _, kl, rl = oracle.match_stereo_lines(cam, mc, sc, d["seg_l"][c:e], d["angle_l"][c:e], d["loct_l"][c:e], d["ldesc_l"][c:e],
                                              d["seg_r"][cr:er], d["ldesc_r"][cr:er])
        recs_p.append(rp); recs_l.append(rl); counts.append((k, kl))
    cat = lambda recs, key: np.concatenate([r[key] for r in recs])
    fb = T.FrameBatch(pt_off=np.concatenate([[0], np.cumsum([c[0] for c in counts])]),
                      ls_off=np.concatenate([[0], np.cumsum([c[1] for c in counts])]),
                      pdesc=cat(recs_p, "desc"), ldesc=cat(recs_l, "desc"), pt_P=cat(recs_p, "P"), pt_pl=cat(recs_p, "pl"),
                      pt_sigma2=cat(recs_p, "sigma2"), ls_sP=cat(recs_l, "sP"), ls_eP=cat(recs_l, "eP"), ls_le=cat(recs_l, "le"),
                      ls_spl=cat(recs_l, "spl"), ls_epl=cat(recs_l, "epl"), ls_sigma2=cat(recs_l, "sigma2"),
                      ls_level=cat(recs_l, "level"))
    return fb, np.array(counts, np.int32)


@pytest.mark.gpu
@pytest.mark.parametrize("B,n_pt,n_ls", [(3, 900, 200), (1, 300, 0), (6, 1500, 350)])
def test_gpu_track_stereo_batch_vs_oracle_chain(engine, oracle, B, n_pt, n_ls):
    """plstvo_track_stereo_batch == oracle(matchStereoPoints / Lines of both frames) -> oracle(f2fTracking + optimizePose):
    survivor counts and match counts equal, pose within 1e-9 rad / 1e-8 m, and close to the ground-truth motion."""
    import ref_numpy as R
    prev, curr, Tgt, cam = SS.make_stereo_pairs(B, n_pt=n_pt, n_ls=max(n_ls, 1), seed=B)
    mc, sc, cfg = T.default_stereo_match_config(), T.default_stereo_config(), T.kitti_config()
    if n_ls == 0:
        cfg.has_lines = 0
    res, n_stereo = engine.track_stereo_batch(cam, cfg, mc, sc, prev, curr)
    fbp, cp = _oracle_frames(oracle, cam, mc, sc, prev)
    fbc, cc = _oracle_frames(oracle, cam, mc, sc, curr)
    np.testing.assert_array_equal(n_stereo, np.concatenate([cp, cc], 1))
    ref = oracle.track_batch(cam, cfg, fbp, fbc)["results"]
    for p in range(B):
        assert res["status"][p] == ref["status"][p] and res["good"][p] == ref["good"][p] == 1
        assert res["n_matched_pt"][p] == ref["n_matched_pt"][p] and res["n_matched_ls"][p] == ref["n_matched_ls"][p]
        assert res["n_inliers"][p] == ref["n_inliers"][p]
        ang, tr = R.pose_error(res["DT"][p], ref["DT"][p])
        assert ang < 1e-9 and tr < 1e-8
        ang, tr = R.pose_error(res["DT_opt"][p], Tgt[p])            # DT_opt: prev -> curr, like the generator's T
        assert ang < 5e-3 and tr < 5e-2


@pytest.mark.gpu
def test_gpu_track_stereo_batch_edge_cases(engine, oracle):
    """Frames whose stereo step leaves too few features (optimizePose's FEW_BEFORE branch), priors passed through, and the
    error paths of the fused entry point."""
    prev, curr, Tgt, cam = SS.make_stereo_pairs(3, n_pt=400, n_ls=60, seed=21)
    mc, sc, cfg = T.default_stereo_match_config(), T.default_stereo_config(), T.kitti_config()
    # pair 1: the previous frame's right-image features lie outside the image -> outside the grid (GridStructure::at's
    # out_of_bounds list) -> no stereo match at all
    a, b = prev["pr_off"][1], prev["pr_off"][2]
    prev["kp_r"] = prev["kp_r"].copy()
    prev["kp_r"][a:b, 0] -= 5000.0
    c, e = prev["lr_off"][1], prev["lr_off"][2]
    prev["seg_r"] = prev["seg_r"].copy()
    prev["seg_r"][c:e, 0] -= 5000.0
    prev["seg_r"][c:e, 2] -= 5000.0
    pri = T.identity_priors(3)
    pri["Tfw"][2][:3, 3] = [1.0, 2.0, 3.0]                          # chained into Tfw of pair 2 (:377)
    res, n_st = engine.track_stereo_batch(cam, cfg, mc, sc, prev, curr, priors=pri)
    fbp, cp = _oracle_frames(oracle, cam, mc, sc, prev)
    fbc, cc = _oracle_frames(oracle, cam, mc, sc, curr)
    np.testing.assert_array_equal(n_st, np.concatenate([cp, cc], 1))
    ref = oracle.track_batch(cam, cfg, fbp, fbc, priors=pri)["results"]
    np.testing.assert_array_equal(res["status"], ref["status"])
    np.testing.assert_array_equal(res["n_inliers"], ref["n_inliers"])
    assert n_st[1, 0] == 0 and n_st[1, 1] == 0 and res["status"][1] == T.ST_FEW_BEFORE and res["good"][1] == 0 and res["err_norm"][1] == -1.0
    np.testing.assert_allclose(res["Tfw"], ref["Tfw"], atol=1e-8)
    np.testing.assert_allclose(res["DT"], ref["DT"], atol=1e-8)
    # errors
    bad = dict(curr)
    bad["pl_off"], bad["pr_off"], bad["ll_off"], bad["lr_off"] = (curr[k][:-1] for k in ("pl_off", "pr_off", "ll_off", "lr_off"))
    with pytest.raises(RuntimeError):
        engine.track_stereo_batch(cam, cfg, mc, sc, prev, bad)     # different numbers of frames


@pytest.mark.gpu
def test_gpu_track_stereo_sequence(engine, oracle):
    """Sequence mode: NF consecutive frames, every frame through the stereo step once.  The NF - 1 results equal, byte for byte,
    those of plstvo_track_stereo_batch on (frames[:-1], frames[1:]); chained on the host they equal the handler fed frame by
    frame (priors = the previous frame's chained pose), and follow the generator's trajectory."""
    import ref_numpy as R
    from stvo_pl_b200 import handler as Hd
    NF = 7
    frames, rel, cam = SS.make_stereo_sequence(NF, n_pt=900, n_ls=180, seed=8)
    mc, sc, cfg = T.default_stereo_match_config(), T.default_stereo_config(), T.kitti_config()
    res, n_st = engine.track_stereo_sequence(cam, cfg, mc, sc, frames)
    assert len(res) == NF - 1 and n_st.shape == (NF, 2) and (res["good"] == 1).all()
    ref, n_b = engine.track_stereo_batch(cam, cfg, mc, sc, SS.stereo_frames_slice(frames, 0, NF - 1), SS.stereo_frames_slice(frames, 1, NF))
    assert res.tobytes() == ref.tobytes()
    np.testing.assert_array_equal(n_st[:-1], n_b[:, :2])
    np.testing.assert_array_equal(n_st[1:], n_b[:, 2:])
    # chaining: host scan == pairs solved one after the other with the previous chained pose as prior
    chained = Hd.chain_poses(res.copy())
    Tfw, cov = np.eye(4), np.eye(6)
    for k in range(NF - 1):
        pri = T.identity_priors(1)
        pri["Tfw"][0], pri["Tfw_cov"][0] = Tfw, cov
        one, _ = engine.track_stereo_batch(cam, cfg, mc, sc, SS.stereo_frames_slice(frames, k, k + 1),
                                           SS.stereo_frames_slice(frames, k + 1, k + 2), priors=pri)
        Tfw, cov = one["Tfw"][0], one["Tfw_cov"][0]
        np.testing.assert_allclose(chained["Tfw"][k], Tfw, atol=1e-9)
        np.testing.assert_allclose(chained["Tfw_cov"][k], cov, rtol=1e-7, atol=1e-12)
    # trajectory: DT_opt (prev -> curr) against the generator's relative motion
    for k in range(NF - 1):
        ang, tr = R.pose_error(res["DT_opt"][k], rel[k])
        assert ang < 5e-3 and tr < 5e-2
    # a single frame: nothing to do; two frames: one result
    r1, _ = engine.track_stereo_sequence(cam, cfg, mc, sc, SS.stereo_frames_slice(frames, 0, 1))
    assert len(r1) == 0
    r2, _ = engine.track_stereo_sequence(cam, cfg, mc, sc, SS.stereo_frames_slice(frames, 2, 4))
    assert r2.tobytes() == res[2:3].tobytes()


@pytest.mark.gpu
def test_gpu_track_stereo_async_equals_blocking(engine):
    """plstvo_track_stereo_batch_async / _sequence_async + plstvo_wait, two batches in flight, different inputs per slot."""
    mc, sc, cfg = T.default_stereo_match_config(), T.default_stereo_config(), T.kitti_config()
    batches = [SS.make_stereo_pairs(4, n_pt=500 + 100 * k, n_ls=100, seed=60 + k) for k in range(3)]
    cam = batches[0][3]
    ref = [engine.track_stereo_batch(cam, cfg, mc, sc, b[0], b[1]) for b in batches]
    pin = lambda d: {k: engine.pinned.copy(np.ascontiguousarray(v, T.STEREO_FEATURE_DTYPES[k])) for k, v in d.items()}
    cs = [(T.stereo_features_as_c(pin(b[0])), T.stereo_features_as_c(pin(b[1]))) for b in batches]
    res = [engine.pinned.empty((4,), T.POSE_RESULT_DTYPE) for _ in batches]
    nst = [engine.pinned.empty((4, 4), np.int32) for _ in batches]
    tickets = []
    for k, ((pc, _kp), (cc, _kc)) in enumerate(cs):
        tickets.append(engine.track_stereo_batch_async(cam, cfg, mc, sc, pc, cc, res[k], nst[k]))
        if k >= 1:
            engine.wait(tickets[k - 1])
    engine.wait(tickets[-1])
    for k in range(3):
        assert res[k].tobytes() == ref[k][0].tobytes()
        np.testing.assert_array_equal(nst[k], ref[k][1])
    frames, _, scam = SS.make_stereo_sequence(5, n_pt=600, n_ls=100, seed=3)
    rs, ns = engine.track_stereo_sequence(scam, cfg, mc, sc, frames)
    fc, _kf = T.stereo_features_as_c(pin(frames))
    r2, n2 = engine.pinned.empty((4,), T.POSE_RESULT_DTYPE), engine.pinned.empty((5, 2), np.int32)
    engine.wait(engine.track_stereo_sequence_async(scam, cfg, mc, sc, fc, r2, n2))
    assert r2.tobytes() == rs.tobytes()
    np.testing.assert_array_equal(n2, ns)


@pytest.mark.gpu
def test_python_handler_from_raw_features_equals_sequence_call(engine):
    """The Python mirror of the reference's main loop (initialize / insertStereoPair / optimizePose / updateFrame, frames built by
    handler.stereo_features from raw stereo features) walks a sequence frame by frame; its poses equal the batched sequence call
    chained on the host."""
    from stvo_pl_b200 import handler as Hd
    NF = 5
    frames, rel, cam = SS.make_stereo_sequence(NF, n_pt=700, n_ls=140, seed=12)
    cfg = T.kitti_config()
    h = Hd.StereoFrameHandler(cam, cfg, engine)
    h.initialize(Hd.stereo_features(engine, cam, SS.stereo_frames_slice(frames, 0, 1)))
    poses = []
    for k in range(1, NF):
        h.insertStereoPair(Hd.stereo_features(engine, cam, SS.stereo_frames_slice(frames, k, k + 1)), k)
        h.optimizePose()
        poses.append((h.curr_frame.DT.copy(), h.curr_frame.Tfw.copy(), h.n_inliers))
        h.needNewKF()
        h.updateFrame()
    res, _ = engine.track_stereo_sequence(cam, cfg, T.default_stereo_match_config(), T.default_stereo_config(), frames)
    chained = Hd.chain_poses(res.copy())
    for k in range(NF - 1):
        np.testing.assert_allclose(poses[k][0], res["DT"][k], atol=1e-12)
        np.testing.assert_allclose(poses[k][1], chained["Tfw"][k], atol=1e-9)
        assert poses[k][2] == res["n_inliers"][k]
