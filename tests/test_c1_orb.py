"""BASELINE config C1 ("KITTI plumbing, points only") realised as SURVEY 8(c) describes it: KITTI-00 calibration, synthetic
textured 1241x376 stereo images of a scene with known depth and known camera motion, key points and descriptors from OpenCV's own
ORB (the detector call of src/stereoFrame.cpp:112-115 with config_kitti.yaml:54-61), then the stereo step, f2fTracking and
optimizePose.  CPU: the oracle chain recovers the motion.  GPU: plstvo_track_stereo_batch equals the oracle chain on these real
descriptors (low-entropy neighbourhoods, repeated structure: nothing like the uniform random descriptors of the other tests)."""
import numpy as np
import pytest

from stvo_pl_b200 import types as T

cv2 = pytest.importorskip("cv2")

Z0, TX, TZ = 12.0, 0.25, 0.40   # fronto-parallel textured plane at 12 m; the camera moves 0.25 m right and 0.40 m forward


def _texture(W, H, seed):
    rng = np.random.default_rng(seed)
    img = (rng.random((H // 6 + 3, (W + 400) // 6 + 3)) * 255).astype(np.uint8)      # blocky noise -> corners at every scale
    img = cv2.resize(img, (W + 400, H), interpolation=cv2.INTER_NEAREST)
    fine = (rng.random((H, W + 400)) * 60).astype(np.uint8)
    return cv2.GaussianBlur(cv2.add(img // 2 + 40, fine), (5, 5), 1.0)


def _view(tex, cam, tx, tz, side):
    """Image of the textured plane from a camera displaced by (tx, 0, tz); side = 0 left, 1 right (baseline b further right).
    Texture pixel (x, y) is the plane point seen at image pixel (x - 200, y) by the first left camera."""
    s = Z0 / (Z0 - tz)
    M = np.float32([[s, 0, cam.cx - s * (200.0 + cam.cx) - cam.fx * (tx + side * cam.b) / (Z0 - tz)], [0, s, cam.cy - s * cam.cy]])
    return cv2.warpAffine(tex, M, (cam.width, cam.height), flags=cv2.INTER_LINEAR)


def make_c1_features(seed=0):
    cam = T.kitti_camera()
    W, H = cam.width, cam.height
    tex = _texture(W, H, seed)
    orb = cv2.ORB_create(2000, 1.2, 1, 19, 0, 2, cv2.ORB_FAST_SCORE, 31, 20)
    frames = []
    for k in range(2):
        left, right = _view(tex, cam, k * TX, k * TZ, 0), _view(tex, cam, k * TX, k * TZ, 1)
        feats = {}
        for side, img in (("l", left), ("r", right)):
            kp, des = orb.detectAndCompute(img, None)
            feats["kp_" + side] = np.array([p.pt for p in kp], np.float32).reshape(-1, 2)
            feats["desc_" + side] = np.asarray(des, np.uint8).reshape(-1, 32)
            if side == "l":
                feats["oct_l"] = np.array([p.octave for p in kp], np.int32)
        frames.append(feats)

    def as_stereo(f):
        z32, zf = np.zeros((0, 32), np.uint8), np.zeros((0, 4), np.float32)
        return dict(pl_off=[0, len(f["kp_l"])], pr_off=[0, len(f["kp_r"])], kp_l=f["kp_l"], kp_r=f["kp_r"], poct_l=f["oct_l"],
                    pdesc_l=f["desc_l"], pdesc_r=f["desc_r"], ll_off=[0, 0], lr_off=[0, 0], seg_l=zf, seg_r=zf,
                    angle_l=np.zeros(0, np.float32), loct_l=np.zeros(0, np.int32), ldesc_l=z32, ldesc_r=z32)
    return as_stereo(frames[0]), as_stereo(frames[1]), cam


def _configs():
    mc, sc, cfg = T.default_stereo_match_config(), T.default_stereo_config(), T.kitti_config()
    mc.min_ratio_12_p = 0.75                      # config_kitti.yaml:19
    cfg.has_lines = 0                             # C1: points only
    return mc, sc, cfg


def _oracle_chain(oracle, cam, mc, sc, cfg, prev, curr):
    fbs = []
    for d in (prev, curr):
        _, k, r = oracle.match_stereo_points(cam, mc, sc, d["kp_l"], d["poct_l"], d["pdesc_l"], d["kp_r"], d["pdesc_r"])
        z3, z2 = np.zeros((0, 3)), np.zeros((0, 2))
        fbs.append((k, T.FrameBatch(pt_off=[0, k], ls_off=[0, 0], pdesc=r["desc"], ldesc=np.zeros((0, 32), np.uint8), pt_P=r["P"],
                                    pt_pl=r["pl"], pt_sigma2=r["sigma2"], ls_sP=z3, ls_eP=z3, ls_le=z3, ls_spl=z2, ls_epl=z2,
                                    ls_sigma2=np.zeros(0), ls_level=np.zeros(0, np.int32))))
    res = oracle.track_batch(cam, cfg, fbs[0][1], fbs[1][1])["results"][0]
    return res, (fbs[0][0], fbs[1][0]), fbs


def test_c1_oracle_chain_recovers_the_motion(oracle):
    prev, curr, cam = make_c1_features()
    mc, sc, cfg = _configs()
    res, (k0, k1), fbs = _oracle_chain(oracle, cam, mc, sc, cfg, prev, curr)
    assert len(prev["kp_l"]) > 1500 and k0 > 400 and k1 > 400              # ORB found features, the stereo step kept many
    Z = fbs[0][1].pt_P[:, 2]
    assert abs(np.median(Z) - Z0) < 0.3                                     # lifted depth = the plane's
    assert res["good"] == 1 and res["n_matched_pt"] > 200
    t = res["DT_opt"][:3, 3]                                               # prev -> curr: the scene moves by (-TX, 0, -TZ)
    assert abs(t[0] + TX) < 0.03 and abs(t[1]) < 0.03 and abs(t[2] + TZ) < 0.10
    R = res["DT_opt"][:3, :3]
    assert np.arccos(min(1.0, (np.trace(R) - 1) / 2)) < 5e-3


@pytest.mark.gpu
def test_c1_gpu_equals_oracle_chain_on_orb_features(engine, oracle):
    import ref_numpy as RN
    prev, curr, cam = make_c1_features(seed=1)
    mc, sc, cfg = _configs()
    ref, (k0, k1), _ = _oracle_chain(oracle, cam, mc, sc, cfg, prev, curr)
    res, n_st = engine.track_stereo_batch(cam, cfg, mc, sc, prev, curr)
    assert (n_st[0, 0], n_st[0, 2]) == (k0, k1)
    assert res["status"][0] == ref["status"] and res["good"][0] == ref["good"] == 1
    assert res["n_matched_pt"][0] == ref["n_matched_pt"] and res["n_inliers"][0] == ref["n_inliers"]
    ang, tr = RN.pose_error(res["DT"][0], ref["DT"])
    assert ang < 1e-9 and tr < 1e-8
