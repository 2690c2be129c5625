"""Generates the committed golden fixtures in tests/golden/ (run in the dev container; needs cv2).

  match_*.npz : inputs + the output of OpenCV's cv::BFMatcher(NORM_HAMMING,false)::knnMatch(.,.,2) — the
                third-party routine the reference calls at src/matching.cpp:47-48 — taken from the
                opencv-python build in this image, with src/matching.cpp:50-61 (ratio test in float) and
                :76-86 (mutual filter) applied on top in float32 numpy.  These PIN the matching half.
  pose_*.npz  : small seeded problems with the pose results of the independent numpy/scipy restatement
                (tests/ref_numpy.py) and of the C oracle.  The reference has no tests or fixtures for the
                pose half and cannot be compiled here, so these are regression vectors, not reference
                outputs (pose parity is "unpinned").

Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import cv2  # noqa: E402

from stvo_pl_b200 import synth, types as T  # noqa: E402


def cv2_knn2(d1, d2):
    bfm = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    ms = bfm.knnMatch(d1, d2, 2)
    idx = np.full((len(d1), 2), -1, np.int32)
    dist = np.full((len(d1), 2), -1.0, np.float32)
    for i, m in enumerate(ms):
        for k, dm in enumerate(m):
            idx[i, k], dist[i, k] = dm.trainIdx, dm.distance
    return idx, dist


def match_nnr_from_knn(idx, dist, nnr):
    """src/matching.cpp:50-61 on top of the knnMatch output (float compare)."""
    nnr = np.float32(nnr)
    ok = (idx[:, 1] >= 0) & (dist[:, 0] < (dist[:, 1] * nnr).astype(np.float32))
    return np.where(ok, idx[:, 0], -1).astype(np.int32)


def match_from_cv2(d1, d2, nnr):
    i12, s12 = cv2_knn2(d1, d2)
    i21, s21 = cv2_knn2(d2, d1)
    m12 = match_nnr_from_knn(i12, s12, nnr)
    m21 = match_nnr_from_knn(i21, s21, nnr)
    mutual = m12.copy()
    for i1, i2 in enumerate(m12):  # src/matching.cpp:80-86
        if i2 >= 0 and m21[i2] != i1:
            mutual[i1] = -1
    return i12, s12, m12, mutual


def match_cases():
    rng = np.random.default_rng(7)
    cases = {}
    cases["random_300x400"] = (rng.integers(0, 256, (300, 32), dtype=np.uint8),
                               rng.integers(0, 256, (400, 32), dtype=np.uint8), 0.9)
    # low entropy: every byte 0x00 or 0xFF -> distances are multiples of 8, heavy ties
    cases["ties_300x400"] = ((rng.integers(0, 2, (300, 32), dtype=np.uint8) * 255).astype(np.uint8),
                             (rng.integers(0, 2, (400, 32), dtype=np.uint8) * 255).astype(np.uint8), 0.9)
    # only 2 bytes vary: many exact duplicates, distance-0 ties
    d = np.zeros((200, 32), np.uint8)
    d[:, :2] = rng.integers(0, 4, (200, 2), dtype=np.uint8)
    e = np.zeros((180, 32), np.uint8)
    e[:, :2] = rng.integers(0, 4, (180, 2), dtype=np.uint8)
    cases["duplicates_200x180"] = (d, e, 0.75)
    cases["n2_is_2"] = (rng.integers(0, 256, (40, 32), dtype=np.uint8), rng.integers(0, 256, (2, 32), dtype=np.uint8), 0.9)
    cases["n2_is_3"] = (rng.integers(0, 256, (40, 32), dtype=np.uint8), rng.integers(0, 256, (3, 32), dtype=np.uint8), 0.75)
    cases["n1_is_1"] = (rng.integers(0, 256, (1, 32), dtype=np.uint8), rng.integers(0, 256, (50, 32), dtype=np.uint8), 0.9)
    prev, curr, _, _ = synth.make_batch("kitti", 1, n_pt=600, n_ls=500)
    cases["synth_lines_500x500"] = (prev.ldesc, curr.ldesc, 0.75)
    cases["synth_points_600x600"] = (prev.pdesc, curr.pdesc, 0.75)
    prev, curr, _, _ = synth.make_batch("kitti", 1, n_pt=500, n_ls=100, tie_stress=True)
    cases["synth_tiestress_500x500"] = (prev.pdesc, curr.pdesc, 0.9)
    return cases


def main():
    for name, (d1, d2, nnr) in match_cases().items():
        d1, d2 = np.ascontiguousarray(d1), np.ascontiguousarray(d2)
        idx, dist, m12, mutual = match_from_cv2(d1, d2, nnr)
        np.savez_compressed(os.path.join(HERE, f"match_{name}.npz"), d1=d1, d2=d2, nnr=np.float64(nnr),
                            knn_idx=idx, knn_dist=dist, m12_nnr=m12, m12_mutual=mutual,
                            cv2_version=cv2.__version__)
        print(name, "accepted", int((m12 >= 0).sum()), "mutual", int((mutual >= 0).sum()))

    # ---- pose regression vectors ----
    from oracle.oracle import Oracle
    import ref_numpy as R
    orc = Oracle()
    for shape, cfgf, npt, nls, B in [("kitti", T.kitti_config, 300, 80, 3), ("euroc", T.euroc_config, 200, 60, 2)]:
        prev, curr, Tgt, cam = synth.make_batch(shape, B, n_pt=npt, n_ls=nls)
        cfg = cfgf()
        o = orc.track_batch(cam, cfg, prev, curr)
        matched = T.matched_from_frames(prev, curr, o["m12_pt"], o["m12_ls"], cfg.lsd_scale)
        ref = [R.problem_from_matched(cam, cfg, matched, p).optimize_pose() for p in range(B)]
        np.savez_compressed(
            os.path.join(HERE, f"pose_{shape}.npz"), shape=shape, B=B, n_pt=npt, n_ls=nls,
            T_gt=Tgt, m12_pt=o["m12_pt"], m12_ls=o["m12_ls"],
            oracle_DT=o["results"]["DT"], oracle_DT_opt=o["results"]["DT_opt"], oracle_err=o["results"]["err_norm"],
            oracle_cov=o["results"]["DT_cov"], oracle_n_inl=np.stack([o["results"]["n_inliers_pt"], o["results"]["n_inliers_ls"]], 1),
            oracle_inlier_pt=o["inlier_pt"], oracle_inlier_ls=o["inlier_ls"],
            numpy_DT=np.stack([r["DT"] for r in ref]), numpy_err=np.array([r["err_norm"] for r in ref]))
        for p in range(B):
            print(shape, p, R.pose_error(o["results"]["DT"][p], ref[p]["DT"]), o["results"]["err_norm"][p], ref[p]["err_norm"])


if __name__ == "__main__":
    main()
