"""Synthetic left/right stereo feature sets for StVO::matchGrid (the stereo step, src/stereoFrame.cpp:120-173, :309-398):
keypoints / segments in the left image, their right-image counterparts shifted left by a disparity, descriptors that differ
in a few bits, plus clutter.  Grid quantities are formed exactly like the caller does (inv_width = 64 / W, inv_height = 48 / H,
integer truncation for the query cells, doubles for the Bresenham rasterisation of the train segments)."""
from __future__ import annotations

import numpy as np

from .types import GRID_COLS, GRID_ROWS


def make_stereo_points(n_l, n_r, W=1241, H=376, seed=0, overlap=0.8, bitflip=0.08, max_disp=120.0, tie_stress=False):
    rng = np.random.default_rng(seed)
    inv_w, inv_h = GRID_COLS / W, GRID_ROWS / H
    pl = np.stack([rng.uniform(0, W, n_l), rng.uniform(0, H, n_l)], 1)
    pr = np.stack([rng.uniform(0, W, n_r), rng.uniform(0, H, n_r)], 1)
    if tie_stress:
        d1 = (rng.integers(0, 2, (n_l, 32), dtype=np.uint8) * 255).astype(np.uint8)
        d2 = (rng.integers(0, 2, (n_r, 32), dtype=np.uint8) * 255).astype(np.uint8)
    else:
        d1 = rng.integers(0, 256, (n_l, 32), dtype=np.uint8)
        d2 = rng.integers(0, 256, (n_r, 32), dtype=np.uint8)
    k = int(min(n_l, n_r) * overlap)
    src, dst = rng.permutation(n_l)[:k], rng.permutation(n_r)[:k]
    disp = rng.uniform(1.0, max_disp, k)
    pr[dst, 0] = np.clip(pl[src, 0] - disp, 0, W - 1e-3)
    pr[dst, 1] = np.clip(pl[src, 1] + rng.normal(0, 0.3, k), 0, H - 1e-3)
    flips = (rng.random((k, 32, 8)) < bitflip)
    d2[dst] = d1[src] ^ np.packbits(flips, axis=2).reshape(k, 32)
    q_cell = np.stack([(pl[:, 0] * inv_w).astype(np.int32), (pl[:, 1] * inv_h).astype(np.int32)], 1)
    t_cell = np.stack([(pr[:, 0] * inv_w).astype(np.int32), (pr[:, 1] * inv_h).astype(np.int32)], 1)
    return q_cell, d1, t_cell, d2


def make_stereo_lines(n_l, n_r, W=1241, H=376, seed=0, overlap=0.8, bitflip=0.08, max_disp=120.0):
    rng = np.random.default_rng(seed)
    inv_w, inv_h = GRID_COLS / W, GRID_ROWS / H

    def segs(n):
        s = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
        ang, ln = rng.uniform(0, np.pi, n), rng.uniform(5, 200, n)   # includes segments shorter than one cell
        e = s + np.stack([ln * np.cos(ang), ln * np.sin(ang)], 1)
        return s, np.clip(e, 0, [W - 1e-3, H - 1e-3])
    sl, el = segs(n_l)
    sr, er = segs(n_r)
    d1 = rng.integers(0, 256, (n_l, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n_r, 32), dtype=np.uint8)
    k = int(min(n_l, n_r) * overlap)
    src, dst = rng.permutation(n_l)[:k], rng.permutation(n_r)[:k]
    disp = rng.uniform(1.0, max_disp, (k, 1))
    sr[dst] = np.clip(sl[src] - np.concatenate([disp, np.zeros((k, 1))], 1), 0, [W - 1e-3, H - 1e-3])
    er[dst] = np.clip(el[src] - np.concatenate([disp, np.zeros((k, 1))], 1), 0, [W - 1e-3, H - 1e-3])
    flips = (rng.random((k, 32, 8)) < bitflip)
    d2[dst] = d1[src] ^ np.packbits(flips, axis=2).reshape(k, 32)
    q_line = np.stack([(sl[:, 0] * inv_w).astype(np.int32), (sl[:, 1] * inv_h).astype(np.int32),
                       (el[:, 0] * inv_w).astype(np.int32), (el[:, 1] * inv_h).astype(np.int32)], 1)
    t_line = np.stack([sr[:, 0] * inv_w, sr[:, 1] * inv_h, er[:, 0] * inv_w, er[:, 1] * inv_h], 1)
    v = np.stack([(er[:, 0] - sr[:, 0]) * inv_w, (er[:, 1] - sr[:, 1]) * inv_h], 1)   # stereoFrame.cpp:331-333
    with np.errstate(all="ignore"):
        t_dir = v / np.sqrt((v * v).sum(1, keepdims=True))
    return q_line, d1, t_line, t_dir, d2


# ---- inputs of the 3-D lifting step (src/stereoFrame.cpp:149-172, :348-397) ----
def make_lift_points(n_l, n_r, seed, W=1241, H=376):
    """Left keypoints, a right set and a matchGrid-style m12 with: unmatched rows, matches violating the epipolar bound,
    disparities below / exactly at min_disp, several pyramid levels."""
    rng = np.random.default_rng(seed)
    kp_l = np.stack([rng.uniform(0, W, n_l), rng.uniform(0, H, n_l)], 1).astype(np.float32)
    kp_r = np.stack([rng.uniform(0, W, n_r), rng.uniform(0, H, n_r)], 1).astype(np.float32)
    m12 = np.full(n_l, -1, np.int32)
    k = int(min(n_l, n_r) * 0.8)
    src, dst = rng.permutation(n_l)[:k], rng.permutation(n_r)[:k]
    m12[src] = dst
    disp = rng.uniform(-2.0, 120.0, k).astype(np.float32)
    disp[: k // 10] = 1.0                                             # exactly the default min_disp
    kp_r[dst, 0] = kp_l[src, 0] - disp
    dy = rng.normal(0, 0.6, k).astype(np.float32)
    dy[k // 10: k // 5] = 0.0
    dy[k // 5: k // 4] = 1.0                                          # exactly the default max_dist_epip (when representable)
    kp_r[dst, 1] = kp_l[src, 1] + dy
    octave = rng.integers(0, 8, n_l).astype(np.int32)
    desc = rng.integers(0, 256, (n_l, 32), dtype=np.uint8)
    return kp_l, octave, desc, kp_r, m12


def make_lift_lines(n_l, n_r, seed, W=1241, H=376):
    """Inputs of the line lifting step (src/stereoFrame.cpp:348-397): segments, KeyLine angle / octave, descriptors, m12."""
    rng = np.random.default_rng(seed)

    def segs(n):
        s = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
        ang, ln = rng.uniform(0, np.pi, n), rng.uniform(5, 200, n)
        return s, s + np.stack([ln * np.cos(ang), ln * np.sin(ang)], 1)
    sl, el = segs(n_l)
    sr, er = segs(n_r)
    m12 = np.full(n_l, -1, np.int32)
    k = int(min(n_l, n_r) * 0.8)
    src, dst = rng.permutation(n_l)[:k], rng.permutation(n_r)[:k]
    m12[src] = dst
    ds, de = rng.uniform(-2.0, 120.0, k), None
    de = ds * rng.uniform(0.5, 1.5, k)                                # some pairs fail ls_min_disp_ratio
    # the right segment covers a perturbed sub-range of the left one along the same image line
    t0, t1 = rng.uniform(-0.3, 0.3, k), rng.uniform(0.7, 1.3, k)
    d = el[src] - sl[src]
    sr[dst] = sl[src] + t0[:, None] * d - np.stack([ds, np.zeros(k)], 1)
    er[dst] = sl[src] + t1[:, None] * d - np.stack([de, np.zeros(k)], 1)
    flip = rng.random(k) < 0.3                                        # right endpoints in the opposite order
    sr[dst[flip]], er[dst[flip]] = er[dst[flip]].copy(), sr[dst[flip]].copy()
    h = src[: max(k // 12, 1)]                                        # horizontal left segments
    el[h, 1] = sl[h, 1] + rng.uniform(-0.15, 0.15, len(h))
    hr = dst[max(k // 12, 1): max(k // 8, 2)]                         # exactly horizontal right segments (division by zero)
    er[hr, 1] = sr[hr, 1]
    seg_l = np.concatenate([sl, el], 1).astype(np.float32)
    seg_r = np.concatenate([sr, er], 1).astype(np.float32)
    angle = rng.uniform(-np.pi, np.pi, n_l).astype(np.float32)
    octave = rng.integers(0, 3, n_l).astype(np.int32)
    desc = rng.integers(0, 256, (n_l, 32), dtype=np.uint8)
    return seg_l, angle, octave, desc, seg_r, m12


# ---- raw stereo frames for matchStereoPoints / matchStereoLines (src/stereoFrame.cpp:120-173, :309-398) ----
def make_stereo_frame_points(n_l, n_r, W=1241, H=376, seed=0, overlap=0.8, bitflip=0.08, max_disp=120.0):
    """Left / right key points (float32 pixels), octaves and descriptors of one rectified stereo frame: true matches sit on
    (nearly) the same row with a positive disparity, the rest is clutter.  Returns kp_l, octave_l, d1, kp_r, d2."""
    rng = np.random.default_rng(seed)
    kp_l = np.stack([rng.uniform(0, W, n_l), rng.uniform(0, H, n_l)], 1)
    kp_r = np.stack([rng.uniform(0, W, n_r), rng.uniform(0, H, n_r)], 1)
    d1 = rng.integers(0, 256, (n_l, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n_r, 32), dtype=np.uint8)
    k = int(min(n_l, n_r) * overlap)
    src, dst = rng.permutation(n_l)[:k], rng.permutation(n_r)[:k]
    disp = rng.uniform(0.2, max_disp, k)
    kp_r[dst, 0] = np.clip(kp_l[src, 0] - disp, 0, W - 1e-3)
    dy = rng.normal(0, 0.4, k)
    dy[: k // 3] = 0.0                               # exactly rectified rows (the only survivors with max_dist_epip = 0)
    kp_r[dst, 1] = np.clip(kp_l[src, 1] + dy, 0, H - 1e-3)
    flips = (rng.random((k, 32, 8)) < bitflip)
    d2[dst] = d1[src] ^ np.packbits(flips, axis=2).reshape(k, 32)
    kp_l, kp_r = kp_l.astype(np.float32), kp_r.astype(np.float32)
    kp_r[dst[: k // 3], 1] = kp_l[src[: k // 3], 1]
    return kp_l, rng.integers(0, 8, n_l).astype(np.int32), d1, kp_r, d2


def make_stereo_frame_lines(n_l, n_r, W=1241, H=376, seed=0, overlap=0.8, bitflip=0.08, max_disp=120.0):
    """Left / right key lines (float32 start / end pixels), KeyLine angle / octave and descriptors of one stereo frame.
    Returns seg_l, angle_l, octave_l, d1, seg_r, d2."""
    rng = np.random.default_rng(seed)

    def segs(n):
        s = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
        ang, ln = rng.uniform(0, np.pi, n), rng.uniform(5, 200, n)
        e = s + np.stack([ln * np.cos(ang), ln * np.sin(ang)], 1)
        return s, np.clip(e, 0, [W - 1e-3, H - 1e-3])
    sl, el = segs(n_l)
    sr, er = segs(n_r)
    d1 = rng.integers(0, 256, (n_l, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n_r, 32), dtype=np.uint8)
    k = int(min(n_l, n_r) * overlap)
    src, dst = rng.permutation(n_l)[:k], rng.permutation(n_r)[:k]
    ds = rng.uniform(0.2, max_disp, k)
    de = ds * rng.uniform(0.6, 1.4, k)
    t0, t1 = rng.uniform(-0.2, 0.2, k), rng.uniform(0.8, 1.2, k)
    d = el[src] - sl[src]
    sr[dst] = np.clip(sl[src] + t0[:, None] * d - np.stack([ds, np.zeros(k)], 1), 0, [W - 1e-3, H - 1e-3])
    er[dst] = np.clip(sl[src] + t1[:, None] * d - np.stack([de, np.zeros(k)], 1), 0, [W - 1e-3, H - 1e-3])
    flips = (rng.random((k, 32, 8)) < bitflip)
    d2[dst] = d1[src] ^ np.packbits(flips, axis=2).reshape(k, 32)
    seg_l = np.concatenate([sl, el], 1).astype(np.float32)
    seg_r = np.concatenate([sr, er], 1).astype(np.float32)
    return (seg_l, rng.uniform(-np.pi, np.pi, n_l).astype(np.float32), rng.integers(0, 3, n_l).astype(np.int32), d1, seg_r, d2)


def stereo_cells_points(kp_l, kp_r, W, H):
    """The grid coordinates the caller of matchGrid forms (src/stereoFrame.cpp:47-48, :129-139): float pixel x double inverse
    cell size, truncated."""
    inv_w, inv_h = GRID_COLS / float(W), GRID_ROWS / float(H)
    cell = lambda kp: np.stack([(kp[:, 0].astype(np.float64) * inv_w).astype(np.int32),
                                (kp[:, 1].astype(np.float64) * inv_h).astype(np.int32)], 1)
    return cell(np.asarray(kp_l, np.float32).reshape(-1, 2)), cell(np.asarray(kp_r, np.float32).reshape(-1, 2))


def stereo_cells_lines(seg_l, seg_r, W, H):
    """q_line, t_line, t_dir as src/stereoFrame.cpp:318-337 forms them."""
    inv = np.array([GRID_COLS / float(W), GRID_ROWS / float(H)] * 2)
    seg_l, seg_r = np.asarray(seg_l, np.float32).reshape(-1, 4), np.asarray(seg_r, np.float32).reshape(-1, 4)
    q_line = (seg_l.astype(np.float64) * inv).astype(np.int32)
    t_line = seg_r.astype(np.float64) * inv
    v = (seg_r[:, 2:] - seg_r[:, :2]).astype(np.float64) * inv[:2]          # float - float, then x double (:332)
    with np.errstate(all="ignore"):
        t_dir = v / np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1])[:, None]
    return q_line, t_line, t_dir


# ---- raw stereo features of (prev, curr) frame pairs of one scene, for plstvo_track_stereo_batch ----
def make_stereo_pairs(B, n_pt=1200, n_ls=300, seed=0, clutter=0.15, bitflip=0.06, noise_px=0.3):
    """B independent (prev, curr) stereo frame pairs: a static 3-D scene of points and segments seen from two poses of a
    rectified KITTI-shaped stereo rig.  Returns (prev, curr, T_gt[B,4,4], cam) where prev / curr are dicts with the fields of
    PlStereoFeatures (concatenated over frames)."""
    from .synth import expmap_se3, kitti_camera, projection
    cam = kitti_camera()
    W, H = cam.width, cam.height
    rng = np.random.default_rng(seed)

    def to_right(uv, Z):
        return np.stack([uv[:, 0] - cam.b * cam.fx / Z, uv[:, 1]], 1)

    def flip(d):
        return d ^ np.packbits(rng.random((len(d), 32, 8)) < bitflip, axis=2).reshape(len(d), 32)

    def clutter_pts(n):
        return np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1), rng.integers(0, 256, (n, 32), dtype=np.uint8)

    frames = {"prev": [], "curr": []}
    Ts = []
    for _ in range(B):
        T = expmap_se3(np.concatenate([rng.normal([0.0, 0.0, -0.6], 0.05), rng.normal(0, 0.01, 3)]))   # prev -> curr
        Ts.append(T)
        # scene in the prev camera frame
        u, v, Z = rng.uniform(30, W - 30, n_pt), rng.uniform(20, H - 20, n_pt), np.exp(rng.uniform(np.log(4), np.log(40), n_pt))
        P = np.stack([(u - cam.cx) * Z / cam.fx, (v - cam.cy) * Z / cam.fy, Z], 1)
        dP = rng.integers(0, 256, (n_pt, 32), dtype=np.uint8)
        octv = rng.integers(0, 4, n_pt).astype(np.int32)
        su, sv = rng.uniform(60, W - 60, n_ls), rng.uniform(40, H - 40, n_ls)
        Zs = np.exp(rng.uniform(np.log(5), np.log(30), n_ls))
        ang, ln = rng.uniform(0.35, np.pi - 0.35, n_ls), rng.uniform(40, 120, n_ls)      # never horizontal
        eu, ev = su + ln * np.cos(ang), sv + ln * np.sin(ang) * np.sign(rng.uniform(-1, 1, n_ls))
        Ze = Zs * rng.uniform(0.97, 1.03, n_ls)
        sP = np.stack([(su - cam.cx) * Zs / cam.fx, (sv - cam.cy) * Zs / cam.fy, Zs], 1)
        eP = np.stack([(eu - cam.cx) * Ze / cam.fx, (ev - cam.cy) * Ze / cam.fy, Ze], 1)
        dL = rng.integers(0, 256, (n_ls, 32), dtype=np.uint8)
        for name, Tf in (("prev", np.eye(4)), ("curr", T)):
            Pc = P @ Tf[:3, :3].T + Tf[:3, 3]
            uvl = projection(cam, Pc) + rng.normal(0, noise_px, (n_pt, 2))
            uvr = to_right(uvl, Pc[:, 2])
            uvr[:, 0] += rng.normal(0, noise_px, n_pt)
            nc = int(clutter * n_pt)
            cl, cdl = clutter_pts(nc)
            cr, cdr = clutter_pts(nc)
            perm_l, perm_r = rng.permutation(n_pt + nc), rng.permutation(n_pt + nc)
            kp_l = np.concatenate([uvl, cl])[perm_l].astype(np.float32)
            kp_r = np.concatenate([uvr, cr])[perm_r].astype(np.float32)
            inv_l, inv_r = np.argsort(perm_l), np.argsort(perm_r)                    # rectified: identical float rows
            kp_r[inv_r[:n_pt], 1] = kp_l[inv_l[:n_pt], 1]
            sc, ec = sP @ Tf[:3, :3].T + Tf[:3, 3], eP @ Tf[:3, :3].T + Tf[:3, 3]
            sl, el = projection(cam, sc) + rng.normal(0, noise_px, (n_ls, 2)), projection(cam, ec) + rng.normal(0, noise_px, (n_ls, 2))
            sr, er = to_right(sl, sc[:, 2]), to_right(el, ec[:, 2])
            permL, permR = rng.permutation(n_ls), rng.permutation(n_ls)
            frames[name].append(dict(
                kp_l=kp_l, kp_r=kp_r, poct_l=np.concatenate([octv, rng.integers(0, 4, nc).astype(np.int32)])[perm_l],
                pdesc_l=np.concatenate([flip(dP), cdl])[perm_l], pdesc_r=np.concatenate([flip(dP), cdr])[perm_r],
                seg_l=np.concatenate([sl, el], 1)[permL].astype(np.float32), seg_r=np.concatenate([sr, er], 1)[permR].astype(np.float32),
                angle_l=rng.uniform(-np.pi, np.pi, n_ls).astype(np.float32), loct_l=np.zeros(n_ls, np.int32),
                ldesc_l=flip(dL)[permL], ldesc_r=flip(dL)[permR]))

    def cat(fl):
        out = {k: np.concatenate([f[k] for f in fl]) for k in fl[0]}
        for off, key in (("pl_off", "kp_l"), ("pr_off", "kp_r"), ("ll_off", "seg_l"), ("lr_off", "seg_r")):
            out[off] = np.concatenate([[0], np.cumsum([len(f[key]) for f in fl])]).astype(np.int32)
        return out
    return cat(frames["prev"]), cat(frames["curr"]), np.stack(Ts), cam


def make_stereo_sequence(NF, n_pt=1200, n_ls=300, seed=0, clutter=0.15, bitflip=0.06, noise_px=0.3):
    """NF consecutive stereo frames of one static scene seen from a camera moving forward (PlStereoFeatures fields, frames
    concatenated).  Returns (frames, T_rel[NF - 1, 4, 4], cam): T_rel[k] maps frame k's camera coordinates to frame k + 1's."""
    from .synth import expmap_se3, kitti_camera, projection
    cam = kitti_camera()
    W, H = cam.width, cam.height
    rng = np.random.default_rng(seed)

    def flip(d):
        return d ^ np.packbits(rng.random((len(d), 32, 8)) < bitflip, axis=2).reshape(len(d), 32)

    # a corridor of points and segments along the camera's path (frame 0 coordinates), long enough for NF frames of 0.5 m
    total = 62.0 + 0.5 * NF
    n_pool = int(n_pt * total / 28.0) + 64
    P = np.stack([rng.uniform(-40, 40, n_pool), rng.uniform(-12, 12, n_pool), rng.uniform(3.0, total, n_pool)], 1)
    dP, octv = rng.integers(0, 256, (n_pool, 32), dtype=np.uint8), rng.integers(0, 4, n_pool).astype(np.int32)
    l_pool = int(n_ls * total / 14.0) + 64
    sP = np.stack([rng.uniform(-25, 25, l_pool), rng.uniform(-8, 8, l_pool), rng.uniform(5.0, total, l_pool)], 1)
    ang, ln = rng.uniform(0.4, np.pi - 0.4, l_pool), rng.uniform(0.5, 2.5, l_pool)                # metres, never horizontal
    eP = sP + np.stack([ln * np.cos(ang), ln * np.sin(ang), rng.normal(0, 0.2, l_pool)], 1)
    dL = rng.integers(0, 256, (l_pool, 32), dtype=np.uint8)
    T_abs, rel, frames = np.eye(4), [], []
    for k in range(NF):
        if k:
            # bounded absolute pose (the camera stays inside the corridor however long the sequence is): forward 0.5 m per frame,
            # a gentle lateral weave, small independent attitude jitter
            cam_pos = np.array([0.4 * np.sin(k / 9.0), 0.08 * np.sin(k / 5.0), 0.5 * k]) + rng.normal(0, 0.01, 3)
            Rwc = expmap_se3(np.concatenate([np.zeros(3), rng.normal(0, 0.006, 3)]))[:3, :3]      # camera -> world
            T_new = np.eye(4)
            T_new[:3, :3], T_new[:3, 3] = Rwc.T, -Rwc.T @ cam_pos                               # world -> camera
            rel.append(T_new @ np.linalg.inv(T_abs))
            T_abs = T_new
        Pc = P @ T_abs[:3, :3].T + T_abs[:3, 3]
        uv = projection(cam, Pc)
        vis = (Pc[:, 2] > 2.0) & (Pc[:, 2] < 60.0) & (uv[:, 0] > 5) & (uv[:, 0] < W - 5) & (uv[:, 1] > 5) & (uv[:, 1] < H - 5)
        idx = np.nonzero(vis)[0][:n_pt]
        uvl = uv[idx] + rng.normal(0, noise_px, (len(idx), 2))
        uvr = np.stack([uvl[:, 0] - cam.b * cam.fx / Pc[idx, 2] + rng.normal(0, noise_px, len(idx)), uvl[:, 1]], 1)
        nc = int(clutter * len(idx))
        cl = np.stack([rng.uniform(0, W, nc), rng.uniform(0, H, nc)], 1)
        cr = np.stack([rng.uniform(0, W, nc), rng.uniform(0, H, nc)], 1)
        pl_, pr_ = rng.permutation(len(idx) + nc), rng.permutation(len(idx) + nc)
        kp_l = np.concatenate([uvl, cl])[pl_].astype(np.float32)
        kp_r = np.concatenate([uvr, cr])[pr_].astype(np.float32)
        kp_r[np.argsort(pr_)[:len(idx)], 1] = kp_l[np.argsort(pl_)[:len(idx)], 1]
        sc_, ec_ = sP @ T_abs[:3, :3].T + T_abs[:3, 3], eP @ T_abs[:3, :3].T + T_abs[:3, 3]
        s2, e2 = projection(cam, sc_), projection(cam, ec_)
        visl = (sc_[:, 2] > 2) & (ec_[:, 2] > 2) & (sc_[:, 2] < 45) & (np.minimum(s2[:, 0], e2[:, 0]) > 5) & (np.maximum(s2[:, 0], e2[:, 0]) < W - 5) & \
               (np.minimum(s2[:, 1], e2[:, 1]) > 5) & (np.maximum(s2[:, 1], e2[:, 1]) < H - 5) & (np.abs(s2[:, 1] - e2[:, 1]) > 8)
        li = np.nonzero(visl)[0][:n_ls]
        sl, el = s2[li] + rng.normal(0, noise_px, (len(li), 2)), e2[li] + rng.normal(0, noise_px, (len(li), 2))
        sr = np.stack([sl[:, 0] - cam.b * cam.fx / sc_[li, 2], sl[:, 1]], 1)
        er = np.stack([el[:, 0] - cam.b * cam.fx / ec_[li, 2], el[:, 1]], 1)
        qL, qR = rng.permutation(len(li)), rng.permutation(len(li))
        frames.append(dict(
            kp_l=kp_l, kp_r=kp_r, poct_l=np.concatenate([octv[idx], rng.integers(0, 4, nc).astype(np.int32)])[pl_],
            pdesc_l=np.concatenate([flip(dP[idx]), rng.integers(0, 256, (nc, 32), dtype=np.uint8)])[pl_],
            pdesc_r=np.concatenate([flip(dP[idx]), rng.integers(0, 256, (nc, 32), dtype=np.uint8)])[pr_],
            seg_l=np.concatenate([sl, el], 1)[qL].astype(np.float32), seg_r=np.concatenate([sr, er], 1)[qR].astype(np.float32),
            angle_l=rng.uniform(-np.pi, np.pi, len(li)).astype(np.float32), loct_l=np.zeros(len(li), np.int32),
            ldesc_l=flip(dL[li])[qL], ldesc_r=flip(dL[li])[qR]))
    out = {k: np.concatenate([f[k] for f in frames]) for k in frames[0]}
    for off, key in (("pl_off", "kp_l"), ("pr_off", "kp_r"), ("ll_off", "seg_l"), ("lr_off", "seg_r")):
        out[off] = np.concatenate([[0], np.cumsum([len(f[key]) for f in frames])]).astype(np.int32)
    return out, np.stack(rel) if rel else np.zeros((0, 4, 4)), cam


def stereo_frames_slice(d, lo, hi):
    """Frames [lo, hi) of a PlStereoFeatures dict as a new dict (offsets rebased)."""
    out = {}
    for off, keys in (("pl_off", ("kp_l", "poct_l", "pdesc_l")), ("pr_off", ("kp_r", "pdesc_r")),
                      ("ll_off", ("seg_l", "angle_l", "loct_l", "ldesc_l")), ("lr_off", ("seg_r", "ldesc_r"))):
        a, b = d[off][lo], d[off][hi]
        out[off] = (d[off][lo:hi + 1] - a).astype(np.int32)
        for k in keys:
            out[k] = d[k][a:b]
    return out
