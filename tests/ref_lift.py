"""Independent Python restatement of the 3-D lifting of the stereo matches, written from the reference text
(src/stereoFrame.cpp:149-172 points, :348-397 lines, :405-415, :473-508; src/pinholeStereoCamera.cpp:221-229;
src/stereoFeatures.cpp:41-47, :107-115).  Scalar loops with numpy float32 / Python float (= IEEE double) arithmetic in the
reference's own order, so the C oracle can be held to exact equality.  Test infrastructure only."""
import math

import numpy as np


def sigma2_of_level(level, scale):
    s = 1.0
    for _ in range(int(level)):
        s *= scale
    return 1.0 / (s * s)


def back_projection(cam, u, v, disp):
    bd = cam.b / disp
    return [bd * (u - cam.cx), bd * (v - cam.cy), bd * cam.fx]


def overlap_stereo(sc, spl_obs, epl_obs, spl_proj, epl_proj):
    overlap = 1.0
    if abs(epl_obs - spl_obs) > sc.line_horiz_th:
        sln, eln = min(spl_obs, epl_obs), max(spl_obs, epl_obs)
        spn, epn = min(spl_proj, epl_proj), max(spl_proj, epl_proj)
        length = eln - spn
        if epn < sln or spn > eln:
            overlap = 0.0
        elif epn > eln and spn < sln:
            overlap = eln - sln
        else:
            overlap = min(eln, epn) - max(sln, spn)
        overlap = overlap / length if length > float(np.float32(0.01)) else 0.0
        overlap = min(overlap, 1.0)
    return overlap


def lift_points(cam, sc, kp_l, octave_l, desc_l, kp_r, m12):
    kp_l, kp_r = np.asarray(kp_l, np.float32).reshape(-1, 2), np.asarray(kp_r, np.float32).reshape(-1, 2)
    out = dict(pl=[], disp=[], P=[], sigma2=[], level=[], desc=[], src_idx=[])
    for i1, i2 in enumerate(m12):
        if i2 < 0:
            continue
        if float(abs(np.float32(kp_l[i1, 1] - kp_r[i2, 1]))) <= sc.max_dist_epip:      # float - float, fabsf
            disp = float(np.float32(kp_l[i1, 0] - kp_r[i2, 0]))                         # float - float -> double
            if disp >= sc.min_disp:
                u, v = float(kp_l[i1, 0]), float(kp_l[i1, 1])
                out["pl"].append([u, v]); out["disp"].append(disp); out["P"].append(back_projection(cam, u, v, disp))
                out["sigma2"].append(sigma2_of_level(octave_l[i1], sc.orb_scale_factor)); out["level"].append(int(octave_l[i1]))
                out["desc"].append(np.asarray(desc_l[i1])); out["src_idx"].append(i1)
    return out


def lift_lines(cam, sc, seg_l, angle_l, octave_l, desc_l, seg_r, m12):
    seg_l, seg_r = np.asarray(seg_l, np.float32).reshape(-1, 4), np.asarray(seg_r, np.float32).reshape(-1, 4)
    keys = ("spl", "epl", "sdisp", "edisp", "sP", "eP", "le", "angle", "sigma2", "level", "desc", "src_idx")
    out = {k: [] for k in keys}
    with np.errstate(all="ignore"):
        for i1, i2 in enumerate(m12):
            if i2 < 0:
                continue
            spl, epl = [float(seg_l[i1, 0]), float(seg_l[i1, 1])], [float(seg_l[i1, 2]), float(seg_l[i1, 3])]
            le = np.array([spl[1] - epl[1], epl[0] - spl[0], spl[0] * epl[1] - spl[1] * epl[0]])   # (spl,1) x (epl,1)
            le = le / np.float64(math.sqrt(le[0] * le[0] + le[1] * le[1]))
            spr, epr = [float(seg_r[i2, 0]), float(seg_r[i2, 1])], [float(seg_r[i2, 2]), float(seg_r[i2, 3])]
            overlap = overlap_stereo(sc, spl[1], epl[1], spr[1], epr[1])
            # :366-367, with the in-place overwrite of sp_r
            den = np.float64(spr[1] - epr[1])
            spr[0] = float((np.float64(spr[0] * (spl[1] - epr[1])) + np.float64(epr[0] * (spr[1] - spl[1]))) / den)
            spr[1] = spl[1]
            den = np.float64(spr[1] - epr[1])
            epr[0] = float((np.float64(spr[0] * (epl[1] - epr[1])) + np.float64(epr[0] * (spr[1] - epl[1]))) / den)
            epr[1] = epl[1]
            ds, de = spl[0] - spr[0], epl[0] - epr[0]
            mn = de if de < ds else ds          # std::min / std::max comparison order (matters for NaN only)
            mx = de if ds < de else ds
            if float(np.float64(mn) / np.float64(mx)) < sc.ls_min_disp_ratio:
                ds, de = -1.0, -1.0
            if (ds >= sc.min_disp and de >= sc.min_disp and abs(spl[1] - epl[1]) > sc.line_horiz_th
                    and abs(spr[1] - epr[1]) > sc.line_horiz_th and overlap > sc.stereo_overlap_th):
                out["spl"].append(spl); out["epl"].append(epl); out["sdisp"].append(ds); out["edisp"].append(de)
                out["sP"].append(back_projection(cam, spl[0], spl[1], ds)); out["eP"].append(back_projection(cam, epl[0], epl[1], de))
                out["le"].append(le.tolist()); out["angle"].append(float(angle_l[i1]))
                out["sigma2"].append(sigma2_of_level(octave_l[i1], sc.lsd_scale)); out["level"].append(int(octave_l[i1]))
                out["desc"].append(np.asarray(desc_l[i1])); out["src_idx"].append(i1)
    return out
