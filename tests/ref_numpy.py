"""Independent numpy/scipy restatement of the pose half of the hot path, used ONLY to cross-check the C
oracle's hand-rolled algebra and control flow (tests, CPU).  It is written vectorised and uses library
algebra (numpy.linalg.solve / inv / eigvalsh, scipy.linalg.expm / logm) where the oracle restates Eigen by
hand, so an error in one is unlikely to be mirrored in the other.  Citations: src/stereoFrameHandler.cpp.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


def hat_se3(x):
    t, w = x[:3], x[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    M[:3, 3] = t
    return M


def expmap_se3(x):
    return sla.expm(hat_se3(np.asarray(x, float)))


def logmap_se3(T):
    M = np.real(sla.logm(T))
    return np.array([M[0, 3], M[1, 3], M[2, 3], M[2, 1], M[0, 2], M[1, 0]])


def inverse_se3(T):
    return np.linalg.inv(T)


def cauchy(r):
    return 1.0 / (1.0 + r * r)


def mad_stdv(res):
    """vector_stdv_mad (src/auxiliar.cpp:444-460) incl. the fabsf float rounding."""
    if len(res) == 0:
        return 0.0
    r = np.sort(res)
    med = r[len(r) // 2]
    dev = np.sort(np.abs((r - med).astype(np.float32)).astype(np.float64))
    return 1.4826 * dev[len(r) // 2]


def mean_stdv_mad(res):
    """vector_mean_stdv_mad (src/auxiliar.cpp:387-430)."""
    n = len(res)
    if n == 0:
        return 0.0, 0.0
    stdv = mad_stdv(res)
    sel = res < 2.0 * stdv
    k = int(sel.sum())
    if k >= int(0.2 * n):
        mean = res[sel].sum() / k if k else np.nan
    else:
        mean = res.mean()
    return mean, stdv


def overlap(spl, epl, sp, ep):
    """lineSegmentOverlap (src/stereoFrame.cpp:510-616), vectorised over lines."""
    l = epl - spl
    vert = np.abs(spl[:, 0] - epl[:, 0]) < 1.0
    horiz = (~vert) & (np.abs(spl[:, 1] - epl[:, 1]) < 1.0)
    a = spl[:, 1] - epl[:, 1]
    b = epl[:, 0] - spl[:, 0]
    c = spl[:, 0] * epl[:, 1] - epl[:, 0] * spl[:, 1]
    with np.errstate(all="ignore"):
        lxy = 1.0 / (a * a + b * b)
        sx = (b * (b * sp[:, 0] - a * sp[:, 1]) - a * c) * lxy
        ex = (b * (b * ep[:, 0] - a * ep[:, 1]) - a * c) * lxy
        ls_g, le_g = (sx - spl[:, 0]) / l[:, 0], (ex - spl[:, 0]) / l[:, 0]
        ls_v, le_v = (sp[:, 1] - spl[:, 1]) / l[:, 1], (ep[:, 1] - spl[:, 1]) / l[:, 1]
        ls_h, le_h = (sp[:, 0] - spl[:, 0]) / l[:, 0], (ep[:, 0] - spl[:, 0]) / l[:, 0]
    ls = np.where(vert, ls_v, np.where(horiz, ls_h, ls_g))
    le = np.where(vert, le_v, np.where(horiz, le_h, le_g))
    lo, hi = np.minimum(ls, le), np.maximum(ls, le)
    out = hi - lo
    out = np.where(hi > 1.0, 1.0 - lo, out)
    out = np.where(lo < 0.0, hi, out)
    out = np.where((hi < 0.0) | (lo > 1.0), 0.0, out)
    out = np.where((lo < 0.0) & (hi > 1.0), 1.0, out)
    return out


def _jac(fgz2, g, dx, dy):
    gx, gy, gz = g[:, 0], g[:, 1], g[:, 2]
    return np.stack([fgz2 * dx * gz, fgz2 * dy * gz, -fgz2 * (gx * dx + gy * dy),
                     -fgz2 * (gx * gy * dx + gy * gy * dy + gz * gz * dy),
                     fgz2 * (gx * gx * dx + gz * gz * dx + gx * gy * dy),
                     fgz2 * (gx * gz * dy - gy * gz * dx)], axis=1)


class Problem:
    """One matched_pt / matched_ls problem."""

    def __init__(self, cam, cfg, P, obs, s2p, sP, eP, le, spl, epl, s2l):
        self.cam, self.cfg = cam, cfg
        self.P, self.obs, self.s2p = P, obs, s2p
        self.sP, self.eP, self.le, self.spl, self.epl, self.s2l = sP, eP, le, spl, epl, s2l
        self.inl_p = np.ones(len(P), bool)
        self.inl_l = np.ones(len(sP), bool)
        self.evals = 0

    def proj(self, X):
        c = self.cam
        return np.stack([c.cx + c.fx * X[:, 0] / X[:, 2], c.cy + c.fy * X[:, 1] / X[:, 2]], axis=1)

    def point_res(self, DT):
        X = self.P @ DT[:3, :3].T + DT[:3, 3]
        e = self.proj(X) - self.obs
        return X, e, np.hypot(e[:, 0], e[:, 1])

    def line_res(self, DT):
        Xs = self.sP @ DT[:3, :3].T + DT[:3, 3]
        Xe = self.eP @ DT[:3, :3].T + DT[:3, 3]
        ps, pe = self.proj(Xs), self.proj(Xe)
        ds = self.le[:, 0] * ps[:, 0] + self.le[:, 1] * ps[:, 1] + self.le[:, 2]
        de = self.le[:, 0] * pe[:, 0] + self.le[:, 1] * pe[:, 1] + self.le[:, 2]
        return Xs, Xe, ps, pe, ds, de, np.hypot(ds, de)

    def evaluate(self, DT, robust=False):
        """optimizeFunctions :549-694 / optimizeFunctionsRobust :696-962."""
        self.evals += 1
        th, fx = self.cfg.homog_th, self.cam.fx
        H, g, e, N = np.zeros((6, 6)), np.zeros(6), 0.0, 0
        s_p = s_l = 1.0
        X, err, n = self.point_res(DT)
        Xs, Xe, ps, pe, ds, de, nl = self.line_res(DT)
        if robust:
            s_p = min(max(mad_stdv(n[self.inl_p]), 1e-4), np.sqrt(7.815))
            s_l = min(max(mad_stdv(nl[self.inl_l]), 1e-4), np.sqrt(7.815))
        if len(self.P):
            J = _jac(fx / np.maximum(th, X[:, 2] ** 2), X, err[:, 0], err[:, 1]) / np.maximum(th, n)[:, None]
            r = n if robust else n * np.sqrt(self.s2p)
            w = cauchy(r / s_p) if robust else cauchy(r)
            m = self.inl_p
            H += (J[m] * w[m, None]).T @ J[m]
            g += (J[m] * (r[m] * w[m])[:, None]).sum(0)
            e += (r[m] ** 2 * w[m]).sum()
            N += int(m.sum())
        if len(self.sP):
            lx, ly = self.le[:, 0], self.le[:, 1]
            Js = _jac(fx / np.maximum(th, Xs[:, 2] ** 2), Xs, lx, ly)
            Je = _jac(fx / np.maximum(th, Xe[:, 2] ** 2), Xe, lx, ly)
            J = (Js * ds[:, None] + Je * de[:, None]) / np.maximum(th, nl)[:, None]
            r = nl if robust else nl * np.sqrt(self.s2l)
            w = (cauchy(r / s_l) if robust else cauchy(r)) * overlap(self.spl, self.epl, ps, pe)
            m = self.inl_l
            H += (J[m] * w[m, None]).T @ J[m]
            g += (J[m] * (r[m] * w[m])[:, None]).sum(0)
            e += (r[m] ** 2 * w[m]).sum()
            N += int(m.sum())
        return H, g, e / N

    def gn(self, DT, max_iters):
        """gaussNewtonOptimization :394-431."""
        cfg = self.cfg
        err_prev = 999999999.9
        self.evals = 0
        H = np.zeros((6, 6))
        err = 0.0
        for it in range(max_iters):
            H, g, err = self.evaluate(DT)
            if err > err_prev:
                if it > 0:
                    break
                return DT, None, -1.0
            if err < cfg.min_error or abs(err - err_prev) < cfg.min_error_change:
                break
            inc = np.linalg.solve(H, g)
            DT = DT @ inverse_se3(expmap_se3(inc))
            if np.linalg.norm(inc[:3]) < cfg.min_error_change and np.linalg.norm(inc[3:]) < cfg.min_error_change:
                break
            err_prev = err
        return DT, np.linalg.inv(H), err

    def gnr(self, DT, max_iters):
        """gaussNewtonOptimizationRobust :433-480."""
        cfg = self.cfg
        DT0 = DT.copy()
        err_prev = 999999999.9
        self.evals = 0
        good = True
        H = np.zeros((6, 6))
        err = 0.0
        for it in range(max_iters):
            H, g, err = self.evaluate(DT, robust=True)
            if abs(err - err_prev) < cfg.min_error_change or err < cfg.min_error:
                break
            inc = np.linalg.solve(H, g)
            if np.linalg.slogdet(H)[1] < 0.0:
                good = False
                break
            DT = DT @ inverse_se3(expmap_se3(inc))
            if np.linalg.norm(inc) < cfg.min_error_change:
                break
            err_prev = err
        if good:
            return DT, np.linalg.inv(H), err
        return DT0, np.eye(6), -1.0

    def good_solution(self, DT, cov, err):
        """isGoodSolution :292-305 (lower triangle, like SelfAdjointEigenSolver)."""
        if cov is None:
            return False
        w = np.linalg.eigvalsh(cov, UPLO="L")
        return not (w[0] < 0.0 or w[-1] > 1.0 or err < 0.0 or err > 1.0 or not np.isfinite(DT).all())

    def remove_outliers(self, DT):
        """removeOutliers :988-1067."""
        if self.cfg.has_points and len(self.P):
            _, _, n = self.point_res(DT)
            res = n * np.sqrt(self.s2p)
            mean, stdv = mean_stdv_mad(res)
            self.inl_p &= ~(np.abs(res - mean) > self.cfg.inlier_k * stdv)
        if self.cfg.has_lines and len(self.sP):
            nl = self.line_res(DT)[-1]
            res = nl * np.sqrt(self.s2l)
            mean, stdv = mean_stdv_mad(res)
            self.inl_l &= ~(np.abs(res - mean) > self.cfg.inlier_k * stdv)

    def optimize_pose(self):
        """optimizePose :307-392 without motion model.  Returns dict like PlPoseResult."""
        cfg = self.cfg
        DT = np.eye(4)
        cov, err = None, -1.0
        status, it1, it2 = 0, 0, 0
        n_inl = int(self.inl_p.sum() + self.inl_l.sum())
        run = self.gn if cfg.solver_mode == 0 else self.gnr
        if n_inl >= cfg.min_features:
            DT_, cov, err = run(DT.copy(), cfg.max_iters)
            it1 = self.evals
            if self.good_solution(DT_, cov, err):
                self.remove_outliers(DT_)
                if int(self.inl_p.sum() + self.inl_l.sum()) >= cfg.min_features:
                    DT, cov, err = run(DT.copy(), cfg.max_iters_ref)
                    it2 = self.evals
                else:
                    DT, status = np.eye(4), 3
            else:
                DT, cov, err = self.gnr(DT.copy(), cfg.max_iters_ref)
                it2, status = self.evals, 1
        else:
            status = 2
        good = self.good_solution(DT, cov, err) and not np.array_equal(DT, np.eye(4))
        out = dict(DT_opt=DT, status=status, iters_stage1=it1, iters_stage2=it2, good=int(good),
                   inl_p=self.inl_p.copy(), inl_l=self.inl_l.copy())
        if good:
            out.update(DT=expmap_se3(logmap_se3(inverse_se3(DT))), DT_cov=cov, err_norm=err,
                       DT_cov_eig=np.linalg.eigvalsh(cov, UPLO="L"))
        else:
            out.update(DT=np.eye(4), DT_cov=np.zeros((6, 6)), err_norm=-1.0, DT_cov_eig=np.zeros(6))
        return out


def problem_from_matched(cam, cfg, m, p) -> Problem:
    a, b = m.pt_off[p], m.pt_off[p + 1]
    c, d = m.ls_off[p], m.ls_off[p + 1]
    return Problem(cam, cfg, m.pt_P[a:b], m.pt_pl_obs[a:b], m.pt_sigma2[a:b], m.ls_sP[c:d], m.ls_eP[c:d],
                   m.ls_le_obs[c:d], m.ls_spl[c:d], m.ls_epl[c:d], m.ls_sigma2[c:d])


def pose_error(T_a, T_b):
    """(rotation angle [rad] of T_a T_b^-1, translation difference [m])."""
    D = T_a @ np.linalg.inv(T_b)
    c = np.clip((np.trace(D[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)
    ang = np.arccos(c)
    if ang < 1e-6:  # acos loses precision near 1: use the skew part
        S = 0.5 * (D[:3, :3] - D[:3, :3].T)
        ang = np.linalg.norm([S[2, 1], S[0, 2], S[1, 0]])
    return float(ang), float(np.linalg.norm(T_a[:3, 3] - T_b[:3, 3]))
