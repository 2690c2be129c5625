"""Known-answer tests derivable from the reference source alone plus library
cross-checks (scipy expm/logm, numpy.linalg) of the oracle's restated Eigen pieces."""
import numpy as np
import scipy.linalg as sla

import ref_numpy as R
from stvo_pl_b200 import types as T


def test_cauchy(oracle):  # src/auxiliar.cpp:556-559
    assert oracle.robust_weight_cauchy(0.0) == 1.0
    assert oracle.robust_weight_cauchy(1.0) == 0.5


def test_expmap_logmap(oracle):  # src/auxiliar.cpp:113-173
    np.testing.assert_array_equal(oracle.expmap_se3(np.zeros(6)), np.eye(4))
    rng = np.random.default_rng(3)
    for scale in (1e-8, 1e-3, 0.3, 2.5):
        for _ in range(10):
            x = rng.normal(0, 1, 6) * np.array([1, 1, 1, scale, scale, scale])
            if np.linalg.norm(x[3:]) >= np.pi:
                continue
            Tm = oracle.expmap_se3(x)
            if np.linalg.norm(x[3:]) >= 1e-6:
                np.testing.assert_allclose(Tm, R.expmap_se3(x), atol=1e-12)
                np.testing.assert_allclose(oracle.logmap_se3(Tm), x, atol=1e-9)
            np.testing.assert_allclose(oracle.inverse_se3(Tm) @ Tm, np.eye(4), atol=1e-12)


def test_expmap_small_angle_branch(oracle):
    """theta < 1e-6 -> R = I and t is used as is (src/auxiliar.cpp:131-133)."""
    x = np.array([0.1, -0.2, 0.3, 3e-7, 0, 0])
    Tm = oracle.expmap_se3(x)
    np.testing.assert_array_equal(Tm[:3, :3], np.eye(3))
    np.testing.assert_array_equal(Tm[:3, 3], x[:3])


def test_adjoint_unccomp(oracle):  # src/auxiliar.cpp:175-197
    rng = np.random.default_rng(4)
    Tm = R.expmap_se3(rng.normal(0, 0.5, 6))
    Ad = oracle.adjoint_se3(Tm)
    Rm, t = Tm[:3, :3], Tm[:3, 3]
    S = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    exp = np.zeros((6, 6))
    exp[:3, :3] = Rm
    exp[:3, 3:] = S @ Rm
    exp[3:, 3:] = Rm
    np.testing.assert_allclose(Ad, exp, atol=1e-14)
    c1, ci = np.eye(6) * 0.5, np.diag(np.arange(1, 7) * 1e-3)
    np.testing.assert_allclose(oracle.unccomp_se3(Tm, c1, ci), c1 + exp @ ci @ exp.T, atol=1e-14)


def test_projection_roundtrip(oracle):  # src/pinholeStereoCamera.cpp:221-237
    cam = T.kitti_camera()
    for u, v, d in [(10.0, 20.0, 3.0), (600.5, 180.25, 45.0), (1240.0, 375.0, 1.0)]:
        P = oracle.back_projection(cam, u, v, d)
        np.testing.assert_allclose(oracle.projection(cam, P), [u, v], atol=1e-10)
        assert abs(P[2] - cam.b * cam.fx / d) < 1e-12


def test_mad(oracle):  # src/auxiliar.cpp:387-460
    assert oracle.vector_stdv_mad([1, 2, 3, 4, 100]) == 1.4826 * 1.0
    assert oracle.vector_stdv_mad([]) == 0.0
    m, s = oracle.vector_mean_stdv_mad([])
    assert m == 0.0 and s == 0.0
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 100, 1001):
        res = np.abs(rng.normal(0, 1, n)) + (rng.random(n) < 0.1) * 20
        m, s = oracle.vector_mean_stdv_mad(res)
        m2, s2 = R.mean_stdv_mad(res)
        assert s == s2
        assert (np.isnan(m) and np.isnan(m2)) or abs(m - m2) <= 1e-12 * max(1, abs(m2))
        assert oracle.vector_stdv_mad(res) == R.mad_stdv(res)


def test_mad_uses_float_rounding(oracle):
    """fabsf: the deviation is rounded to float (src/auxiliar.cpp:400)."""
    res = np.array([0.0, 1.0 + 1e-12, 5.0])
    s = oracle.vector_stdv_mad(res)          # median = 1+1e-12, deviations {1+1e-12, 0, 4-1e-12} -> float
    assert s == 1.4826 * float(np.float32(1.0 + 1e-12))


def test_overlap_known_answers(oracle):  # src/stereoFrame.cpp:510-616
    s, e = [100.0, 50.0], [160.0, 130.0]
    assert abs(oracle.line_segment_overlap(s, e, s, e) - 1.0) < 1e-12           # with itself
    d = np.array(e) - np.array(s)
    assert oracle.line_segment_overlap(s, e, np.array(e) + 0.5 * d, np.array(e) + 1.5 * d) == 0.0  # disjoint collinear
    assert abs(oracle.line_segment_overlap(s, e, np.array(s) + 0.25 * d, np.array(s) + 0.75 * d) - 0.5) < 1e-12
    assert oracle.line_segment_overlap(s, e, np.array(s) - d, np.array(e) + d) == 1.0  # covers
    # vertical branch |dx| < 1, horizontal branch |dy| < 1
    assert abs(oracle.line_segment_overlap([10, 0], [10.5, 100], [300, 25], [-7, 75]) - 0.5) < 1e-12
    assert abs(oracle.line_segment_overlap([0, 10], [100, 10.5], [25, 300], [75, -7]) - 0.5) < 1e-12
    rng = np.random.default_rng(6)
    spl, epl = rng.uniform(0, 500, (200, 2)), rng.uniform(0, 500, (200, 2))
    epl[:20, 0] = spl[:20, 0] + rng.uniform(-0.9, 0.9, 20)
    epl[20:40, 1] = spl[20:40, 1] + rng.uniform(-0.9, 0.9, 20)
    sp, ep = rng.uniform(0, 500, (200, 2)), rng.uniform(0, 500, (200, 2))
    ref = R.overlap(spl, epl, sp, ep)
    got = np.array([oracle.line_segment_overlap(spl[i], epl[i], sp[i], ep[i]) for i in range(200)])
    np.testing.assert_allclose(got, ref, atol=1e-12)
    assert ((got >= 0) & (got <= 1)).all()


def test_is_finite(oracle):  # src/auxiliar.cpp:353-355
    assert oracle.is_finite(np.eye(4))
    a = np.eye(4)
    a[1, 2] = np.nan
    assert not oracle.is_finite(a)
    a[1, 2] = np.inf
    assert not oracle.is_finite(a)


def test_qr_inverse_eig_vs_numpy(oracle):
    rng = np.random.default_rng(8)
    for k in range(20):
        J = rng.normal(0, 1, (40, 6)) * np.array([1, 1, 1, 30, 30, 30]) ** (k % 3)
        H = J.T @ J
        g = rng.normal(0, 1, 6)
        x, lad, rank = oracle.qr6_solve(H, g)
        assert rank == 6
        np.testing.assert_allclose(x, np.linalg.solve(H, g), rtol=1e-9, atol=1e-12)
        assert abs(lad - np.linalg.slogdet(H)[1]) < 1e-9
        Hi = oracle.inv6(H)
        np.testing.assert_allclose(Hi, np.linalg.inv(H), rtol=1e-8, atol=1e-14)
        np.testing.assert_allclose(oracle.eig6_sym(Hi), np.linalg.eigvalsh(Hi, UPLO="L"), rtol=1e-9, atol=1e-18)
    # rank-deficient: least-squares style truncated solve must not blow up
    H = np.diag([1.0, 2.0, 3.0, 0.0, 0.0, 0.0])
    x, lad, rank = oracle.qr6_solve(H, np.ones(6))
    assert rank == 3 and np.isfinite(x).all()
    # non-symmetric input: only the lower triangle is read, like SelfAdjointEigenSolver
    A = np.diag([1.0, 2, 3, 4, 5, 6])
    A[0, 5] = 100.0
    np.testing.assert_allclose(oracle.eig6_sym(A), [1, 2, 3, 4, 5, 6], atol=1e-12)
