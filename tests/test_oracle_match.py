"""Oracle (matching half) against the committed cv2.BFMatcher golden vectors and, when cv2 is importable,
against cv2 live.  Pins oracle/plstvo_oracle.c: orc_knn2 / orc_match_nnr / orc_match."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = sorted(glob.glob(os.path.join(GOLDEN, "match_*.npz")))


def test_golden_files_present():
    assert len(CASES) >= 9


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_oracle_vs_cv2_golden(oracle, path):
    g = np.load(path)
    d1, d2, nnr = g["d1"], g["d2"], float(g["nnr"])
    idx, dist = oracle.knn2(d1, d2)
    np.testing.assert_array_equal(idx, g["knn_idx"])
    np.testing.assert_array_equal(dist, g["knn_dist"].astype(np.int32))
    if len(d2) >= 2:
        n, m12 = oracle.match_nnr(d1, d2, nnr)
        np.testing.assert_array_equal(m12, g["m12_nnr"])
        assert n == int((m12 >= 0).sum())
        for threads in (False, True):
            n, m = oracle.match(d1, d2, nnr, True, threads)
            np.testing.assert_array_equal(m, g["m12_mutual"])
            assert n == int((m >= 0).sum())


def test_oracle_vs_cv2_live(oracle):
    cv2 = pytest.importorskip("cv2")
    from golden.make_golden import match_from_cv2
    rng = np.random.default_rng(123)
    for n1, n2, lowent in [(257, 190, False), (128, 333, True), (64, 64, True)]:
        if lowent:
            d1 = (rng.integers(0, 2, (n1, 32), dtype=np.uint8) * 255).astype(np.uint8)
            d2 = (rng.integers(0, 2, (n2, 32), dtype=np.uint8) * 255).astype(np.uint8)
        else:
            d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
            d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        for nnr in (0.75, 0.9, 1.0):
            idx, dist, m12, mutual = match_from_cv2(d1, d2, nnr)
            oi, od = oracle.knn2(d1, d2)
            np.testing.assert_array_equal(oi, idx)
            np.testing.assert_array_equal(oracle.match_nnr(d1, d2, nnr)[1], m12)
            np.testing.assert_array_equal(oracle.match(d1, d2, nnr)[1], mutual)


def test_distance_known_answers(oracle):
    """src/matching.cpp:93-109: all-zero vs all-one rows = 256; identical rows = 0."""
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert oracle.distance(z, o) == 256
    assert oracle.distance(o, o) == 0
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert oracle.distance(a, b) == int(np.unpackbits(a ^ b).sum())


def test_ratio_is_float32(oracle):
    """d0 < d1 * nnr in float: 9 < 10 * 0.9f is False (0.9f = 0.899999976), in double it would also be
    False; 18 < 20 * 0.9f -> 18 < 17.9999995 False.  With nnr = 1.0 equality fails (strict <)."""
    q = np.zeros((1, 32), np.uint8)
    t = np.zeros((2, 32), np.uint8)
    t[0, 0] = 0xFF  # distance 8
    t[0, 1] = 0x01  # distance 9
    t[1, 0] = 0xFF
    t[1, 1] = 0x03  # distance 10
    n, m = oracle.match_nnr(q, t, 0.9)
    assert n == 0 and m[0] == -1
    n, m = oracle.match_nnr(q, t, 0.91)
    assert n == 1 and m[0] == 0


def test_degenerate_sizes(oracle):
    d = np.random.default_rng(1).integers(0, 256, (5, 32), dtype=np.uint8)
    n, m = oracle.match_nnr(d, d[:1], 0.9)    # n2 < 2: reference UB, defined as "no match"
    assert n == 0 and (m == -1).all()
    n, m = oracle.match(d, d[:0], 0.9)
    assert n == 0 and (m == -1).all()
    n, m = oracle.match(d[:0], d, 0.9)
    assert n == 0 and len(m) == 0


def test_mutual_matching_is_symmetric_property(oracle):
    """StVO::match with bestLRMatches (src/matching.cpp:63-91): i -> j survives only if j -> i in the reverse problem, so the
    result of match(d2, d1) is the inverse map of match(d1, d2).  Randomised over sizes, entropies and ratios."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(2, 70), st.integers(2, 70), st.integers(0, 2**31 - 1), st.sampled_from([0.6, 0.75, 0.9, 1.0]),
           st.sampled_from(["random", "ties"]))
    def prop(n1, n2, seed, nnr, mode):
        rng = np.random.default_rng(seed)
        if mode == "ties":
            d1 = (rng.integers(0, 2, (n1, 32), dtype=np.uint8) * 255).astype(np.uint8)
            d2 = (rng.integers(0, 2, (n2, 32), dtype=np.uint8) * 255).astype(np.uint8)
        else:
            d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
            d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
            k = min(n1, n2) // 2                      # plant near-duplicates so that matches exist
            d2[:k] = d1[:k] ^ (rng.random((k, 32)) < 0.05).astype(np.uint8)
        n12, m12 = oracle.match(d1, d2, nnr, True)
        n21, m21 = oracle.match(d2, d1, nnr, True)
        assert n12 == n21 == int((m12 >= 0).sum())
        for i, j in enumerate(m12):
            if j >= 0:
                assert m21[j] == i
    prop()
