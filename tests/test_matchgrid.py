"""StVO::matchGrid (the stereo step): C oracle vs an independent plain-Python restatement on CPU; CUDA kernel
vs the oracle through the C-ABI on the GPU.  Match indices must be BIT-EXACT."""
import numpy as np
import pytest

import ref_matchgrid as RG
from stvo_pl_b200 import stereo_synth as SS, types as T

W_STEREO = T.PlGridWindow(left=10, right=0, up=0, down=0)     # matching_s_ws = 10 (src/stereoFrame.cpp:141-143)
W_WIDE = T.PlGridWindow(left=3, right=2, up=1, down=2)


def test_line_cells_vs_python(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        x1, y1, x2, y2 = rng.uniform(-2, 66), rng.uniform(-2, 50), rng.uniform(-2, 66), rng.uniform(-2, 50)
        got = [tuple(c) for c in oracle.line_cells(x1, y1, x2, y2)]
        assert got == RG.line_cells(x1, y1, x2, y2)
    assert [tuple(c) for c in oracle.line_cells(3.2, 4.9, 3.9, 4.1)] == [(3, 4)]


@pytest.mark.parametrize("n_l,n_r,w,ratio,best_lr,tie", [
    (300, 280, W_STEREO, 0.75, True, False), (300, 280, W_STEREO, 0.75, False, False),
    (200, 260, W_WIDE, 0.9, True, False), (150, 150, W_STEREO, 0.9, True, True), (40, 3, W_WIDE, 0.75, True, False)])
def test_oracle_points_vs_python(oracle, n_l, n_r, w, ratio, best_lr, tie):
    q_cell, d1, t_cell, d2 = SS.make_stereo_points(n_l, n_r, seed=n_l + n_r, tie_stress=tie)
    n, m12 = oracle.match_grid_points(q_cell, d1, t_cell, d2, w, ratio, best_lr)
    n_ref, ref = RG.points(q_cell, d1, t_cell, d2, w, ratio, best_lr)
    np.testing.assert_array_equal(m12, ref)
    assert n == n_ref and (tie or not best_lr or n > 0.3 * min(n_l, n_r) or n_r < 10)


@pytest.mark.parametrize("n_l,n_r,w,best_lr", [(120, 130, W_STEREO, True), (120, 130, W_WIDE, False), (60, 80, W_WIDE, True)])
def test_oracle_lines_vs_python(oracle, n_l, n_r, w, best_lr):
    q_line, d1, t_line, t_dir, d2 = SS.make_stereo_lines(n_l, n_r, seed=7 * n_l)
    n, m12 = oracle.match_grid_lines(q_line, d1, t_line, t_dir, d2, w, 0.75, 0.75, best_lr)
    n_ref, ref = RG.lines(q_line, d1, t_line, t_dir, d2, w, 0.75, 0.75, best_lr)
    np.testing.assert_array_equal(m12, ref)
    assert n == n_ref


def test_gate_is_sequential_in_query_order(oracle):
    """The bestLRMatches gate (:145-150): an earlier query with an equal-or-better distance hides the train from a later
    one even if the earlier query ends up matching something else."""
    d = np.zeros((3, 32), np.uint8)
    d[1, 0] = 0x01                      # query 1 is at distance 1 from train 0, query 0 at distance 0
    t = np.zeros((1, 32), np.uint8)
    cells = np.array([[5, 5]] * 3, np.int32)
    n, m12 = oracle.match_grid_points(cells, d, cells[:1], t, W_WIDE, 0.75, True)
    assert list(m12) == [0, -1, -1]     # query 2 ties the record (0 < 0 is false): never considers the train
    n, m12 = oracle.match_grid_points(cells, d, cells[:1], t, W_WIDE, 0.75, False)
    assert list(m12) == [0, 0, 0]       # single candidate: best_d < INT_MAX * ratio


@pytest.mark.gpu
@pytest.mark.parametrize("n_l,n_r,w,ratio,best_lr,tie", [
    (2000, 2000, W_STEREO, 0.75, True, False), (2000, 1900, W_STEREO, 0.9, False, False),
    (1000, 1200, W_WIDE, 0.75, True, False), (600, 600, W_STEREO, 0.9, True, True), (40, 3, W_WIDE, 0.75, True, False),
    (5, 0, W_STEREO, 0.75, True, False)])
def test_gpu_points_vs_oracle(engine, oracle, n_l, n_r, w, ratio, best_lr, tie):
    frames = [SS.make_stereo_points(max(n_l - 37 * k, 1), max(n_r - 11 * k, 0), seed=100 + k, tie_stress=tie) for k in range(3)]
    q_off = np.concatenate([[0], np.cumsum([len(f[1]) for f in frames])])
    t_off = np.concatenate([[0], np.cumsum([len(f[3]) for f in frames])])
    cat = lambda i, wd: np.concatenate([f[i].reshape(-1, wd) for f in frames])
    total, m12, counts = engine.match_grid_points(q_off, cat(0, 2), cat(1, 32), t_off, cat(2, 2), cat(3, 32), w, ratio, best_lr)
    for k, f in enumerate(frames):
        n_ref, ref = oracle.match_grid_points(f[0], f[1], f[2], f[3], w, ratio, best_lr)
        np.testing.assert_array_equal(m12[q_off[k]:q_off[k + 1]], ref)
        assert counts[k] == n_ref
    assert total == counts.sum()


@pytest.mark.gpu
@pytest.mark.parametrize("n_l,n_r,w,best_lr", [(500, 500, W_STEREO, True), (300, 350, W_WIDE, False), (100, 120, W_WIDE, True)])
def test_gpu_lines_vs_oracle(engine, oracle, n_l, n_r, w, best_lr):
    frames = [SS.make_stereo_lines(n_l - 13 * k, n_r - 7 * k, seed=200 + k) for k in range(3)]
    q_off = np.concatenate([[0], np.cumsum([len(f[1]) for f in frames])])
    t_off = np.concatenate([[0], np.cumsum([len(f[4]) for f in frames])])
    cat = lambda i, wd: np.concatenate([f[i].reshape(-1, wd) for f in frames])
    total, m12, counts = engine.match_grid_lines(q_off, cat(0, 4), cat(1, 32), t_off, cat(2, 4), cat(3, 2), cat(4, 32), w,
                                                 0.75, 0.75, best_lr)
    for k, f in enumerate(frames):
        n_ref, ref = oracle.match_grid_lines(f[0], f[1], f[2], f[3], f[4], w, 0.75, 0.75, best_lr)
        np.testing.assert_array_equal(m12[q_off[k]:q_off[k + 1]], ref)
        assert counts[k] == n_ref


def _clustered_points(n_l, n_r, seed, tie=False):
    """Key points crowded into a handful of cells: query windows hold far more than the 128 candidates of the fast path."""
    q_cell, d1, t_cell, d2 = SS.make_stereo_points(n_l, n_r, seed=seed, tie_stress=tie)
    rng = np.random.default_rng(seed)
    t_cell = t_cell.copy()
    q_cell = q_cell.copy()
    hot_t = rng.random(n_r) < 0.6
    t_cell[hot_t] = np.stack([rng.integers(20, 22, hot_t.sum()), rng.integers(10, 11, hot_t.sum())], 1)
    hot_q = rng.random(n_l) < 0.5
    q_cell[hot_q] = np.stack([rng.integers(21, 30, hot_q.sum()), rng.integers(10, 11, hot_q.sum())], 1)
    return q_cell, d1, t_cell, d2


@pytest.mark.gpu
@pytest.mark.parametrize("best_lr,tie", [(True, False), (False, False), (True, True)])
def test_gpu_points_clustered_frame_takes_the_unbounded_path(engine, oracle, best_lr, tie):
    """One dense frame between two ordinary ones: > 128 candidates per window (the reference has no limit,
    src/matching.cpp:128-139); the dense frame must stay exact and must not disturb its neighbours."""
    frames = [SS.make_stereo_points(700, 650, seed=300, tie_stress=tie), _clustered_points(1500, 1400, 301, tie),
              SS.make_stereo_points(500, 520, seed=302, tie_stress=tie)]
    # the dense frame really overflows the fast path
    cells = frames[1][2]
    per_cell = np.bincount(cells[:, 0] * 48 + cells[:, 1])
    assert per_cell.max() > 128
    q_off = np.concatenate([[0], np.cumsum([len(f[1]) for f in frames])])
    t_off = np.concatenate([[0], np.cumsum([len(f[3]) for f in frames])])
    cat = lambda i, wd: np.concatenate([f[i].reshape(-1, wd) for f in frames])
    total, m12, counts = engine.match_grid_points(q_off, cat(0, 2), cat(1, 32), t_off, cat(2, 2), cat(3, 32), W_STEREO, 0.75, best_lr)
    for k, f in enumerate(frames):
        n_ref, ref = oracle.match_grid_points(f[0], f[1], f[2], f[3], W_STEREO, 0.75, best_lr)
        np.testing.assert_array_equal(m12[q_off[k]:q_off[k + 1]], ref)
        assert counts[k] == n_ref
    assert total == counts.sum()


@pytest.mark.gpu
@pytest.mark.parametrize("best_lr", [True, False])
def test_gpu_lines_clustered_frame_takes_the_unbounded_path(engine, oracle, best_lr):
    """Long segments rasterised into many cells plus two windows per query: the candidate set of a query exceeds 128."""
    q_line, d1, t_line, t_dir, d2 = SS.make_stereo_lines(400, 900, seed=410)
    rng = np.random.default_rng(411)
    t_line = t_line.copy()
    hot = rng.random(len(t_line)) < 0.7        # most right-image segments cross the same band of the grid
    n_hot = int(hot.sum())
    t_line[hot] = np.stack([rng.uniform(8, 12, n_hot), rng.uniform(18, 22, n_hot), rng.uniform(40, 50, n_hot),
                            rng.uniform(18, 26, n_hot)], 1)
    v = t_line[:, 2:4] - t_line[:, 0:2]
    t_dir = v / np.linalg.norm(v, axis=1, keepdims=True)
    q_line = q_line.copy()
    hq = rng.random(len(q_line)) < 0.5
    q_line[hq] = np.stack([rng.integers(24, 30, hq.sum()), rng.integers(18, 24, hq.sum()), rng.integers(30, 44, hq.sum()),
                           rng.integers(18, 24, hq.sum())], 1)
    frames = [SS.make_stereo_lines(120, 130, seed=412), (q_line, d1, t_line, t_dir, d2)]
    q_off = np.concatenate([[0], np.cumsum([len(f[1]) for f in frames])])
    t_off = np.concatenate([[0], np.cumsum([len(f[4]) for f in frames])])
    cat = lambda i, wd: np.concatenate([f[i].reshape(-1, wd) for f in frames])
    total, m12, counts = engine.match_grid_lines(q_off, cat(0, 4), cat(1, 32), t_off, cat(2, 4), cat(3, 2), cat(4, 32), W_WIDE,
                                                 0.75, 0.75, best_lr)
    for k, f in enumerate(frames):
        n_ref, ref = oracle.match_grid_lines(f[0], f[1], f[2], f[3], f[4], W_WIDE, 0.75, 0.75, best_lr)
        np.testing.assert_array_equal(m12[q_off[k]:q_off[k + 1]], ref)
        assert counts[k] == n_ref
    assert counts[1] >= 0
