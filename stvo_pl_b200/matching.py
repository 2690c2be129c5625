"""include/matching.h surface over the C-ABI: StVO::matchNNR / StVO::match (src/matching.cpp:41-91).

Same names, argument meaning and error behaviour as the reference: descriptors are N x 32 uint8 matrices, `nnr` is
narrowed to float, the result is (count, matches_12) with -1 for "no match"; size problems raise.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine

_default: Engine | None = None


def default_engine() -> Engine:
    global _default
    if _default is None:
        _default = Engine()
    return _default


def _check(desc) -> np.ndarray:
    d = np.asarray(desc)
    if d.dtype != np.uint8 or d.ndim != 2 or (d.shape[0] and d.shape[1] != 32):
        raise RuntimeError("[matchNNR] descriptors must be N x 32 uint8")     # cv::Mat N x 32 CV_8UC1
    return np.ascontiguousarray(d)


def matchNNR(desc1, desc2, nnr: float, engine: Engine | None = None):
    """int matchNNR(const cv::Mat&, const cv::Mat&, float nnr, std::vector<int>& matches_12) — src/matching.cpp:41."""
    eng = engine or default_engine()
    return eng.match_nnr(_check(desc1), _check(desc2), float(np.float32(nnr)))


def match(desc1, desc2, nnr: float, best_lr_matches: bool = True, engine: Engine | None = None):
    """int match(...) — src/matching.cpp:63; `best_lr_matches` = Config::bestLRMatches()."""
    eng = engine or default_engine()
    return eng.match(_check(desc1), _check(desc2), float(np.float32(nnr)), best_lr_matches)
