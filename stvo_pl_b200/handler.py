"""StVO::StereoFrameHandler surface (include/stereoFrameHandler.h:34-101) over the C-ABI, for frames of
pre-extracted features (feature extraction is the reference's own, out of scope here).

Call sequence of app/imagesStVO.cpp:88-124:
    h = StereoFrameHandler(cam, cfg); h.initialize(frame0)
    for frame in frames: h.insertStereoPair(frame); h.optimizePose(); ...read h.curr_frame...; h.updateFrame()
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import types as T
from .engine import Engine


@dataclass
class StereoFrame:
    """What the path reads of StereoFrame (include/stereoFrame.h:59-115) + the per-frame results."""
    features: T.FrameBatch                    # a batch of exactly one frame
    frame_idx: int = 0
    Tfw: np.ndarray = field(default_factory=lambda: np.eye(4))
    Tfw_cov: np.ndarray = field(default_factory=lambda: np.eye(6))
    DT: np.ndarray = field(default_factory=lambda: np.eye(4))
    DT_cov: np.ndarray = field(default_factory=lambda: np.zeros((6, 6)))
    DT_cov_eig: np.ndarray = field(default_factory=lambda: np.zeros(6))
    err_norm: float = -1.0


class StereoFrameHandler:
    def __init__(self, cam: T.PlCamera, cfg: Optional[T.PlConfig] = None, engine: Optional[Engine] = None):
        self.cam, self.cfg = cam, cfg or T.default_config()
        self.engine = engine or Engine()
        self.prev_frame: Optional[StereoFrame] = None
        self.curr_frame: Optional[StereoFrame] = None
        self.matched_pt = np.zeros(0, np.int64)     # prev indices of matched_pt, ascending (list order)
        self.matched_ls = np.zeros(0, np.int64)
        self.n_inliers = self.n_inliers_pt = self.n_inliers_ls = 0
        self._pending = None

    def initialize(self, features: T.FrameBatch, idx: int = 0):
        """src/stereoFrameHandler.cpp:35-52: Tfw = I, Tfw_cov = I, DT = I."""
        assert features.B == 1
        self.prev_frame = StereoFrame(features, idx)
        self.curr_frame = self.prev_frame

    def insertStereoPair(self, features: T.FrameBatch, idx: int = 0):
        """src/stereoFrameHandler.cpp:54-60: new frame + f2fTracking.  The fused device call already computes the
        pose; optimizePose() publishes it."""
        assert features.B == 1 and self.prev_frame is not None
        self.curr_frame = StereoFrame(features, idx)
        self.f2fTracking()

    def f2fTracking(self):
        """src/stereoFrameHandler.cpp:106-129."""
        p = self.prev_frame
        pri = np.zeros(1, dtype=T.PRIOR_DTYPE)
        pri["Tfw"][0], pri["Tfw_cov"][0] = p.Tfw, p.Tfw_cov
        pri["DT"][0], pri["DT_cov"][0], pri["err_norm"][0] = p.DT, p.DT_cov, p.err_norm
        out = self.engine.track_batch(self.cam, self.cfg, p.features, self.curr_frame.features, priors=pri)
        self._pending = out
        self.m12_pt, self.m12_ls = out["m12_pt"], out["m12_ls"]
        self.matched_pt = np.nonzero(self.m12_pt >= 0)[0]
        self.matched_ls = np.nonzero(self.m12_ls >= 0)[0]
        self.n_inliers_pt, self.n_inliers_ls = len(self.matched_pt), len(self.matched_ls)   # :126-128
        self.n_inliers = self.n_inliers_pt + self.n_inliers_ls

    def optimizePose(self):
        """src/stereoFrameHandler.cpp:307-392."""
        r = self._pending["results"][0]
        c = self.curr_frame
        c.DT, c.DT_cov, c.DT_cov_eig, c.err_norm = r["DT"].copy(), r["DT_cov"].copy(), r["DT_cov_eig"].copy(), float(r["err_norm"])
        c.Tfw, c.Tfw_cov = r["Tfw"].copy(), r["Tfw_cov"].copy()
        self.inlier_pt = self._pending["inlier_pt"][self.matched_pt].astype(bool)    # flags in matched_pt order
        self.inlier_ls = self._pending["inlier_ls"][self.matched_ls].astype(bool)
        self.n_inliers_pt, self.n_inliers_ls, self.n_inliers = int(r["n_inliers_pt"]), int(r["n_inliers_ls"]), int(r["n_inliers"])
        self.status = int(r["status"])

    def updateFrame(self):
        """src/stereoFrameHandler.cpp:62-102 (the adaptive FAST threshold belongs to feature extraction)."""
        self.matched_pt = np.zeros(0, np.int64)
        self.matched_ls = np.zeros(0, np.int64)
        self.prev_frame, self.curr_frame = self.curr_frame, None
