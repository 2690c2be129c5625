"""ctypes binding of the CUDA library (stvo_pl_b200/lib/libplstvo_b200.so) through its C-ABI
(include/plstvo.h).  There is no CPU fallback: a missing extension or a missing sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLSTVO_LIB") or os.path.join(_HERE, "lib", "libplstvo_b200.so")   # PLSTVO_LIB: A/B builds

ERRORS = {-1: "PLSTVO_E_INVALID", -2: "PLSTVO_E_TOO_LARGE", -3: "PLSTVO_E_CUDA", -4: "PLSTVO_E_NO_DEVICE",
          -5: "PLSTVO_E_SIZE"}


class PlstvoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise RuntimeError(f"CUDA extension missing: {path}. Build it with `python -m stvo_pl_b200.build` "
                           "(there is no CPU fallback).")
    L = C.CDLL(path)
    vp, i32p, u8p, dp, fp = C.c_void_p, T.c_int32_p, T.c_uint8_p, T.c_double_p, T.c_float_p
    cam, cfg = C.POINTER(T.PlCamera), C.POINTER(T.PlConfig)
    fb, mb = C.POINTER(T.PlFrameBatch), C.POINTER(T.PlMatchedBatch)
    sig = {
        "plstvo_version": (C.c_int, []),
        "plstvo_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "plstvo_destroy": (None, [vp]),
        "plstvo_last_error": (C.c_char_p, [vp]),
        "plstvo_default_config": (None, [cfg]),
        "plstvo_kitti_config": (None, [cfg]),
        "plstvo_match_nnr": (C.c_int, [vp, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_float, i32p]),
        "plstvo_match": (C.c_int, [vp, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_float, C.c_int, i32p]),
        "plstvo_match_batch": (C.c_int, [vp, C.c_int, u8p, i32p, u8p, i32p, C.c_float, C.c_int, i32p, i32p]),
        "plstvo_match_grid_points": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, T.PlGridWindow, C.c_int, C.c_double, i32p, i32p,
                                               u8p, i32p, i32p, u8p, i32p, i32p]),
        "plstvo_match_grid_lines": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, T.PlGridWindow, C.c_int, C.c_double,
                                              C.c_double, i32p, i32p, u8p, i32p, dp, dp, u8p, i32p, i32p]),
        "plstvo_default_stereo_config": (None, [C.POINTER(T.PlStereoConfig)]),
        "plstvo_stereo_lift_points": (C.c_int, [vp, cam, C.POINTER(T.PlStereoConfig), C.c_int, i32p, fp, i32p, u8p, i32p, fp,
                                                i32p, dp, dp, dp, dp, i32p, u8p, i32p, i32p]),
        "plstvo_stereo_lift_lines": (C.c_int, [vp, cam, C.POINTER(T.PlStereoConfig), C.c_int, i32p, fp, fp, i32p, u8p, i32p,
                                               fp, i32p, dp, dp, dp, dp, dp, dp, dp, dp, dp, i32p, u8p, i32p, i32p]),
        "plstvo_default_stereo_match_config": (None, [C.POINTER(T.PlStereoMatchConfig)]),
        "plstvo_match_stereo_points": (C.c_int, [vp, cam, C.POINTER(T.PlStereoMatchConfig), C.POINTER(T.PlStereoConfig), C.c_int,
                                                 i32p, fp, i32p, u8p, i32p, fp, u8p, i32p, dp, dp, dp, dp, i32p, u8p, i32p, i32p]),
        "plstvo_match_stereo_lines": (C.c_int, [vp, cam, C.POINTER(T.PlStereoMatchConfig), C.POINTER(T.PlStereoConfig), C.c_int,
                                                i32p, fp, fp, i32p, u8p, i32p, fp, u8p, i32p, dp, dp, dp, dp, dp, dp, dp, dp, dp,
                                                i32p, u8p, i32p, i32p]),
        "plstvo_track_stereo_batch": (C.c_int, [vp, cam, cfg, C.POINTER(T.PlStereoMatchConfig), C.POINTER(T.PlStereoConfig),
                                                C.POINTER(T.PlStereoFeatures), C.POINTER(T.PlStereoFeatures), vp, vp, i32p]),
        "plstvo_track_stereo_sequence": (C.c_int, [vp, cam, cfg, C.POINTER(T.PlStereoMatchConfig), C.POINTER(T.PlStereoConfig),
                                                   C.POINTER(T.PlStereoFeatures), vp, vp, i32p]),
        "plstvo_track_stereo_batch_async": (C.c_int, [vp, cam, cfg, C.POINTER(T.PlStereoMatchConfig), C.POINTER(T.PlStereoConfig),
                                                      C.POINTER(T.PlStereoFeatures), C.POINTER(T.PlStereoFeatures), vp, vp, i32p]),
        "plstvo_track_stereo_sequence_async": (C.c_int, [vp, cam, cfg, C.POINTER(T.PlStereoMatchConfig), C.POINTER(T.PlStereoConfig),
                                                         C.POINTER(T.PlStereoFeatures), vp, vp, i32p]),
        "plstvo_f2f_tracking": (C.c_int, [vp, cfg, fb, fb, i32p, i32p, i32p]),
        "plstvo_optimize_pose": (C.c_int, [vp, cam, cfg, mb, vp, vp, u8p, u8p]),
        "plstvo_track_batch": (C.c_int, [vp, cam, cfg, fb, fb, vp, vp, i32p, i32p, u8p, u8p]),
        "plstvo_track_batch_async": (C.c_int, [vp, cam, cfg, fb, fb, vp, vp, i32p, i32p, u8p, u8p]),
        "plstvo_wait": (C.c_int, [vp, C.c_int]),
        "plstvo_batch_upload": (C.c_int, [vp, cam, cfg, fb, fb, vp, C.POINTER(vp)]),
        "plstvo_batch_run": (C.c_int, [vp, vp]),
        "plstvo_batch_run_timed": (C.c_int, [vp, vp, C.c_int, C.c_int, C.POINTER(C.c_float)]),
        "plstvo_batch_download": (C.c_int, [vp, vp, vp, i32p, i32p, u8p, u8p]),
        "plstvo_batch_free": (None, [vp, vp]),
        "plstvo_synchronize": (C.c_int, [vp]),
        "plstvo_host_alloc": (vp, [C.c_size_t]),
        "plstvo_host_free": (None, [vp]),
        "plstvo_launch_count": (C.c_int64, [vp]),
        "plstvo_batch_kernel_times": (C.c_int, [vp, vp, C.c_int, dp, dp, i32p, i32p]),
        "plstvo_batch_stage_times": (C.c_int, [vp, vp, C.c_int, dp, i32p]),
        "plstvo_gn_eval_stream": (C.c_int, [vp, cam, cfg, mb, dp, C.c_int, dp, dp, dp, C.POINTER(C.c_float)]),
        "plstvo_popc_rate": (C.c_int, [vp, dp]),
        "plstvo_debug_algebra": (C.c_int, [vp, C.c_int, dp, dp, dp, dp, dp, dp]),
        "plstvo_debug_select": (C.c_int, [vp, C.c_int, i32p, dp, i32p, dp, C.c_int, dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    return L


EXPORTED_SYMBOLS = [
    "plstvo_version", "plstvo_create", "plstvo_destroy", "plstvo_last_error", "plstvo_default_config",
    "plstvo_kitti_config", "plstvo_match_nnr", "plstvo_match", "plstvo_match_batch", "plstvo_match_grid_points",
    "plstvo_match_grid_lines", "plstvo_default_stereo_config", "plstvo_stereo_lift_points", "plstvo_stereo_lift_lines",
    "plstvo_default_stereo_match_config", "plstvo_match_stereo_points", "plstvo_match_stereo_lines", "plstvo_track_stereo_batch",
    "plstvo_track_stereo_sequence", "plstvo_track_stereo_batch_async", "plstvo_track_stereo_sequence_async",
    "plstvo_f2f_tracking",
    "plstvo_optimize_pose", "plstvo_track_batch", "plstvo_track_batch_async", "plstvo_wait", "plstvo_batch_upload", "plstvo_batch_run",
    "plstvo_batch_run_timed", "plstvo_batch_download", "plstvo_batch_free", "plstvo_synchronize",
    "plstvo_host_alloc", "plstvo_host_free", "plstvo_launch_count", "plstvo_batch_kernel_times", "plstvo_batch_stage_times",
    "plstvo_gn_eval_stream", "plstvo_popc_rate", "plstvo_debug_algebra", "plstvo_debug_select"]


def _p(a: Optional[np.ndarray], typ):
    return C.cast(None, typ) if a is None else a.ctypes.data_as(typ)


def bind_to_gpu_numa(device: int) -> Optional[int]:
    """Restricts the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned buffers allocated afterwards
    are local to the GPU's PCIe root (matters with one process per GPU on a two-socket host: the e2e path moves ~40 GB/s per
    GPU).  Returns the node, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class PinnedPool:
    """Pinned host arrays (cudaMallocHost) so that H2D/D2H run at full PCIe rate without staging."""

    def __init__(self, lib):
        self.lib = lib
        self._ptrs = []

    def empty(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = self.lib.plstvo_host_alloc(max(n, 1))
        if not ptr:
            raise MemoryError("cudaMallocHost failed")
        self._ptrs.append(ptr)
        buf = (C.c_uint8 * max(n, 1)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def copy(self, a: Optional[np.ndarray]) -> Optional[np.ndarray]:
        if a is None:
            return None
        out = self.empty(a.shape, a.dtype)
        out[...] = a
        return out

    def pin_frames(self, f: T.FrameBatch) -> T.FrameBatch:
        return T.FrameBatch(**{k: self.copy(getattr(f, k)) for k in (
            "pt_off", "ls_off", "pdesc", "ldesc", "pt_P", "pt_pl", "pt_sigma2", "ls_sP", "ls_eP", "ls_le",
            "ls_spl", "ls_epl", "ls_sigma2", "ls_level")})

    def close(self):
        for p in self._ptrs:
            self.lib.plstvo_host_free(p)
        self._ptrs = []


class DeviceBatch:
    """Inputs resident in HBM (PlDeviceBatch)."""

    def __init__(self, eng: "Engine", handle, prev: T.FrameBatch):
        self.eng, self.handle = eng, handle
        self.B, self.n_pt, self.n_ls = prev.B, prev.n_pt, prev.n_ls

    def run(self):
        self.eng._ck(self.eng.lib.plstvo_batch_run(self.eng.ctx, self.handle))

    def run_timed(self, iters: int, flush_l2: bool = False) -> float:
        ms = C.c_float(0)
        self.eng._ck(self.eng.lib.plstvo_batch_run_timed(self.eng.ctx, self.handle, iters, int(flush_l2), C.byref(ms)))
        return float(ms.value)

    def kernel_times(self, iters: int = 3):
        a, b = np.zeros(1), np.zeros(1)
        nt, npairs = np.zeros(1, np.int32), np.zeros(1, np.int32)
        self.eng._ck(self.eng.lib.plstvo_batch_kernel_times(
            self.eng.ctx, self.handle, iters, a.ctypes.data_as(T.c_double_p), b.ctypes.data_as(T.c_double_p),
            nt.ctypes.data_as(T.c_int32_p), npairs.ctypes.data_as(T.c_int32_p)))
        return dict(ms_match=float(a[0]), ms_solve=float(b[0]), n_tiles=int(nt[0]), n_pairs=int(npairs[0]))

    def stage_times(self, iters: int = 3):
        ms, cnt = np.zeros(8), np.zeros(4, np.int32)
        self.eng._ck(self.eng.lib.plstvo_batch_stage_times(self.eng.ctx, self.handle, iters, ms.ctypes.data_as(T.c_double_p),
                                                           cnt.ctypes.data_as(T.c_int32_p)))
        return dict(ms_expand=float(ms[0]), ms_distance=float(ms[1]), ms_resolve=float(ms[2]), ms_lists=float(ms[3]),
                    ms_solve=float(ms[3] + ms[4]), ms_optimize_pose=float(ms[4]), ms_gn_stage1=float(ms[5]),
                    ms_outliers=float(ms[6]), ms_gn_stage2=float(ms[7]),
                    tensor_core_form=bool(cnt[0] & 1), streamed_solver=bool(cnt[0] & 2), delegated_to_fp64=int(cnt[0] >> 8), n_items=int(cnt[1]),
                    n_problems=int(cnt[2]), n_pairs=int(cnt[3]))

    def download(self):
        res = np.zeros(self.B, dtype=T.POSE_RESULT_DTYPE)
        m12_pt, m12_ls = np.full(self.n_pt, -1, np.int32), np.full(self.n_ls, -1, np.int32)
        inl_pt, inl_ls = np.zeros(self.n_pt, np.uint8), np.zeros(self.n_ls, np.uint8)
        self.eng._ck(self.eng.lib.plstvo_batch_download(
            self.eng.ctx, self.handle, res.ctypes.data, _p(m12_pt, T.c_int32_p), _p(m12_ls, T.c_int32_p),
            _p(inl_pt, T.c_uint8_p), _p(inl_ls, T.c_uint8_p)))
        return dict(results=res, m12_pt=m12_pt, m12_ls=m12_ls, inlier_pt=inl_pt, inlier_ls=inl_ls)

    def free(self):
        if self.handle:
            self.eng.lib.plstvo_batch_free(self.eng.ctx, self.handle)
            self.handle = None


class Engine:
    """One PlContext (one GPU)."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        ctx = C.c_void_p()
        rc = self.lib.plstvo_create(device, C.byref(ctx))
        if rc != 0:
            raise PlstvoError(rc, "plstvo_create failed: no sm_100 CUDA device visible (no CPU fallback exists)")
        self.ctx = ctx
        self.pinned = PinnedPool(self.lib)

    def close(self):
        if self.ctx:
            self.pinned.close()
            self.lib.plstvo_destroy(self.ctx)
            self.ctx = None

    def _ck(self, rc: int) -> int:
        if rc < 0:
            raise PlstvoError(rc, (self.lib.plstvo_last_error(self.ctx) or b"").decode())
        return rc

    @property
    def launches(self) -> int:
        return int(self.lib.plstvo_launch_count(self.ctx))

    # ---- matching.h surface ----
    def match_nnr(self, d1, d2, nnr):
        d1, d2 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32), np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        m12 = np.full(len(d1), -1, np.int32)
        n = self._ck(self.lib.plstvo_match_nnr(self.ctx, _p(d1, T.c_uint8_p), len(d1), _p(d2, T.c_uint8_p), len(d2),
                                               32, C.c_float(nnr), _p(m12, T.c_int32_p)))
        return n, m12

    def match(self, d1, d2, nnr, best_lr=True):
        d1, d2 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32), np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        m12 = np.full(len(d1), -1, np.int32)
        n = self._ck(self.lib.plstvo_match(self.ctx, _p(d1, T.c_uint8_p), len(d1), _p(d2, T.c_uint8_p), len(d2), 32,
                                           C.c_float(nnr), int(best_lr), _p(m12, T.c_int32_p)))
        return n, m12

    def match_batch(self, d1, off1, d2, off2, nnr, best_lr=True):
        d1, d2 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32), np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        off1, off2 = np.ascontiguousarray(off1, np.int32), np.ascontiguousarray(off2, np.int32)
        B = len(off1) - 1
        m12 = np.full(len(d1), -1, np.int32)
        counts = np.zeros(B, np.int32)
        n = self._ck(self.lib.plstvo_match_batch(self.ctx, B, _p(d1, T.c_uint8_p), _p(off1, T.c_int32_p),
                                                 _p(d2, T.c_uint8_p), _p(off2, T.c_int32_p), C.c_float(nnr),
                                                 int(best_lr), _p(m12, T.c_int32_p), _p(counts, T.c_int32_p)))
        return n, m12, counts

    def match_grid_points(self, q_off, q_cell, d1, t_off, t_cell, d2, window: T.PlGridWindow, ratio: float,
                          best_lr=True, rows=T.GRID_ROWS, cols=T.GRID_COLS):
        """StVO::matchGrid, points overload (src/matching.cpp:111-177), batched over frames."""
        q_off, t_off = np.ascontiguousarray(q_off, np.int32), np.ascontiguousarray(t_off, np.int32)
        q_cell, t_cell = np.ascontiguousarray(q_cell, np.int32).reshape(-1, 2), np.ascontiguousarray(t_cell, np.int32).reshape(-1, 2)
        d1, d2 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32), np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        B = len(q_off) - 1
        m12, counts = np.full(len(d1), -1, np.int32), np.zeros(B, np.int32)
        n = self._ck(self.lib.plstvo_match_grid_points(
            self.ctx, B, rows, cols, window, int(best_lr), float(ratio), _p(q_off, T.c_int32_p), _p(q_cell, T.c_int32_p),
            _p(d1, T.c_uint8_p), _p(t_off, T.c_int32_p), _p(t_cell, T.c_int32_p), _p(d2, T.c_uint8_p),
            _p(m12, T.c_int32_p), _p(counts, T.c_int32_p)))
        return n, m12, counts

    def match_grid_lines(self, q_off, q_line, d1, t_off, t_line, t_dir, d2, window: T.PlGridWindow, ratio: float,
                         line_sim_th: float, best_lr=True, rows=T.GRID_ROWS, cols=T.GRID_COLS):
        """StVO::matchGrid, lines overload (src/matching.cpp:179-258), batched over frames."""
        q_off, t_off = np.ascontiguousarray(q_off, np.int32), np.ascontiguousarray(t_off, np.int32)
        q_line = np.ascontiguousarray(q_line, np.int32).reshape(-1, 4)
        t_line = np.ascontiguousarray(t_line, np.float64).reshape(-1, 4)
        t_dir = np.ascontiguousarray(t_dir, np.float64).reshape(-1, 2)
        d1, d2 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32), np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        B = len(q_off) - 1
        m12, counts = np.full(len(d1), -1, np.int32), np.zeros(B, np.int32)
        n = self._ck(self.lib.plstvo_match_grid_lines(
            self.ctx, B, rows, cols, window, int(best_lr), float(ratio), float(line_sim_th), _p(q_off, T.c_int32_p),
            _p(q_line, T.c_int32_p), _p(d1, T.c_uint8_p), _p(t_off, T.c_int32_p), _p(t_line, T.c_double_p),
            _p(t_dir, T.c_double_p), _p(d2, T.c_uint8_p), _p(m12, T.c_int32_p), _p(counts, T.c_int32_p)))
        return n, m12, counts

    def stereo_lift_points(self, cam, scfg: T.PlStereoConfig, l_off, kp_l, octave_l, desc_l, r_off, kp_r, m12):
        """Stereo matches -> PointFeature records (src/stereoFrame.cpp:149-172), batched over frames.  Every output array has
        one slot per left keypoint; frame p's survivors are dense from l_off[p], counts[p] of them."""
        l_off, r_off = np.ascontiguousarray(l_off, np.int32), np.ascontiguousarray(r_off, np.int32)
        kp_l, kp_r = np.ascontiguousarray(kp_l, np.float32).reshape(-1, 2), np.ascontiguousarray(kp_r, np.float32).reshape(-1, 2)
        octave_l, m12 = np.ascontiguousarray(octave_l, np.int32), np.ascontiguousarray(m12, np.int32)
        desc_l = np.ascontiguousarray(desc_l, np.uint8).reshape(-1, 32)
        B, n = len(l_off) - 1, len(kp_l)
        out = dict(pl=np.zeros((n, 2)), disp=np.zeros(n), P=np.zeros((n, 3)), sigma2=np.zeros(n), level=np.zeros(n, np.int32),
                   desc=np.zeros((n, 32), np.uint8), src_idx=np.full(n, -1, np.int32), counts=np.zeros(B, np.int32))
        total = self._ck(self.lib.plstvo_stereo_lift_points(
            self.ctx, C.byref(cam), C.byref(scfg), B, _p(l_off, T.c_int32_p), _p(kp_l, T.c_float_p), _p(octave_l, T.c_int32_p),
            _p(desc_l, T.c_uint8_p), _p(r_off, T.c_int32_p), _p(kp_r, T.c_float_p), _p(m12, T.c_int32_p),
            _p(out["pl"], T.c_double_p), _p(out["disp"], T.c_double_p), _p(out["P"], T.c_double_p),
            _p(out["sigma2"], T.c_double_p), _p(out["level"], T.c_int32_p), _p(out["desc"], T.c_uint8_p),
            _p(out["src_idx"], T.c_int32_p), _p(out["counts"], T.c_int32_p)))
        return total, out

    def stereo_lift_lines(self, cam, scfg: T.PlStereoConfig, l_off, seg_l, angle_l, octave_l, desc_l, r_off, seg_r, m12):
        """Stereo matches -> LineFeature records (src/stereoFrame.cpp:348-397), batched over frames."""
        l_off, r_off = np.ascontiguousarray(l_off, np.int32), np.ascontiguousarray(r_off, np.int32)
        seg_l, seg_r = np.ascontiguousarray(seg_l, np.float32).reshape(-1, 4), np.ascontiguousarray(seg_r, np.float32).reshape(-1, 4)
        angle_l = np.ascontiguousarray(angle_l, np.float32)
        octave_l, m12 = np.ascontiguousarray(octave_l, np.int32), np.ascontiguousarray(m12, np.int32)
        desc_l = np.ascontiguousarray(desc_l, np.uint8).reshape(-1, 32)
        B, n = len(l_off) - 1, len(seg_l)
        out = dict(spl=np.zeros((n, 2)), epl=np.zeros((n, 2)), sdisp=np.zeros(n), edisp=np.zeros(n), sP=np.zeros((n, 3)),
                   eP=np.zeros((n, 3)), le=np.zeros((n, 3)), angle=np.zeros(n), sigma2=np.zeros(n),
                   level=np.zeros(n, np.int32), desc=np.zeros((n, 32), np.uint8), src_idx=np.full(n, -1, np.int32),
                   counts=np.zeros(B, np.int32))
        total = self._ck(self.lib.plstvo_stereo_lift_lines(
            self.ctx, C.byref(cam), C.byref(scfg), B, _p(l_off, T.c_int32_p), _p(seg_l, T.c_float_p), _p(angle_l, T.c_float_p),
            _p(octave_l, T.c_int32_p), _p(desc_l, T.c_uint8_p), _p(r_off, T.c_int32_p), _p(seg_r, T.c_float_p),
            _p(m12, T.c_int32_p), _p(out["spl"], T.c_double_p), _p(out["epl"], T.c_double_p), _p(out["sdisp"], T.c_double_p),
            _p(out["edisp"], T.c_double_p), _p(out["sP"], T.c_double_p), _p(out["eP"], T.c_double_p),
            _p(out["le"], T.c_double_p), _p(out["angle"], T.c_double_p), _p(out["sigma2"], T.c_double_p),
            _p(out["level"], T.c_int32_p), _p(out["desc"], T.c_uint8_p), _p(out["src_idx"], T.c_int32_p),
            _p(out["counts"], T.c_int32_p)))
        return total, out

    def match_stereo_points(self, cam, mcfg: T.PlStereoMatchConfig, scfg: T.PlStereoConfig, l_off, kp_l, octave_l, desc_l,
                            r_off, kp_r, desc_r, out=None):
        """StereoFrame::matchStereoPoints (src/stereoFrame.cpp:120-173): grid cells, matchGrid and the lifting in one device
        pass.  Returns (total, out) like stereo_lift_points, plus out["m12"]."""
        l_off, r_off = np.ascontiguousarray(l_off, np.int32), np.ascontiguousarray(r_off, np.int32)
        kp_l, kp_r = np.ascontiguousarray(kp_l, np.float32).reshape(-1, 2), np.ascontiguousarray(kp_r, np.float32).reshape(-1, 2)
        octave_l = np.ascontiguousarray(octave_l, np.int32)
        desc_l, desc_r = np.ascontiguousarray(desc_l, np.uint8).reshape(-1, 32), np.ascontiguousarray(desc_r, np.uint8).reshape(-1, 32)
        B, n = len(l_off) - 1, len(kp_l)
        if out is None:
            out = self.stereo_outputs(n, B, lines=False)
        total = self._ck(self.lib.plstvo_match_stereo_points(
            self.ctx, C.byref(cam), C.byref(mcfg), C.byref(scfg), B, _p(l_off, T.c_int32_p), _p(kp_l, T.c_float_p),
            _p(octave_l, T.c_int32_p), _p(desc_l, T.c_uint8_p), _p(r_off, T.c_int32_p), _p(kp_r, T.c_float_p),
            _p(desc_r, T.c_uint8_p), _p(out["m12"], T.c_int32_p), _p(out["pl"], T.c_double_p), _p(out["disp"], T.c_double_p),
            _p(out["P"], T.c_double_p), _p(out["sigma2"], T.c_double_p), _p(out["level"], T.c_int32_p),
            _p(out["desc"], T.c_uint8_p), _p(out["src_idx"], T.c_int32_p), _p(out["counts"], T.c_int32_p)))
        return total, out

    def stereo_outputs(self, n: int, B: int, lines: bool, pinned: bool = False) -> dict:
        """Output arrays of match_stereo_points / _lines for n left features in B frames (pinned=True: cudaMallocHost)."""
        mk = (lambda shape, dt: self.pinned.empty(shape, dt)) if pinned else (lambda shape, dt: np.zeros(shape, dt))
        if lines:
            spec = dict(m12=((n,), np.int32), spl=((n, 2), np.float64), epl=((n, 2), np.float64), sdisp=((n,), np.float64),
                        edisp=((n,), np.float64), sP=((n, 3), np.float64), eP=((n, 3), np.float64), le=((n, 3), np.float64),
                        angle=((n,), np.float64), sigma2=((n,), np.float64), level=((n,), np.int32), desc=((n, 32), np.uint8),
                        src_idx=((n,), np.int32), counts=((B,), np.int32))
        else:
            spec = dict(m12=((n,), np.int32), pl=((n, 2), np.float64), disp=((n,), np.float64), P=((n, 3), np.float64),
                        sigma2=((n,), np.float64), level=((n,), np.int32), desc=((n, 32), np.uint8), src_idx=((n,), np.int32),
                        counts=((B,), np.int32))
        out = {k: mk(shape, dt) for k, (shape, dt) in spec.items()}
        out["m12"][...] = -1
        out["src_idx"][...] = -1
        return out

    def match_stereo_lines(self, cam, mcfg: T.PlStereoMatchConfig, scfg: T.PlStereoConfig, l_off, seg_l, angle_l, octave_l,
                           desc_l, r_off, seg_r, desc_r, out=None):
        """StereoFrame::matchStereoLines (src/stereoFrame.cpp:309-398) in one device pass."""
        l_off, r_off = np.ascontiguousarray(l_off, np.int32), np.ascontiguousarray(r_off, np.int32)
        seg_l, seg_r = np.ascontiguousarray(seg_l, np.float32).reshape(-1, 4), np.ascontiguousarray(seg_r, np.float32).reshape(-1, 4)
        angle_l, octave_l = np.ascontiguousarray(angle_l, np.float32), np.ascontiguousarray(octave_l, np.int32)
        desc_l, desc_r = np.ascontiguousarray(desc_l, np.uint8).reshape(-1, 32), np.ascontiguousarray(desc_r, np.uint8).reshape(-1, 32)
        B, n = len(l_off) - 1, len(seg_l)
        if out is None:
            out = self.stereo_outputs(n, B, lines=True)
        total = self._ck(self.lib.plstvo_match_stereo_lines(
            self.ctx, C.byref(cam), C.byref(mcfg), C.byref(scfg), B, _p(l_off, T.c_int32_p), _p(seg_l, T.c_float_p),
            _p(angle_l, T.c_float_p), _p(octave_l, T.c_int32_p), _p(desc_l, T.c_uint8_p), _p(r_off, T.c_int32_p),
            _p(seg_r, T.c_float_p), _p(desc_r, T.c_uint8_p), _p(out["m12"], T.c_int32_p), _p(out["spl"], T.c_double_p),
            _p(out["epl"], T.c_double_p), _p(out["sdisp"], T.c_double_p), _p(out["edisp"], T.c_double_p),
            _p(out["sP"], T.c_double_p), _p(out["eP"], T.c_double_p), _p(out["le"], T.c_double_p), _p(out["angle"], T.c_double_p),
            _p(out["sigma2"], T.c_double_p), _p(out["level"], T.c_int32_p), _p(out["desc"], T.c_uint8_p),
            _p(out["src_idx"], T.c_int32_p), _p(out["counts"], T.c_int32_p)))
        return total, out

    def track_stereo_batch(self, cam, cfg, mcfg, scfg, prev: dict, curr: dict, priors=None, results=None):
        """Raw stereo features of B (prev, curr) frame pairs -> poses: matchStereoPoints / Lines for both frames, f2fTracking and
        optimizePose, the lifted records resident in HBM.  prev / curr: dicts with the PlStereoFeatures fields.
        Returns (results, n_stereo[B, 4])."""
        ps, keep_p = T.stereo_features_as_c(prev)
        cs, keep_c = T.stereo_features_as_c(curr)
        B = ps.B
        if results is None:
            results = np.zeros(B, dtype=T.POSE_RESULT_DTYPE)
        n_stereo = np.zeros((B, 4), np.int32)
        self._ck(self.lib.plstvo_track_stereo_batch(self.ctx, C.byref(cam), C.byref(cfg), C.byref(mcfg), C.byref(scfg), C.byref(ps),
                                                    C.byref(cs), priors.ctypes.data if priors is not None else None,
                                                    results.ctypes.data, _p(n_stereo, T.c_int32_p)))
        del keep_p, keep_c
        return results, n_stereo

    def track_stereo_batch_async(self, cam, cfg, mcfg, scfg, prev_c, curr_c, results, n_stereo, priors=None) -> int:
        """Streaming form: prev_c / curr_c are PlStereoFeatures structs built once by T.stereo_features_as_c over PINNED arrays
        (keep the returned keep-alive dicts); results / n_stereo are caller-owned (pinned) arrays, valid after wait(ticket)."""
        return self._ck(self.lib.plstvo_track_stereo_batch_async(
            self.ctx, C.byref(cam), C.byref(cfg), C.byref(mcfg), C.byref(scfg), C.byref(prev_c), C.byref(curr_c),
            priors.ctypes.data if priors is not None else None, results.ctypes.data, _p(n_stereo, T.c_int32_p)))

    def track_stereo_sequence_async(self, cam, cfg, mcfg, scfg, frames_c, results, n_stereo, priors=None) -> int:
        return self._ck(self.lib.plstvo_track_stereo_sequence_async(
            self.ctx, C.byref(cam), C.byref(cfg), C.byref(mcfg), C.byref(scfg), C.byref(frames_c),
            priors.ctypes.data if priors is not None else None, results.ctypes.data, _p(n_stereo, T.c_int32_p)))

    def track_stereo_sequence(self, cam, cfg, mcfg, scfg, frames: dict, priors=None, results=None):
        """NF consecutive frames of raw stereo features -> NF - 1 poses (pair p = frames p, p + 1); every frame goes through the
        stereo step once.  Returns (results, n_stereo[NF, 2])."""
        fs, keep = T.stereo_features_as_c(frames)
        NF = fs.B
        if results is None:
            results = np.zeros(max(NF - 1, 0), dtype=T.POSE_RESULT_DTYPE)
        n_stereo = np.zeros((NF, 2), np.int32)
        self._ck(self.lib.plstvo_track_stereo_sequence(self.ctx, C.byref(cam), C.byref(cfg), C.byref(mcfg), C.byref(scfg),
                                                       C.byref(fs), priors.ctypes.data if priors is not None else None,
                                                       results.ctypes.data, _p(n_stereo, T.c_int32_p)))
        del keep
        return results, n_stereo

    # ---- stereoFrameHandler.h surface ----
    def f2f_tracking(self, cfg, prev: T.FrameBatch, curr: T.FrameBatch):
        m12_pt, m12_ls = np.full(prev.n_pt, -1, np.int32), np.full(prev.n_ls, -1, np.int32)
        n = np.zeros((prev.B, 2), np.int32)
        pc, cc = prev.as_c(), curr.as_c()
        self._ck(self.lib.plstvo_f2f_tracking(self.ctx, C.byref(cfg), C.byref(pc), C.byref(cc),
                                              _p(m12_pt, T.c_int32_p), _p(m12_ls, T.c_int32_p), _p(n, T.c_int32_p)))
        return m12_pt, m12_ls, n

    def optimize_pose(self, cam, cfg, matched: T.MatchedBatch, priors=None):
        res = np.zeros(matched.B, dtype=T.POSE_RESULT_DTYPE)
        inl_pt = np.zeros(int(matched.pt_off[-1]), np.uint8)
        inl_ls = np.zeros(int(matched.ls_off[-1]), np.uint8)
        mc = matched.as_c()
        self._ck(self.lib.plstvo_optimize_pose(self.ctx, C.byref(cam), C.byref(cfg), C.byref(mc),
                                               priors.ctypes.data if priors is not None else None, res.ctypes.data,
                                               _p(inl_pt, T.c_uint8_p), _p(inl_ls, T.c_uint8_p)))
        return res, inl_pt, inl_ls

    def track_batch(self, cam, cfg, prev: T.FrameBatch, curr: T.FrameBatch, priors=None, out=None):
        """insertStereoPair's f2fTracking + optimizePose for B pairs, host buffers in / out."""
        if out is None:
            out = dict(results=np.zeros(prev.B, dtype=T.POSE_RESULT_DTYPE),
                       m12_pt=np.full(prev.n_pt, -1, np.int32), m12_ls=np.full(prev.n_ls, -1, np.int32),
                       inlier_pt=np.zeros(prev.n_pt, np.uint8), inlier_ls=np.zeros(prev.n_ls, np.uint8))
        pc, cc = prev.as_c(), curr.as_c()
        self._ck(self.lib.plstvo_track_batch(
            self.ctx, C.byref(cam), C.byref(cfg), C.byref(pc), C.byref(cc),
            priors.ctypes.data if priors is not None else None, out["results"].ctypes.data,
            _p(out["m12_pt"], T.c_int32_p), _p(out["m12_ls"], T.c_int32_p), _p(out["inlier_pt"], T.c_uint8_p),
            _p(out["inlier_ls"], T.c_uint8_p)))
        return out

    def track_batch_async(self, cam, cfg, prev: T.FrameBatch, curr: T.FrameBatch, out, priors=None) -> int:
        """Streaming form: returns a ticket; `prev`, `curr`, `priors` and `out` must stay alive and untouched until
        wait(ticket).  Use pinned arrays (self.pinned) so that the copies really are asynchronous."""
        pc, cc = prev.as_c(), curr.as_c()
        self._keep = getattr(self, "_keep", {})
        t = self._ck(self.lib.plstvo_track_batch_async(
            self.ctx, C.byref(cam), C.byref(cfg), C.byref(pc), C.byref(cc),
            priors.ctypes.data if priors is not None else None, out["results"].ctypes.data,
            _p(out["m12_pt"], T.c_int32_p), _p(out["m12_ls"], T.c_int32_p), _p(out["inlier_pt"], T.c_uint8_p),
            _p(out["inlier_ls"], T.c_uint8_p)))
        self._keep[t] = (prev, curr, priors, out)
        return t

    def wait(self, ticket: int):
        self._ck(self.lib.plstvo_wait(self.ctx, ticket))
        getattr(self, "_keep", {}).pop(ticket, None)

    def pinned_outputs(self, prev: T.FrameBatch):
        e = self.pinned.empty
        return dict(results=e((prev.B,), T.POSE_RESULT_DTYPE), m12_pt=e((prev.n_pt,), np.int32),
                    m12_ls=e((prev.n_ls,), np.int32), inlier_pt=e((prev.n_pt,), np.uint8),
                    inlier_ls=e((prev.n_ls,), np.uint8))

    def upload(self, cam, cfg, prev: T.FrameBatch, curr: T.FrameBatch, priors=None) -> DeviceBatch:
        h = C.c_void_p()
        pc, cc = prev.as_c(), curr.as_c()
        self._ck(self.lib.plstvo_batch_upload(self.ctx, C.byref(cam), C.byref(cfg), C.byref(pc), C.byref(cc),
                                              priors.ctypes.data if priors is not None else None, C.byref(h)))
        return DeviceBatch(self, h, prev)

    def gn_eval_stream(self, cam, cfg, matched: T.MatchedBatch, DT, iters: int = 1):
        """optimizeFunctions for B problems streamed from HBM (C5 roofline kernel); returns H, g, e, ms_total."""
        B = matched.B
        DT = np.ascontiguousarray(DT, np.float64).reshape(B, 16)
        H, g, e = np.zeros((B, 6, 6)), np.zeros((B, 6)), np.zeros(B)
        ms = C.c_float(0)
        mc = matched.as_c()
        self._ck(self.lib.plstvo_gn_eval_stream(self.ctx, C.byref(cam), C.byref(cfg), C.byref(mc), _p(DT, T.c_double_p),
                                                iters, _p(H, T.c_double_p), _p(g, T.c_double_p), _p(e, T.c_double_p),
                                                C.byref(ms)))
        return H, g, e, float(ms.value)

    def debug_select(self, lists, ks, pivots=None):
        """Block-wide radix selection (test hook): the ks[p]-th smallest of lists[p]; with pivots, of |x - pivot| rounded to float."""
        off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int32)
        v = np.ascontiguousarray(np.concatenate([np.asarray(x, np.float64) for x in lists]) if len(lists) else np.zeros(0))
        ks = np.ascontiguousarray(ks, np.int32)
        out = np.zeros(len(lists))
        pv = None if pivots is None else np.ascontiguousarray(pivots, np.float64)
        self._ck(self.lib.plstvo_debug_select(self.ctx, len(lists), _p(off, T.c_int32_p), _p(v, T.c_double_p), _p(ks, T.c_int32_p),
                                              _p(pv, T.c_double_p) if pv is not None else None, 0 if pv is None else 1,
                                              _p(out, T.c_double_p)))
        return out

    def debug_algebra(self, H, g):
        """On-chip 6x6 routines (test hook): returns x = H^-1 g (col-pivot QR), log|det H|, inv(H), eigvals(H)."""
        H = np.ascontiguousarray(H, np.float64).reshape(-1, 36)
        g = np.ascontiguousarray(g, np.float64).reshape(-1, 6)
        n = len(H)
        x, lad, inv, eig = np.zeros((n, 6)), np.zeros(n), np.zeros((n, 36)), np.zeros((n, 6))
        self._ck(self.lib.plstvo_debug_algebra(self.ctx, n, _p(H, T.c_double_p), _p(g, T.c_double_p), _p(x, T.c_double_p),
                                               _p(lad, T.c_double_p), _p(inv, T.c_double_p), _p(eig, T.c_double_p)))
        return x, lad, inv.reshape(n, 6, 6), eig

    def synchronize(self):
        self._ck(self.lib.plstvo_synchronize(self.ctx))

    def popc_rate(self) -> float:
        r = np.zeros(1)
        self._ck(self.lib.plstvo_popc_rate(self.ctx, r.ctypes.data_as(T.c_double_p)))
        return float(r[0])
