"""The reference-side shim as FILES (shim/matching_b200.cpp, shim/stereoFrameHandler_b200.cpp): compiled against stand-in
headers that carry the reference's class and member names (shim/standin/, Eigen surface from oracle/ref_shim/), linked
against libplstvo_b200.so, and — on the GPU box — driven exactly like app/imagesStVO.cpp:96-97 drives the handler."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from stvo_pl_b200 import synth, types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "shim", "_build", "libshim_test.so")
SRC = [os.path.join(ROOT, "shim", n) for n in ("matching_b200.cpp", "stereoFrameHandler_b200.cpp", os.path.join("standin", "shim_driver.cpp"))]


def build_shim() -> str:
    from stvo_pl_b200 import build as b
    lib = b.build()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = SRC + [os.path.join(ROOT, "shim", "standin", n) for n in ("matching.h", "stereoFrameHandler.h", "config.h")] + [lib]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-std=c++11", "-O1", "-fPIC", "-shared", "-Wall", "-Wno-unused", "-Wno-ignored-qualifiers", "-Werror=return-type",
           "-I", os.path.join(ROOT, "shim", "standin"), "-I", os.path.join(ROOT, "include"), "-o", OUT] + SRC + \
          ["-L", os.path.dirname(lib), "-lplstvo_b200", "-Wl,-rpath," + os.path.dirname(lib), "-Wl,--allow-shlib-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return OUT


def test_shim_compiles_and_links_against_the_library():
    so = build_shim()
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    for name in ("_ZN4StVO8matchNNRERKN2cv3MatES3_fRSt6vectorIiSaIiEE", "_ZN4StVO5matchERKN2cv3MatES3_fRSt6vectorIiSaIiEE",
                 "_ZN4StVO18StereoFrameHandler11f2fTrackingEv", "_ZN4StVO18StereoFrameHandler12optimizePoseEv"):
        assert name in syms, name
    undefined = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout
    for name in ("plstvo_create", "plstvo_match_nnr", "plstvo_match", "plstvo_track_batch", "plstvo_last_error"):
        assert name in undefined, name      # resolved by libplstvo_b200.so at load time


@pytest.mark.gpu
def test_shim_drives_the_engine_like_the_reference_app(engine, oracle):
    L = C.CDLL(build_shim())
    cfg = T.kitti_config()
    prev, curr, Tgt, cam = synth.make_batch("kitti", 3, n_pt=700, n_ls=180)
    direct = engine.track_batch(cam, cfg, prev, curr)
    for p in range(3):
        a, b = prev.select([p]), curr.select([p])
        ac, bc = a.as_c(), b.as_c()
        DT, cov, err, Tfw = np.zeros(16), np.zeros(36), C.c_double(0), np.zeros(16)
        ninl, nmp, nml = np.zeros(3, np.int32), C.c_int32(0), C.c_int32(0)
        ip, il = np.zeros(a.n_pt + 1, np.uint8), np.zeros(a.n_ls + 1, np.uint8)
        dp = lambda x: x.ctypes.data_as(T.c_double_p)
        rc = L.shim_track_pair(C.byref(cam), C.byref(cfg), C.byref(ac), C.byref(bc), dp(DT), dp(cov), C.byref(err), dp(Tfw),
                               ninl.ctypes.data_as(T.c_int32_p), C.byref(nmp), C.byref(nml), ip.ctypes.data_as(T.c_uint8_p),
                               il.ctypes.data_as(T.c_uint8_p))
        assert rc == 0
        r = direct["results"][p]
        np.testing.assert_array_equal(DT.reshape(4, 4), r["DT"].reshape(4, 4))       # same library call underneath: bit-identical
        np.testing.assert_array_equal(cov.reshape(6, 6), r["DT_cov"].reshape(6, 6))
        assert err.value == r["err_norm"]
        assert (nmp.value, nml.value) == (r["n_matched_pt"], r["n_matched_ls"])
        assert tuple(ninl) == (r["n_inliers_pt"], r["n_inliers_ls"], r["n_inliers"])
        sel = direct["m12_pt"][prev.pt_off[p]:prev.pt_off[p + 1]] >= 0                # matched_pt is in ascending i1
        np.testing.assert_array_equal(ip[:nmp.value], direct["inlier_pt"][prev.pt_off[p]:prev.pt_off[p + 1]][sel])
        sel = direct["m12_ls"][prev.ls_off[p]:prev.ls_off[p + 1]] >= 0
        np.testing.assert_array_equal(il[:nml.value], direct["inlier_ls"][prev.ls_off[p]:prev.ls_off[p + 1]][sel])
    # matching.h surface
    d1 = prev.pdesc[:500].copy()
    d2 = curr.pdesc[:480].copy()
    for mutual in (0, 1):
        m12 = np.full(500, -7, np.int32)
        n = L.shim_match(d1.ctypes.data_as(T.c_uint8_p), 500, d2.ctypes.data_as(T.c_uint8_p), 480, C.c_float(0.75), mutual,
                         m12.ctypes.data_as(T.c_int32_p))
        ref = oracle.match(d1, d2, 0.75, best_lr=True)[1] if mutual else oracle.match_nnr(d1, d2, 0.75)[1]
        np.testing.assert_array_equal(m12, ref)
        assert n == int((ref >= 0).sum())
