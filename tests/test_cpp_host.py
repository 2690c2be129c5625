"""The C++ host layer (include/plstvo.hpp) compiles against the C-ABI and links the product library (CPU: link check
only); on the GPU box the example runs the reference's call sequence on one pair and must agree with the oracle."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import ref_numpy as R
from conftest import ROOT
from stvo_pl_b200 import synth, types as T

LIBDIR = os.path.join(ROOT, "stvo_pl_b200", "lib")


def build_example(tmp_path):
    exe = str(tmp_path / "track_cpp")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "track_cpp.cpp"), "-L", LIBDIR, "-lplstvo_b200",
                    f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True)
    return exe


def test_cpp_host_compiles_and_links(tmp_path):
    exe = build_example(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert json.loads(out)["version"] == 100


def _write_vec(f, a, dtype):
    a = np.ascontiguousarray(a, dtype).ravel()
    f.write(struct.pack("<q", a.size))
    f.write(a.tobytes())


def _write_frame(f, fb: T.FrameBatch):
    _write_vec(f, fb.pdesc, np.uint8); _write_vec(f, fb.ldesc, np.uint8)
    _write_vec(f, fb.pt_P, np.float64); _write_vec(f, fb.pt_pl, np.float64); _write_vec(f, fb.pt_sigma2, np.float64)
    for k in ("ls_sP", "ls_eP", "ls_le", "ls_spl", "ls_epl", "ls_sigma2"):
        _write_vec(f, getattr(fb, k), np.float64)
    _write_vec(f, fb.ls_level, np.int32)


@pytest.mark.gpu
def test_cpp_handler_sequence_vs_oracle(tmp_path, oracle):
    exe = build_example(tmp_path)
    prev, curr, _, cam = synth.make_batch("kitti", 1, n_pt=700, n_ls=150)
    path = str(tmp_path / "pair.bin")
    with open(path, "wb") as f:
        f.write(bytes(cam))
        _write_frame(f, prev)
        _write_frame(f, curr)
    got = json.loads(subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout)
    ref = oracle.track_batch(cam, T.kitti_config(), prev, curr, priors=T.identity_priors(1))["results"][0]
    assert got["good"] == ref["good"] == 1 and got["status"] == ref["status"]
    assert got["n_matched_pt"] == ref["n_matched_pt"] and got["n_inliers"] == ref["n_inliers"]
    ang, tr = R.pose_error(np.array(got["DT"]).reshape(4, 4), ref["DT"])
    assert ang < 1e-9 and tr < 1e-8


@pytest.mark.gpu
def test_cpp_stereo_sequence_vs_fused_call(tmp_path):
    """examples/stereo_cpp.cpp: raw stereo features -> matchStereoPoints / Lines per frame -> handler, in C++; must agree with
    plstvo_track_stereo_batch on the same two frames."""
    from stvo_pl_b200 import stereo_synth as SS
    from stvo_pl_b200.engine import Engine
    exe = str(tmp_path / "stereo_cpp")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "stereo_cpp.cpp"), "-L", LIBDIR, "-lplstvo_b200", f"-Wl,-rpath,{LIBDIR}",
                    "-o", exe], check=True)
    prev, curr, Tgt, cam = SS.make_stereo_pairs(1, n_pt=800, n_ls=160, seed=5)
    path = str(tmp_path / "stereo.bin")
    with open(path, "wb") as f:
        f.write(bytes(cam))
        for d in (prev, curr):
            for key, dt in (("kp_l", np.float32), ("poct_l", np.int32), ("pdesc_l", np.uint8), ("kp_r", np.float32),
                            ("pdesc_r", np.uint8), ("seg_l", np.float32), ("angle_l", np.float32), ("loct_l", np.int32),
                            ("ldesc_l", np.uint8), ("seg_r", np.float32), ("ldesc_r", np.uint8)):
                _write_vec(f, d[key], dt)
    got = json.loads(subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout)
    eng = Engine(0)
    res, n_st = eng.track_stereo_batch(cam, T.kitti_config(), T.default_stereo_match_config(), T.default_stereo_config(), prev, curr)
    eng.close()
    assert got["stereo"] == [int(x) for x in n_st[0]]
    assert got["good"] == int(res["good"][0]) == 1 and got["n_matched_pt"] == int(res["n_matched_pt"][0])
    assert got["n_inliers"] == int(res["n_inliers"][0])
    ang, tr = R.pose_error(np.array(got["DT"]).reshape(4, 4), res["DT"][0])
    assert ang < 1e-12 and tr < 1e-12
