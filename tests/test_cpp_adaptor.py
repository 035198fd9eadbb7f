"""The header-only C++ adaptor (privacy_preserving_sfm_b200/cpp/ppsfm_adaptor.h) that re-creates
the reference API: it must compile against the C-ABI without Eigen (CPU test) and reproduce the
oracle's results when run (GPU test)."""
import os
import subprocess

import numpy as np
import pytest

import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "adaptor_test")


def _build():
    pp.build_library()
    libdir = os.path.join(ROOT, "privacy_preserving_sfm_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(libdir, "cpp"), os.path.join(ROOT, "tests", "cpp", "adaptor_test.cc"),
           "-o", BIN, "-L" + libdir, "-lppsfm_b200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return BIN


def test_adaptor_compiles_without_eigen():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_adaptor_pose_matches_oracle(oracle, tmp_path):
    exe = _build()
    sc = S.make_abs_pose_scene(n=3000, inlier_ratio=0.4, seed=33)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([3000], dtype=np.int64).tofile(f)
        sc["lines"].tofile(f)
        sc["points"].tofile(f)
        sc["aligned"].astype(np.float64).tofile(f)
    subprocess.check_call([exe, "pose", fin, fout])
    out = np.fromfile(fout, dtype=np.float64)
    ok, ninl, ok_ref, n_models = out[:4]
    q, t, q_ref, t_ref, mask = out[4:8], out[8:11], out[11:15], out[15:18], out[18:]
    oracle.set_prng_seed(0)
    o = oracle.make_options(0.012, 0.25, 0.99999, 3.0, 100, 10000)
    ok2, q2, t2, ninl2, mask2, _ = oracle.estimate_absolute_pose_from_lines(
        sc["lines"], sc["aligned"], sc["points"], o)
    assert bool(ok) == ok2 and int(ninl) == ninl2
    assert np.array_equal(q, q2) and np.array_equal(t, t2)          # bit-exact
    assert np.array_equal(mask.astype(np.uint8), mask2)
    okr, qr, tr, _ = oracle.refine_absolute_pose(sc["lines"], sc["points"], mask2, 1,
                                                 [1000.0, 1000.0, 500.0, 500.0], q2, t2)
    assert bool(ok_ref) == okr
    assert np.abs(q_ref - qr).max() < 1e-9 and np.abs(t_ref - tr).max() < 1e-9
    ref6 = oracle.p6l_estimate(sc["lines"][:6], sc["aligned"][:6], sc["points"][:6])
    assert int(n_models) == len(ref6)


@pytest.mark.gpu
def test_adaptor_bundle_adjuster_matches_oracle(oracle, tmp_path):
    exe = _build()
    sc = S.make_ba_scene(num_cams=7, num_points=200, obs_per_point=4, seed=81)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([7, 200, len(sc["obs_cam"])], dtype=np.int64).tofile(f)
        for a in (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"].astype(np.float64),
                  sc["obs_pt"].astype(np.float64), sc["obs_line"], sc["cam_params"]):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)
    subprocess.check_call([exe, "ba", fin, fout])
    out = np.fromfile(fout, dtype=np.float64)
    ok, c0, c1, iters = out[:4]
    q = out[4:4 + 28].reshape(7, 4)
    t = out[32:32 + 21].reshape(7, 3)
    X = out[53:].reshape(200, 3)
    flags = np.zeros(7, np.uint8)
    flags[0], flags[1] = 1, 2
    b = oracle.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                        sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags)
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(num_threads=1, max_num_iterations=20,
                                                           gradient_tolerance=1e-4))
    assert bool(ok) and ok2
    assert abs(c1 - s2.final_cost) <= 1e-9 * s2.final_cost
    assert int(iters) == s2.num_successful_steps + s2.num_unsuccessful_steps
    assert np.abs(q - b.qvecs).max() < 1e-8 and np.abs(t - b.tvecs).max() < 1e-8
    assert np.abs(X - b.points).max() < 1e-7
