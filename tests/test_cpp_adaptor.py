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


@pytest.mark.gpu
def test_adaptor_bundle_adjuster_refines_intrinsics(oracle, tmp_path):
    """BundleAdjustmentOptions::refine_extra_params + config.SetConstantCamera through the C++
    adaptor (ParameterizeCameras, bundle_adjustment.cc:490-528): camera.ParamsData() of the
    variable camera is updated in place, the constant camera stays."""
    exe = _build()
    sc = S.make_ba_scene(num_cams=7, num_points=200, obs_per_point=4, seed=81, noise_px=2.0)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([7, 200, len(sc["obs_cam"])], dtype=np.int64).tofile(f)
        for a in (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"].astype(np.float64),
                  sc["obs_pt"].astype(np.float64), sc["obs_line"], sc["cam_params"]):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)
    subprocess.check_call([exe, "ba_refine", fin, fout])
    out = np.fromfile(fout, dtype=np.float64)
    ok, c0, c1, iters = out[:4]
    q = out[4:4 + 28].reshape(7, 4)
    t = out[32:32 + 21].reshape(7, 3)
    X = out[53:53 + 600].reshape(200, 3)
    cams = out[653:661].reshape(2, 4)
    flags = np.zeros(7, np.uint8)
    flags[0], flags[1] = 1, 2
    cp = sc["cam_params"]
    prm = [cp[0], cp[2], cp[3], 0.08]
    b = oracle.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                        sc["obs_line"], [2, 2], [prm, prm], image_camera=np.arange(7) % 2,
                        pose_flags=flags, camera_const=[0, 1])
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(
        num_threads=1, max_num_iterations=6, gradient_tolerance=1e-4, refine_extra_params=1))
    assert bool(ok) and ok2
    assert abs(c1 - s2.final_cost) <= 1e-8 * s2.final_cost
    assert int(iters) == s2.num_successful_steps + s2.num_unsuccessful_steps
    assert np.abs(q - b.qvecs).max() < 1e-7 and np.abs(t - b.tvecs).max() < 1e-7
    assert np.abs(X - b.points).max() < 1e-6
    assert np.array_equal(cams[1], prm) and cams[0, 3] != 0.08
    assert np.array_equal(cams[0, :3], prm[:3])
    assert abs(cams[0, 3] - b.camera_params[0, 3]) <= 1e-6 * max(1e-3, abs(b.camera_params[0, 3]))


def _scene_file(tmp_path, sc, aligned=None):
    fin = str(tmp_path / "scene.bin")
    with open(fin, "wb") as f:
        np.array([sc["num_cams"], sc["num_points"], len(sc["obs_cam"])], dtype=np.int64).tofile(f)
        for a in (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"].astype(np.float64),
                  sc["obs_pt"].astype(np.float64), sc["obs_line"], sc["cam_params"]):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)
        if aligned is not None:
            np.ascontiguousarray(aligned, dtype=np.float64).tofile(f)
    return fin


def _filter_scene(seed):
    """Point-major scene with corrupted lines, far / behind points and all-aligned tracks."""
    from privacy_preserving_sfm_b200 import filters as F
    sc = S.make_ba_scene(num_cams=10, num_points=300, obs_per_point=6, seed=seed)
    rng = np.random.default_rng(seed + 1)
    O = len(sc["obs_cam"])
    sc["qvecs"], sc["tvecs"] = sc["qvecs_gt"], sc["tvecs_gt"]
    pts = sc["points_gt"].copy()
    line = sc["obs_line"].copy()
    bad = rng.uniform(size=O) < 0.12
    line[bad, 2] += rng.normal(scale=0.02, size=bad.sum())
    pts[rng.choice(300, 12, replace=False)] *= 400.0
    pts[rng.choice(300, 30, replace=False)] += 30.0 * np.array([0.0, 0.0, 1.0])
    aligned = (rng.uniform(size=O) < 0.4).astype(np.uint8)
    track_start = np.searchsorted(sc["obs_pt"], np.arange(301)).astype(np.int64)
    for p in rng.choice(300, 8, replace=False):
        aligned[track_start[p]:track_start[p + 1]] = 1
    sc["points"], sc["obs_line"] = pts, line
    pb = F.FilterProblem(sc["qvecs"], sc["tvecs"], np.zeros(10, np.int32), [1],
                         [list(sc["cam_params"])], [(1000, 1000)], pts, track_start,
                         sc["obs_cam"], line, aligned)
    return sc, aligned, pb


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["filter", "depth"])
def test_adaptor_filters_match_oracle(oracle, tmp_path, mode):
    """ppsfm::FilterPoints3D / FilterObservationsWithNegativeDepth applied to a Reconstruction-like
    object through DeleteObservation / DeletePoint3D / SetError end in the state the oracle's
    masks describe."""
    exe = _build()
    sc, aligned, pb = _filter_scene(41)
    fin, fout = _scene_file(tmp_path, sc, aligned), str(tmp_path / "out.bin")
    subprocess.check_call([exe, mode, fin, fout])
    out = np.fromfile(fout, dtype=np.float64)
    P, O = 300, len(sc["obs_cam"])
    nf, thr = out[0], out[1]
    alive, err = out[2:2 + 2 * P:2], out[3:3 + 2 * P:2]
    attached = out[2 + 2 * P:]
    assert thr == 12.0 / 1000.0                       # ImageToWorldThreshold, PINHOLE f = 1000
    if mode == "filter":
        nf2, od2, pd2, pe2, _ = oracle.filter_points3d(pb, 4.0, 1.5)
        assert np.array_equal(err[alive > 0], pe2[pd2 == 0])
    else:
        nf2, od2, pd2 = oracle.filter_negative_depth(pb)
    assert int(nf) == nf2 and 0 < nf2
    assert np.array_equal(alive > 0, pd2 == 0) and 0 < (pd2 == 0).sum() < P
    assert np.array_equal(attached > 0, od2 == 0)


@pytest.mark.gpu
def test_adaptor_triangulation_matches_oracle(oracle, tmp_path, ctx):
    """ppsfm::EstimateTriangulationBatch / EstimateTriangulation (projection matrices in, the
    reference's signature) against the oracle's EstimateTriangulation on the same tracks."""
    from privacy_preserving_sfm_b200 import filters as F, triangulation as T
    exe = _build()
    sc = S.make_ba_scene(num_cams=9, num_points=150, obs_per_point=6, seed=52)
    sc["qvecs"], sc["tvecs"] = sc["qvecs_gt"], sc["tvecs_gt"]
    rng = np.random.default_rng(3)
    bad = rng.uniform(size=len(sc["obs_cam"])) < 0.1
    sc["obs_line"] = sc["obs_line"].copy()
    sc["obs_line"][bad, 2] += rng.normal(scale=0.05, size=bad.sum())
    fin, fout = _scene_file(tmp_path, sc), str(tmp_path / "out.bin")
    subprocess.check_call([exe, "tri", fin, fout])
    out = np.fromfile(fout, dtype=np.float64)
    assert out[1] == 1.0                              # per-track call == batch on track 0
    P, O = 150, len(sc["obs_cam"])
    ok, xyz = out[2:2 + 4 * P].reshape(P, 4)[:, 0] > 0, out[2:2 + 4 * P].reshape(P, 4)[:, 1:]
    mask = out[2 + 4 * P:] > 0
    track_start = np.searchsorted(sc["obs_pt"], np.arange(P + 1)).astype(np.int64)
    pb = F.FilterProblem(sc["qvecs"], sc["tvecs"], np.zeros(9, np.int32), [1],
                         [list(sc["cam_params"])], [(1000, 1000)], sc["points_gt"], track_start,
                         sc["obs_cam"], sc["obs_line"], np.zeros(O, np.uint8))
    opt = T.EstimateTriangulationOptions(
        min_tri_angle=np.deg2rad(1.5), residual_type=T.ANGULAR_ERROR, max_error=np.deg2rad(2.0),
        confidence=0.9999, min_inlier_ratio=0.02, max_num_trials=10000, exhaustive_threshold=15)
    ok2, xyz2, mask2, _ = oracle.estimate_triangulation_batch(pb, opt)
    assert np.array_equal(ok, ok2) and ok.sum() > 100
    assert np.array_equal(mask[np.repeat(ok, 6)], mask2[np.repeat(ok, 6)])
    # (rotation matrix -> quaternion -> rotation matrix on the way in: 1e-8, not bits)
    assert np.abs(xyz[ok] - xyz2[ok]).max() < 1e-8 * max(1.0, np.abs(xyz2[ok]).max())
    assert np.median(np.abs(xyz[ok] - sc["points_gt"][ok]).max(axis=1)) < 0.02
