"""Four-view initialisation (BASELINE.json configs[0]) — host code, runs
without a GPU.  Pins:
  * the LO-MSAC driver (ppsfm::LocallyOptimizedMSAC) bit-for-bit against the REFERENCE's own
    ransac_lib::LocallyOptimizedMSAC (oracle/_ref/libref_init.so, compiled from the reference's
    header where it lies) on the same solver objects;
  * the estimators against the properties the reference's own tests check
    (src/init/initializer_test.cc:346-434: poses recovered to 1e-6 without / 1e-4 with outliers).
"""
import ctypes as C
import os

import numpy as np
import pytest

from privacy_preserving_sfm_b200 import initializer as I
from privacy_preserving_sfm_b200 import synthetic as S

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref",
                    "libref_init.so")


def _normalised(poses):
    p = poses.copy()
    p[:, :, 3] /= np.linalg.norm(p[1, :, 3])   # "normalize w.r.t. second pose"
    return p


def _relative(poses):
    """P_i P_0^-1 (invariant to the choice of the world frame), translation scale |t_1| = 1."""
    R0, t0 = poses[0, :, :3], poses[0, :, 3]
    out = []
    for P in poses:
        R = P[:, :3] @ R0.T
        out.append(np.concatenate([R, (P[:, 3] - R @ t0)[:, None]], axis=1))
    out = np.stack(out)
    out[:, :, 3] /= np.linalg.norm(out[1, :, 3])
    return out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_initializer_no_outliers(seed):
    lines, aligned, gravity, gt = S.make_init_scene(100, 50, 0, seed=seed)
    ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, gravity)
    assert ok and rep.num_aligned == 50 and rep.num_unaligned == 50
    assert rep.inliers_2d == 50 and rep.inliers_3d == 50 and ratio == 1.0
    assert rep.iterations_2d >= 1000      # min_num_iterations_ (initializer.cc:112)
    est = _normalised(poses)
    for i in range(4):
        assert np.linalg.norm(est[i] - gt[i]) < 1e-6   # initializer_test.cc:375


@pytest.mark.parametrize("seed", [4, 5])
def test_initializer_with_outliers(seed):
    lines, aligned, gravity, gt = S.make_init_scene(100, 50, 10, seed=seed)
    ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, gravity)
    assert ok and 0.7 <= ratio <= 1.0
    est = _normalised(poses)
    for i in range(4):
        assert np.linalg.norm(est[i] - gt[i]) < 1e-4   # initializer_test.cc:427


def test_initializer_tilted_cameras():
    """Gravity not along the camera y axis: exercises the gravity pre-rotation (initializer.cc:72-95)."""
    lines, aligned, gravity, gt = S.make_init_scene(120, 60, 0, seed=7, tilt_deg=12.0)
    ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, gravity)
    assert ok
    assert np.abs(_relative(poses) - _relative(gt)).max() < 1e-6


def test_config1_plumbing():
    """BASELINE.json configs[0]: 2 000 lifted-line tracks (1 000 aligned + 1 000 random), 10 %
    outliers, InitOptions{max_error 0.005}."""
    lines, aligned, gravity, gt = S.make_init_scene(2000, 1000, 200, seed=S.SCENE_SEED)
    ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, gravity,
                                                       I.InitOptions(max_error=0.005))
    assert ok and rep.num_aligned == 1000 and ratio > 0.8
    assert np.abs(_normalised(poses) - gt).max() < 1e-4


def test_contract_violations():
    lines, aligned, gravity, _ = S.make_init_scene(40, 20, 0, seed=9)
    bad = aligned.copy()
    bad[2, :] = 1 - bad[2, :]            # aligned / unaligned split differs between images
    with pytest.raises(ValueError):
        I.initialize_reconstruction(lines, bad, gravity)
    g2 = gravity.copy()
    g2[1] = [1.0, 0.0, 0.0]               # aligned lines no longer parallel to "gravity"
    with pytest.raises(ValueError):
        I.initialize_reconstruction(lines, aligned, g2)
    # too few tracks for a minimal sample: the reference returns false
    idx = np.flatnonzero(aligned[0])[:4]
    ok, _, _, _ = I.initialize_reconstruction(lines[:, idx], aligned[:, idx], gravity)
    assert not ok


@pytest.mark.skipif(not os.path.exists(_REF), reason="oracle/_ref/libref_init.so not built")
@pytest.mark.parametrize("n,n_al,n_out,seed", [(100, 50, 0, 11), (100, 50, 10, 12),
                                                (300, 120, 45, 13), (60, 40, 20, 14)])
def test_lomsac_bit_identical_to_reference_driver(n, n_al, n_out, seed):
    """Same estimators under the reference's ransac_lib::LocallyOptimizedMSAC: every output must
    be bit-identical (the driver consumes the same mt19937 draws and applies the same rules)."""
    lines, aligned, gravity, _ = S.make_init_scene(n, n_al, n_out, seed=seed)
    ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, gravity)
    ref = C.CDLL(_REF)
    dp = C.POINTER(C.c_double)
    ref.ref_initialize_reconstruction.argtypes = [dp, C.POINTER(C.c_uint8), C.c_size_t, dp, dp,
                                                  dp, dp, dp]
    opt = np.array([0.1, 6.0, 0.005])
    poses2 = np.zeros((4, 3, 4))
    ratio2 = C.c_double(0.0)
    rep2 = np.zeros(7)
    rc = ref.ref_initialize_reconstruction(
        np.ascontiguousarray(lines).ctypes.data_as(dp),
        np.ascontiguousarray(aligned).ctypes.data_as(C.POINTER(C.c_uint8)), n,
        np.ascontiguousarray(gravity).ctypes.data_as(dp), opt.ctypes.data_as(dp),
        poses2.ctypes.data_as(dp), C.byref(ratio2), rep2.ctypes.data_as(dp))
    assert (rc == 0) == ok
    assert [rep.inliers_2d, rep.inliers_3d, rep.iterations_2d, rep.iterations_3d] == \
        [int(v) for v in rep2[2:6]]
    assert ratio == ratio2.value
    assert np.array_equal(poses, poses2)


@pytest.mark.parametrize("m,n", [(3, 2), (4, 3)])
def test_fixed_size_least_squares_equals_the_generic_routine(m, n):
    """hd::qr_solve_fixed (cpp/ppsfm_init_math.h: what the triangulations and the GPU scoring
    kernel run) is la::qr_solve operation for operation: same bits on random, rank-deficient and
    badly scaled systems."""
    from privacy_preserving_sfm_b200 import binding
    L = binding.load_library()
    dp = C.POINTER(C.c_double)
    L.ppsfm_init_test_qr.argtypes = [dp, C.c_int, C.c_int, dp, dp, dp]
    rng = np.random.default_rng(m * 10 + n)
    for k in range(500):
        A = rng.normal(size=(m, n))
        b = rng.normal(size=m)
        if k % 5 == 1:
            A[:, 1] = 2.0 * A[:, 0]                 # rank deficient
        if k % 5 == 2:
            A *= 10.0 ** rng.integers(-8, 8)        # badly scaled
        if k % 5 == 3:
            A[:, n - 1] *= 1e-15                    # negligible pivot
        xg, xf = np.zeros(n), np.zeros(n)
        rc = L.ppsfm_init_test_qr(np.ascontiguousarray(A).ctypes.data_as(dp), m, n,
                                  b.ctypes.data_as(dp), xg.ctypes.data_as(dp),
                                  xf.ctypes.data_as(dp))
        assert rc == 0 and np.array_equal(xg, xf)
        if k % 5 == 0:
            assert np.allclose(xg, np.linalg.lstsq(A, b, rcond=None)[0], atol=1e-9)


@pytest.mark.parametrize("n,n_al,n_out,seed", [(100, 50, 10, 4), (160, 90, 30, 15)])
def test_batched_scoring_hooks_reproduce_the_plain_run(tmp_path, n, n_al, n_out, seed):
    """The optional Solver members ScoreModels / Materialize (lazy candidate models scored in one
    batch — what csrc/init_kernels.cu does on the GPU) with a HOST scorer built on the shared
    arithmetic of cpp/ppsfm_init_math.h: initialize_reconstruction must return exactly what the
    plain run returns (tests/cpp/init_scorer_test.cc; no GPU involved)."""
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    exe = str(tmp_path / "init_scorer_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall",
                           "-I" + os.path.join(root, "privacy_preserving_sfm_b200", "cpp"),
                           os.path.join(root, "tests", "cpp", "init_scorer_test.cc"), "-o", exe])
    lines, aligned, gravity, _ = S.make_init_scene(n, n_al, n_out, seed=seed)
    blob = np.concatenate([np.ascontiguousarray(lines, np.float64).ravel(),
                           np.ascontiguousarray(aligned, np.float64).ravel(),
                           np.ascontiguousarray(gravity, np.float64).ravel()]).tobytes()
    r = subprocess.run([exe, str(n)], input=blob, capture_output=True)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
    assert b"identical 1" in r.stdout
