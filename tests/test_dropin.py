"""The drop-in boundary at its strictest: a caller written against the REFERENCE'S OWN headers.

privacy_preserving_sfm_b200/cpp/dropin/estimators_pose_lines.cc defines
colmap::EstimateAbsolutePoseFromLines and colmap::RefineAbsolutePoseFromLines — the two functions of
src/estimators/pose.cc that IncrementalMapper::RegisterNextImage calls
(src/sfm/incremental_mapper.cc:719-735) — on top of libppsfm_b200.so.  It includes the reference's
src/estimators/pose.h, so the compiler checks the definitions against the reference's own
declarations and types (colmap::RANSACOptions, colmap::FeatureLines,
colmap::AbsolutePoseRefinementOptions, colmap::Camera).  tests/cpp/dropin_pose_test.cc is such a
caller; tests/cpp/build_dropin.sh compiles both, plus the reference's base/camera.cc and
base/camera_models.cc from where they lie, where /root/reference exists (Eigen / Ceres / glog /
Boost are the stand-in headers of oracle/ref/shim).  The binary travels to the GPU box."""
import os
import subprocess

import numpy as np
import pytest

from privacy_preserving_sfm_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "dropin_pose_test")
INIT_BIN = os.path.join(ROOT, "tests", "cpp", "dropin_init_test")
SCRIPT = os.path.join(ROOT, "tests", "cpp", "build_dropin.sh")


def test_dropin_compiles_against_the_reference_headers():
    if not os.path.isdir("/root/reference/src/estimators"):
        pytest.skip("/root/reference absent (the binary is prebuilt where it exists)")
    import privacy_preserving_sfm_b200 as pp
    pp.build_library()
    if os.path.exists(BIN):
        os.remove(BIN)
    subprocess.check_call(["bash", SCRIPT], stdout=subprocess.DEVNULL)
    assert os.path.exists(BIN)
    # the caller resolves the reference's symbols to the drop-in, and the drop-in to the C-ABI
    syms = subprocess.run(["nm", "-C", BIN], capture_output=True, text=True).stdout
    assert " T colmap::EstimateAbsolutePoseFromLines(" in syms
    assert " T colmap::RefineAbsolutePoseFromLines(" in syms
    assert " U ppsfm_estimate_absolute_pose_from_lines" in syms
    assert " U ppsfm_refine_absolute_pose_from_lines_ex" in syms
    # cpp/dropin/estimators_triangulation.cc and the GPU-scored build of init_initializer.cc are
    # compile-checked against the reference's declarations (triangulation.h:143-147,
    # initializer.h:103-108)
    tri = subprocess.run(["nm", "-C", os.path.join(ROOT, "tests", "cpp", "dropin_triangulation.o")],
                         capture_output=True, text=True).stdout
    assert " T colmap::EstimateTriangulation(" in tri
    ini = subprocess.run(["nm", "-C", os.path.join(ROOT, "tests", "cpp", "dropin_init_gpu.o")],
                         capture_output=True, text=True).stdout
    assert " T colmap::init::initialize_reconstruction(" in ini
    assert " U ppsfm_initialize_reconstruction_gpu" in ini


def test_dropin_initializer_runs_the_references_own_test_recipe():
    """A caller of the reference's init::initialize_reconstruction (its own header and types)
    linked with cpp/dropin/init_initializer.cc, on the recipe of src/init/initializer_test.cc
    (:346-434): poses to 1e-6 without outliers, 1e-4 with 10 % outliers, and the same bits as the
    library's ppsfm_initialize_reconstruction (both run cpp/ppsfm_init.h on the host)."""
    if not os.path.exists(INIT_BIN):
        pytest.skip("tests/cpp/dropin_init_test not prebuilt (needs /root/reference to build)")
    from privacy_preserving_sfm_b200 import initializer as I
    for (n, n_al, n_out, seed), tol in [((100, 50, 0, 1), 1e-6), ((100, 50, 10, 4), 1e-4)]:
        lines, aligned, gravity, gt = S.make_init_scene(n, n_al, n_out, seed=seed)
        blob = np.concatenate([lines.ravel(), aligned.astype(np.float64).ravel(), gravity.ravel()])
        r = subprocess.run([INIT_BIN, str(n)], input=blob.tobytes(), capture_output=True, check=True)
        rows = r.stdout.decode().strip().split("\n")
        ok, ratio, count = rows[0].split()
        assert ok == "1" and count == "4"
        poses = np.array([[float.fromhex(x) for x in row.split()] for row in rows[1:5]]).reshape(4, 3, 4)
        est = poses.copy()
        est[:, :, 3] /= np.linalg.norm(est[1, :, 3])
        for i in range(4):
            assert np.linalg.norm(est[i] - gt[i]) < tol
        ok2, poses2, ratio2, _ = I.initialize_reconstruction(lines, aligned, gravity)
        assert ok2 and float.fromhex(ratio) == ratio2 and np.array_equal(poses, poses2)


def test_dropin_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available() or not os.path.exists(BIN):
        pytest.skip("needs the prebuilt binary and a machine without a GPU")
    sc = S.make_abs_pose_scene(n=100, inlier_ratio=0.5, seed=1)
    fin = str(tmp_path / "in.bin")
    with open(fin, "wb") as f:
        np.array([100], dtype=np.int64).tofile(f)
        sc["lines"].tofile(f)
        sc["points"].tofile(f)
        sc["aligned"].astype(np.float64).tofile(f)
    r = subprocess.run([BIN, fin, str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "no usable CUDA device" in r.stderr      # no CPU fallback


@pytest.mark.gpu
def test_dropin_caller_matches_the_reference_loop(oracle, tmp_path):
    """The reference-typed caller through the drop-in on the GPU: pose, inlier count and mask are
    those of the reference's own RANSAC loop (oracle/_ref/libref_p6l.so where it travelled, and
    the oracle's wrapper of pose.cc:52-94); the refined pose is the oracle's refinement."""
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/dropin_pose_test not prebuilt (needs /root/reference to build)")
    n = 3000
    sc = S.make_abs_pose_scene(n=n, inlier_ratio=0.4, seed=33)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([n], dtype=np.int64).tofile(f)
        sc["lines"].tofile(f)
        sc["points"].tofile(f)
        sc["aligned"].astype(np.float64).tofile(f)
    subprocess.check_call([BIN, fin, fout])
    out = np.fromfile(fout, dtype=np.float64)
    ok, ninl, ok_ref = out[:3]
    q, t, q_ref, t_ref, mask = out[4:8], out[8:11], out[11:15], out[15:18], out[18:]
    o = oracle.make_options(0.012, 0.25, 0.99999, 3.0, 100, 10000)
    oracle.set_prng_seed(0)
    ok2, q2, t2, ninl2, mask2, rep2 = oracle.estimate_absolute_pose_from_lines(
        sc["lines"], sc["aligned"], sc["points"], o)
    assert bool(ok) == ok2 and ok2 and int(ninl) == ninl2
    assert np.array_equal(q, q2) and np.array_equal(t, t2)          # bit-exact
    assert np.array_equal(mask.astype(np.uint8), mask2)
    import oracle.reference as R
    if os.path.exists(R.LIB_PATH):
        R.set_prng_seed(0)
        rr, rmask = R.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], o)
        assert int(ninl) == rr.num_inliers and np.array_equal(mask.astype(np.uint8), rmask)
        assert np.array_equal(t, np.array(rr.model[9:12]))
    okr, qr, tr, _ = oracle.refine_absolute_pose(sc["lines"], sc["points"], mask2, 1,
                                                 [1000.0, 1000.0, 500.0, 500.0], q2, t2)
    assert bool(ok_ref) == okr
    assert np.abs(q_ref - qr).max() < 1e-9 and np.abs(t_ref - tr).max() < 1e-9
