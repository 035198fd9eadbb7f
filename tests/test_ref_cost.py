"""The oracle's line cost against the REFERENCE'S OWN cost functors and camera models (CPU).

oracle/build_ref.sh compiles src/base/cost_functions.h (BundleAdjustmentLineCostFunction,
BundleAdjustmentConstantPoseLineCostFunction) and src/base/camera_models.{h,cc} from where they
lie under /root/reference into oracle/_ref/libref_cost.so; Ceres, Eigen, glog and Boost are absent
in this image and replaced by the stand-ins of oracle/ref/shim/ (ceres::Jet,
UnitQuaternionRotatePoint and the scalar functions restated from Ceres' published sources).  The
functors are evaluated the way ceres::AutoDiffCostFunction evaluates them — every parameter block
(2; 4, 3, 3, k) seeded as a dual number.  Pinned here bit for bit: the functor text (projection,
closest point on the line, the two WorldToImage calls, the residual) and WorldToImage of all 11
camera models against oracle/ba_oracle.cc + oracle/camera_models_ext.h, which the CUDA
linearisation is tested against (tests/test_gpu_ba.py, tests/test_gpu_ba_intrinsics.py).  Ceres'
solver itself (trust region, Schur, parameterisations) stays unpinned.

Skipped where neither oracle/_ref/libref_cost.so nor /root/reference exists."""
import ctypes
import os
import re

import numpy as np
import pytest

from privacy_preserving_sfm_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# COLMAP model ids (src/base/camera_models.h:117-130) with plausible parameters
MODELS = {
    0: [900.0, 500, 480],                                              # SIMPLE_PINHOLE
    1: [1000.0, 990, 500, 480],                                        # PINHOLE
    2: [900.0, 500, 480, 0.05],                                        # SIMPLE_RADIAL
    3: [900.0, 500, 480, 0.05, -0.01],                                 # RADIAL
    4: [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003],            # OPENCV
    5: [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003],            # OPENCV_FISHEYE
    6: [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.001, 0.02, -0.004, 0.0005],  # FULL_OPENCV
    7: [1000.0, 990, 500, 480, 0.7],                                   # FOV
    8: [900.0, 500, 480, 0.05],                                        # SIMPLE_RADIAL_FISHEYE
    9: [900.0, 500, 480, 0.05, -0.01],                                 # RADIAL_FISHEYE
    10: [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.001, 0.0007, -0.0002, 0.0003],  # THIN_PRISM_FISHEYE
}


@pytest.fixture(scope="module")
def ref():
    import oracle.reference as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_cost.so not built and /root/reference absent")
    return R


def _same_bits(a, b):
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def _observations(count, seed):
    sc = S.make_ba_scene(num_cams=6, num_points=60, obs_per_point=4, seed=seed, noise_px=1.0)
    rng = np.random.default_rng(seed)
    for o in rng.integers(0, len(sc["obs_cam"]), count):
        ci, pi = sc["obs_cam"][o], sc["obs_pt"][o]
        yield sc["obs_line"][o], sc["qvecs"][ci], sc["tvecs"][ci], sc["points"][pi]


@pytest.mark.parametrize("model", sorted(MODELS))
def test_line_cost_functor_bit_identical(oracle, ref, model):
    """cost_functions.h:62-100 with (2; 4, 3, 3, k) dual numbers: residual and the Jacobians with
    respect to qvec, tvec, point3D and camera_params."""
    p = MODELS[model]
    assert ref.camera_num_params(model) == len(p)
    for line, q, t, X in _observations(150, seed=31 + model):
        a = oracle.line_cost_intr(model, p, line, q, t, X)
        b = ref.line_cost_intr(model, p, line, q, t, X)
        for x, y in zip(a, b):
            assert _same_bits(x, y)
        assert np.abs(b[4][:, len(p):]).max(initial=0.0) == 0.0
        # the (2; 4, 3, 3) block the constant-intrinsics solver uses is the same evaluation
        r, jq, jt, jx = oracle.line_cost(model, p, line, q, t, X)
        assert _same_bits(r, b[0]) and _same_bits(jq, b[1]) and _same_bits(jt, b[2]) and _same_bits(jx, b[3])


@pytest.mark.parametrize("model", sorted(MODELS))
def test_constant_pose_functor_bit_identical(oracle, ref, model):
    """cost_functions.h:130-176: the pose baked into the functor, blocks (2; 3, k).  Same residual
    and point / camera Jacobians as the variable-pose functor (the product masks the pose block,
    SURVEY 8 row A12)."""
    p = MODELS[model]
    for line, q, t, X in _observations(100, seed=57 + model):
        a = oracle.line_cost_intr(model, p, line, q, t, X)
        r, jx, jc = ref.constant_pose_line_cost(model, p, line, q, t, X)
        assert _same_bits(r, a[0]) and _same_bits(jx, a[3]) and _same_bits(jc, a[4])


@pytest.mark.parametrize("model", sorted(MODELS))
def test_plain_double_residual(oracle, ref, model):
    """Without Jacobians Ceres evaluates the functor on doubles; the value part of a dual-number
    quotient is a * (1 / b), so the two evaluations may differ in the last bits (they do for the
    fisheye models) — never by more than a few ulp of the projection."""
    p = MODELS[model]
    for line, q, t, X in _observations(100, seed=83 + model):
        rd = ref.line_residual(model, p, line, q, t, X)
        rj = ref.line_cost_intr(model, p, line, q, t, X)[0]
        assert np.abs(rd - rj).max() <= 1e-12 * 1000.0


def test_fov_and_fisheye_branches(oracle, ref):
    # FOV: omega^2 < 1e-4 (series), radius^2 < 1e-4 (series), generic; fisheye: r -> 0
    q, t = np.array([1.0, 0, 0, 0]), np.zeros(3)
    for omega in (1e-3, 0.7):
        for X in ([1e-3, -2e-3, 1.0], [0.3, -0.2, 1.5], [0.0, 0.0, 2.0]):
            for line in ([0.6, 0.8, 0.1], [1.0, 0.0, -0.2]):
                p = [1000.0, 990, 500, 480, omega]
                a = oracle.line_cost_intr(7, p, line, q, t, X)
                b = ref.line_cost_intr(7, p, line, q, t, X)
                assert all(_same_bits(x, y) for x, y in zip(a, b))
    for model in (5, 8, 9, 10):
        for X in ([0.0, 0.0, 2.0], [1e-20, 0.0, 1.0]):
            line = [0.6, 0.8, 0.0]          # the line passes through the principal axis too
            a = oracle.line_cost_intr(model, MODELS[model], line, q, t, X)
            b = ref.line_cost_intr(model, MODELS[model], line, q, t, X)
            assert all(_same_bits(x, y) for x, y in zip(a, b))


def _mask_table(path, name):
    src = open(os.path.join(ROOT, path)).read()
    m = re.search(name + r"\[11\]\s*=\s*\{([^}]*)\}", src)
    assert m, (path, name)
    return [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]


def test_parameter_groups_match_the_reference(ref):
    """FocalLengthIdxs / PrincipalPointIdxs / ExtraParamsIdxs of the 11 models
    (camera_models.cc:42-56) against the bit masks BundleAdjuster::ParameterizeCameras is restated
    with, in the product (csrc/ba_host.cu) and in the oracle (ba_oracle.cc)."""
    want = [[sum(1 << i for i in ref.camera_param_idxs(m, g)) for m in range(11)] for g in range(3)]
    for path, names in [("privacy_preserving_sfm_b200/csrc/ba_host.cu",
                         ("kFocalMask", "kPrincipalMask", "kExtraMask")),
                        ("oracle/ba_oracle.cc", ("kFocal", "kPP", "kExtra"))]:
        for g, name in enumerate(names):
            assert _mask_table(path, name) == want[g], (path, name)
    for m in range(11):   # the three groups partition Camera::Params()
        assert want[0][m] | want[1][m] | want[2][m] == (1 << ref.camera_num_params(m)) - 1
        assert want[0][m] & want[1][m] == want[0][m] & want[2][m] == want[1][m] & want[2][m] == 0


def test_image_to_world_threshold_matches_the_reference(ref):
    """BaseCameraModel::ImageToWorldThreshold (camera_models.h:533-543) against the library's
    host-only ppsfm_image_to_world_threshold (no GPU needed)."""
    import privacy_preserving_sfm_b200 as pp
    L = pp.load_library()
    L.ppsfm_image_to_world_threshold.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                                 ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    L.ppsfm_image_to_world_threshold.restype = ctypes.c_int
    for m, p in MODELS.items():
        for thr in (12.0, 4.0, 0.37):
            arr = (ctypes.c_double * 12)(*p)
            out = ctypes.c_double()
            assert L.ppsfm_image_to_world_threshold(m, arr, thr, ctypes.byref(out)) == 0
            assert out.value == ref.image_to_world_threshold(m, p, thr)


def test_world_to_image_of_every_model(oracle, ref):
    """CameraModelWorldToImage on a grid of normalised coordinates, against the projections the
    oracle's filters use (oracle/filter_oracle.cc -> same WorldToImage restatement), through the
    functor identity: with q = identity, t = 0 and X = (u, v, 1) the residual is
    WorldToImage(u, v) - WorldToImage(closest point on the line)."""
    q, t = np.array([1.0, 0, 0, 0]), np.zeros(3)
    rng = np.random.default_rng(5)
    for m, p in MODELS.items():
        for _ in range(50):
            u, v = rng.uniform(-0.4, 0.4, 2)
            th = rng.uniform(0, 2 * np.pi)
            a, b, c = np.cos(th), np.sin(th), rng.uniform(-0.2, 0.2)
            alpha = a * u + b * v + c
            w0 = ref.world_to_image(m, p, u, v)
            w1 = ref.world_to_image(m, p, u - alpha * a, v - alpha * b)
            r = oracle.line_cost(m, p, [a, b, c], q, t, [u, v, 1.0])[0]
            # (u, v, 1) rotated by the identity quaternion and divided by z = 1 is exact
            assert np.abs(r - (w0 - w1)).max() <= 1e-9
