"""Robust line triangulation: the oracle against the REFERENCE'S OWN sources
(CPU).

oracle/build_ref.sh compiles src/estimators/triangulation.cc (EstimateTriangulation,
TriangulationEstimator), src/base/triangulation.cc (TriangulateMultiViewPoint,
CalculateTriangulationAngle), src/base/projection.cc (CalculateNormalizedLineAngularError,
CalculateSquaredLineReprojectionError with the in-image test, HasPointPositiveDepth),
src/base/camera.cc + camera_models.cc, src/optim/loransac.h + ransac.h,
src/optim/combination_sampler.cc, src/util/math.{h,cc} and src/optim/support_measurement.cc from
where they lie under /root/reference into oracle/_ref/libref_tri.so (stand-ins of
oracle/ref/shim/ for Eigen / glog / Ceres / Boost).  The one piece that is not the reference's
text is the n x 4 JacobiSVD inside TriangulateMultiViewPoint: the stand-in returns the null vector
of eigen_restated::NullVectorNx4, the iteration the oracle uses (Eigen's own JacobiSVD gives the
same vector up to rounding).  Everything else — the residuals of both types, cheirality and
angle tests, sampling order, support comparison, local optimisation, the dynamic trial bound, the
final mask — is pinned bit for bit: success flags, points, inlier masks and trial counts of
oracle/triangulation_oracle.cc equal the reference's.  The CUDA kernel is compared with the oracle
in tests/test_gpu_triangulation.py.

Skipped where neither oracle/_ref/libref_tri.so nor /root/reference exists."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import filters as F
from privacy_preserving_sfm_b200 import synthetic as S
from privacy_preserving_sfm_b200 import triangulation as T


@pytest.fixture(scope="module")
def ref():
    import oracle.reference as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_tri.so not built and /root/reference absent")
    return R


def _tracks(num_cams, num_points, obs, seed, outlier=0.15, varlen=False):
    # the generator of tests/test_gpu_triangulation.py
    sc = S.make_ba_scene(num_cams=num_cams, num_points=num_points, obs_per_point=obs, seed=seed)
    rng = np.random.default_rng(seed + 7)
    order = np.argsort(sc["obs_pt"], kind="stable")
    obs_img, obs_pt, line = sc["obs_cam"][order], sc["obs_pt"][order], sc["obs_line"][order].copy()
    if varlen:
        keep = np.ones(len(obs_img), bool)
        start = np.searchsorted(obs_pt, np.arange(num_points + 1))
        for p in range(num_points):
            drop = rng.integers(0, obs - 1)
            keep[start[p + 1] - drop:start[p + 1]] = False
        obs_img, obs_pt, line = obs_img[keep], obs_pt[keep], line[keep]
    bad = rng.uniform(size=len(obs_img)) < outlier
    line[bad, 2] += rng.choice([-1.0, 1.0], bad.sum()) * rng.uniform(0.1, 0.3, bad.sum())
    track_start = np.searchsorted(obs_pt, np.arange(num_points + 1)).astype(np.int64)
    return F.FilterProblem(sc["qvecs"], sc["tvecs"], np.zeros(num_cams, np.int32), [1],
                           [[1000.0, 1000.0, 500.0, 500.0]], [(1000, 1000)],
                           np.zeros((num_points, 3)), track_start, obs_img, line,
                           np.zeros(len(obs_img), np.uint8))


def _options(**kw):
    # src/sfm/incremental_triangulator.cc:518-533
    o = dict(min_tri_angle=np.deg2rad(1.5), residual_type=T.ANGULAR_ERROR,
             max_error=np.deg2rad(2.0), confidence=0.9999, min_inlier_ratio=0.02,
             max_num_trials=10000, exhaustive_threshold=15)
    o.update(kw)
    return T.EstimateTriangulationOptions(**o)


def _identical(a, b):
    ok, xyz, mask, nt = a
    ok2, xyz2, mask2, nt2 = b
    assert np.array_equal(ok, ok2) and np.array_equal(nt, nt2) and np.array_equal(mask, mask2)
    assert np.array_equal(xyz[ok].view(np.uint64), xyz2[ok2].view(np.uint64))


@pytest.mark.parametrize("residual_type,max_error", [(T.ANGULAR_ERROR, np.deg2rad(2.0)),
                                                     (T.REPROJECTION_ERROR, 4.0)])
def test_exhaustive_sampling_identical(oracle, ref, residual_type, max_error):
    pb = _tracks(12, 400, 8, seed=5)
    opt = _options(residual_type=residual_type, max_error=max_error)
    a, b = oracle.estimate_triangulation_batch(pb, opt), ref.estimate_triangulation_batch(pb, opt)
    _identical(a, b)
    assert a[0].mean() > 0.9 and (a[3] == 56).all()       # C(8, 3) trials per track


def test_adaptive_abort_and_ragged_tracks_identical(oracle, ref):
    pb = _tracks(10, 300, 7, seed=11, varlen=True)          # lengths 2..7, some below 3
    opt = _options(exhaustive_threshold=0, min_num_trials=3)
    a, b = oracle.estimate_triangulation_batch(pb, opt), ref.estimate_triangulation_batch(pb, opt)
    _identical(a, b)
    lens = np.diff(pb.track_start)
    assert (lens < 3).any() and not a[0][lens < 3].any()
    full = np.array([n * (n - 1) * (n - 2) // 6 for n in lens])
    assert (a[3][lens >= 3] < full[lens >= 3]).any()      # the dynamic bound stopped some early


def test_long_tracks_with_trial_cap_identical(oracle, ref):
    # 24 views: C(24, 3) = 2024 combinations, capped by max_num_trials; many outliers so that
    # the local optimisation both wins and loses
    pb = _tracks(30, 20, 24, seed=13, outlier=0.4)
    for kw in (dict(max_num_trials=300, exhaustive_threshold=15),
               dict(max_num_trials=10000, exhaustive_threshold=30),
               dict(max_num_trials=50, min_num_trials=50, exhaustive_threshold=0, confidence=0.5)):
        opt = _options(**kw)
        _identical(oracle.estimate_triangulation_batch(pb, opt), ref.estimate_triangulation_batch(pb, opt))


def test_min_inlier_ratio_cap_identical(oracle, ref):
    # a large min_inlier_ratio makes the RANSAC constructor's cap (ransac.h:144-156) the binding
    # limit on the trial count
    pb = _tracks(12, 200, 10, seed=17, outlier=0.3)
    opt = _options(min_inlier_ratio=0.6, exhaustive_threshold=0, min_num_trials=0)
    a, b = oracle.estimate_triangulation_batch(pb, opt), ref.estimate_triangulation_batch(pb, opt)
    _identical(a, b)
    assert a[3].max() < 120
