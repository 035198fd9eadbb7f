"""Post-BA filters: the oracle against the REFERENCE'S OWN Reconstruction (CPU).

oracle/build_ref.sh compiles src/base/reconstruction.cc with the classes and functions it uses
(src/base/{image,point3d,track,camera,camera_models,pose,projection,triangulation}.cc) from where
they lie under /root/reference into oracle/_ref/libref_filter.so (stand-ins of oracle/ref/shim/
for Eigen / glog / Ceres / Boost / FreeImage; SQLite's header is the reference's vendored one).
oracle/ref/ref_filter.cc builds a colmap::Reconstruction from the flat track-major problem of the
C-ABI through the reference's own AddCamera / AddImage / RegisterImage / AddPoint3D, runs the
reference's Reconstruction::FilterPoints3D / FilterObservationsWithNegativeDepth and reads back
what they deleted.  Pinned against oracle/filter_oracle.cc (which the CUDA kernels equal bit for
bit, tests/test_gpu_filters.py): the number of filtered entries, which observations lost their
point (DeleteObservation's cascade included: a track of <= 3 elements goes as a whole), which points
are gone, and Point3D::Error() — bit-identical for the polynomial camera models.  One rounding
difference is inherent: the reference takes projection centres through Eigen's quaternion-vector
product, the oracle as -R^T t; triangulation angles agree to 1e-15 and no decision of these
scenes sits on the threshold.

Skipped where neither oracle/_ref/libref_filter.so nor /root/reference exists."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import filters as F
from privacy_preserving_sfm_b200 import synthetic as S

MODELS = [(0, [900.0, 500, 480]), (1, [1000.0, 990, 500, 480]), (2, [900.0, 500, 480, 0.05]),
          (3, [900.0, 500, 480, 0.05, -0.01]),
          (4, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003]),
          (5, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003]),
          (7, [1000.0, 990, 500, 480, 0.3]),
          (10, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.004, -0.002, 0.001, -0.001])]


@pytest.fixture(scope="module")
def ref():
    import oracle.reference as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_filter.so not built and /root/reference absent")
    return R


def _problem(num_cams, num_points, obs, seed, model=1, params=(1000.0, 1000.0, 500.0, 500.0),
             corrupt=0.15, varlen=False):
    # the generator of tests/test_gpu_filters.py (+ ragged tracks)
    sc = S.make_ba_scene(num_cams=num_cams, num_points=num_points, obs_per_point=obs, seed=int(seed))
    rng = np.random.default_rng(int(seed) + 1)
    order = np.argsort(sc["obs_pt"], kind="stable")
    obs_img, obs_pt, line = sc["obs_cam"][order], sc["obs_pt"][order], sc["obs_line"][order].copy()
    if varlen:
        keep = np.ones(len(obs_img), bool)
        start = np.searchsorted(obs_pt, np.arange(num_points + 1))
        for p in range(num_points):
            drop = rng.integers(0, obs)
            keep[start[p + 1] - drop:start[p + 1]] = False
        obs_img, obs_pt, line = obs_img[keep], obs_pt[keep], line[keep]
    track_start = np.searchsorted(obs_pt, np.arange(num_points + 1)).astype(np.int64)
    aligned = (rng.uniform(size=len(obs_img)) < 0.4).astype(np.uint8)
    bad = rng.uniform(size=len(obs_img)) < corrupt
    line[bad, 2] += rng.normal(scale=0.02, size=bad.sum())
    pts = sc["points"].copy()
    pts[rng.choice(num_points, max(1, num_points // 25), replace=False)] *= 400.0
    pts[rng.choice(num_points, max(1, num_points // 30), replace=False)] += 30.0 * np.array([0.0, 0.0, 1.0])
    for p in rng.choice(num_points, max(1, num_points // 40), replace=False):
        aligned[track_start[p]:track_start[p + 1]] = 1
    return F.FilterProblem(sc["qvecs"], sc["tvecs"], np.zeros(num_cams, np.int32), [model],
                           [list(params)], [(1000, 1000)], pts, track_start, obs_img, line, aligned)


@pytest.mark.parametrize("model,params", MODELS)
def test_filter_points3d_identical(oracle, ref, model, params):
    pb = _problem(12, 500, 6, seed=3 + model, model=model, params=params)
    nf, od, pd, pe, _ = oracle.filter_points3d(pb, 4.0, 1.5)
    nf2, od2, pd2, pe2 = ref.filter_points3d(pb, 4.0, 1.5)
    assert nf == nf2 and nf > 0
    assert np.array_equal(od, od2) and np.array_equal(pd, pd2)
    assert 0 < pd.sum() < len(pd) and 0 < od.sum() < len(od)
    alive = pd2 == 0
    assert np.array_equal(pe[alive].view(np.uint64), pe2[alive].view(np.uint64))   # Point3D::Error()
    assert (pe2[alive] >= 0).all()


def test_filter_points3d_ragged_tracks_and_thresholds(oracle, ref):
    pb = _problem(10, 400, 7, seed=21, varlen=True)        # track lengths 1..7
    for max_err, min_angle in [(4.0, 1.5), (1.0, 0.5), (12.0, 6.0), (0.2, 0.0)]:
        nf, od, pd, pe, _ = oracle.filter_points3d(pb, max_err, min_angle)
        nf2, od2, pd2, pe2 = ref.filter_points3d(pb, max_err, min_angle)
        assert nf == nf2
        assert np.array_equal(od, od2) and np.array_equal(pd, pd2)
        alive = pd2 == 0
        assert np.array_equal(pe[alive].view(np.uint64), pe2[alive].view(np.uint64))


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_filter_negative_depth_identical(oracle, ref, seed):
    pb = _problem(12, 500, 6, seed=seed, varlen=(seed == 7))
    # put more points behind some cameras so that the <= 3 cascade of DeleteObservation triggers
    rng = np.random.default_rng(seed)
    pts = pb.points.copy()
    sel = rng.choice(len(pts), 60, replace=False)
    pts[sel] += rng.normal(scale=4.0, size=(60, 3))
    pb2 = F.FilterProblem(pb.qvecs, pb.tvecs, pb.image_camera, [1], [[1000.0, 1000.0, 500.0, 500.0]],
                          [(1000, 1000)], pts, pb.track_start, pb.obs_image, pb.obs_line, pb.obs_aligned)
    nf, od, pd = oracle.filter_negative_depth(pb2)
    nf2, od2, pd2 = ref.filter_negative_depth(pb2)
    assert nf == nf2 and nf > 0
    assert np.array_equal(od, od2) and np.array_equal(pd, pd2)
    assert pd2.sum() > 0 and (od2.sum() > nf2)             # whole tracks went with their point
