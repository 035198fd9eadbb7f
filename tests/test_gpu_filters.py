"""Post-BA filters: CUDA kernels vs the CPU restatement of
Reconstruction::FilterPoints3D / FilterObservationsWithNegativeDepth — bit-exact masks."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import filters as F
from privacy_preserving_sfm_b200 import synthetic as S

pytestmark = pytest.mark.gpu

MODELS = [(0, [900.0, 500, 480]), (1, [1000.0, 990, 500, 480]), (2, [900.0, 500, 480, 0.05]),
          (3, [900.0, 500, 480, 0.05, -0.01]),
          (4, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003])]


def _problem(num_cams, num_points, obs, seed, model=1, params=(1000.0, 1000.0, 500.0, 500.0),
             corrupt=0.15):
    sc = S.make_ba_scene(num_cams=num_cams, num_points=num_points, obs_per_point=obs, seed=seed)
    rng = np.random.default_rng(seed + 1)
    order = np.argsort(sc["obs_pt"], kind="stable")
    obs_img, obs_pt, line = sc["obs_cam"][order], sc["obs_pt"][order], sc["obs_line"][order].copy()
    track_start = np.searchsorted(obs_pt, np.arange(num_points + 1)).astype(np.int64)
    aligned = (rng.uniform(size=len(obs_img)) < 0.4).astype(np.uint8)
    # corrupt some lines (large reprojection error), move a few points behind / far away
    bad = rng.uniform(size=len(obs_img)) < corrupt
    line[bad, 2] += rng.normal(scale=0.02, size=bad.sum())
    pts = sc["points"].copy()
    far = rng.choice(num_points, max(1, num_points // 25), replace=False)
    pts[far] *= 400.0                      # tiny triangulation angles
    behind = rng.choice(num_points, max(1, num_points // 30), replace=False)
    pts[behind] += 30.0 * np.array([0.0, 0.0, 1.0])
    allal = rng.choice(num_points, max(1, num_points // 40), replace=False)
    for p in allal:                        # tracks with only gravity-aligned lines are dropped
        aligned[track_start[p]:track_start[p + 1]] = 1
    return F.FilterProblem(sc["qvecs"], sc["tvecs"], np.zeros(num_cams, np.int32), [model],
                           [list(params)], [(1000, 1000)], pts, track_start, obs_img, line, aligned)


@pytest.mark.parametrize("model,params", MODELS)
def test_filter_points3d_matches_oracle(ctx, oracle, model, params):
    pb = _problem(12, 600, 6, seed=3 + model, model=model, params=params)
    nf, od, pd, pe = F.FilterPoints3D(ctx, pb, 4.0, 1.5)
    nf2, od2, pd2, pe2, _ = oracle.filter_points3d(pb, 4.0, 1.5)
    assert nf == nf2 and nf > 0
    assert np.array_equal(od, od2) and np.array_equal(pd, pd2)
    assert 0 < pd.sum() < len(pd) and 0 < od.sum() < len(od)
    assert np.array_equal(pe, pe2)         # same operations in the same order: bit-identical


def test_filter_points3d_short_and_empty_tracks(ctx, oracle):
    pb = _problem(5, 40, 2, seed=9)        # every track shorter than 3 -> everything deleted
    nf, od, pd, _ = F.FilterPoints3D(ctx, pb, 4.0, 1.5)
    assert pd.all() and od.all() and nf == len(od)
    nf2, od2, pd2, _, _ = oracle.filter_points3d(pb, 4.0, 1.5)
    assert nf == nf2 and np.array_equal(od, od2) and np.array_equal(pd, pd2)
    empty = F.FilterProblem(np.zeros((2, 4)) + [1, 0, 0, 0], np.zeros((2, 3)), [0, 0], [1],
                            [[1000.0, 1000.0, 500.0, 500.0]], [(1000, 1000)], np.zeros((3, 3)),
                            np.zeros(4, np.int64), np.zeros(0, np.int32), np.zeros((0, 3)),
                            np.zeros(0, np.uint8))
    nf, od, pd, _ = F.FilterPoints3D(ctx, empty, 4.0, 1.5)
    assert nf == 0 and len(od) == 0 and not pd.any()


def test_filter_negative_depth_matches_oracle(ctx, oracle):
    pb = _problem(10, 500, 5, seed=21)
    pb.points[::7] -= 20.0 * np.array([0.0, 0.0, 1.0])   # (in place: the struct points here)
    nf, od = F.FilterObservationsWithNegativeDepth(ctx, pb)
    nf2, od2 = oracle.filter_negative_depth(pb)
    assert nf == nf2 and np.array_equal(od, od2) and 0 < nf < len(od)


def test_filter_rejects_bad_input(ctx):
    pb = _problem(5, 40, 4, seed=10)
    pb.obs_line[3, 0] += 0.5
    with pytest.raises(Exception):
        F.FilterPoints3D(ctx, pb, 4.0, 1.5)


def test_filter_full_size(ctx, oracle):
    """BASELINE.json configs[3] size (500 cameras / 200k points / 2M observations)."""
    pb = _problem(500, 200000, 10, seed=S.SCENE_SEED, corrupt=0.05)
    nf, od, pd, pe = F.FilterPoints3D(ctx, pb, 4.0, 1.5)
    nf2, od2, pd2, pe2, _ = oracle.filter_points3d(pb, 4.0, 1.5)
    assert nf == nf2 and np.array_equal(od, od2) and np.array_equal(pd, pd2)
    assert np.array_equal(pe, pe2)
