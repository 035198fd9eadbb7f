"""Batched robust line triangulation: CUDA (one thread per track, streaming-QR
null vector) vs the CPU restatement of EstimateTriangulation (full n x 4 one-sided Jacobi SVD).
The two use different but equivalent factorisations and different libm arccosines, so points
agree to ~1e-9 relative and decisions are identical away from threshold ties."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import filters as F
from privacy_preserving_sfm_b200 import synthetic as S
from privacy_preserving_sfm_b200 import triangulation as T

pytestmark = pytest.mark.gpu


def _tracks(num_cams, num_points, obs, seed, outlier=0.15, noise_px=0.5, varlen=False):
    sc = S.make_ba_scene(num_cams=num_cams, num_points=num_points, obs_per_point=obs, seed=seed)
    rng = np.random.default_rng(seed + 7)
    order = np.argsort(sc["obs_pt"], kind="stable")
    obs_img, obs_pt, line = sc["obs_cam"][order], sc["obs_pt"][order], sc["obs_line"][order].copy()
    if varlen:                                   # ragged tracks: drop a random tail of every track
        keep = np.ones(len(obs_img), bool)
        start = np.searchsorted(obs_pt, np.arange(num_points + 1))
        for p in range(num_points):
            drop = rng.integers(0, obs - 1)
            keep[start[p + 1] - drop:start[p + 1]] = False
        obs_img, obs_pt, line = obs_img[keep], obs_pt[keep], line[keep]
    bad = rng.uniform(size=len(obs_img)) < outlier
    line[bad, 2] += rng.choice([-1.0, 1.0], bad.sum()) * rng.uniform(0.1, 0.3, bad.sum())
    track_start = np.searchsorted(obs_pt, np.arange(num_points + 1)).astype(np.int64)
    pb = F.FilterProblem(sc["qvecs"], sc["tvecs"], np.zeros(num_cams, np.int32), [1],
                         [[1000.0, 1000.0, 500.0, 500.0]], [(1000, 1000)],
                         np.zeros((num_points, 3)), track_start, obs_img, line,
                         np.zeros(len(obs_img), np.uint8))
    return pb, sc["points_gt"] if "points_gt" in sc else None, bad


def _mapper_options(**kw):
    # src/sfm/incremental_triangulator.cc:518-533 (min_angle 1.5 deg, create_max_angle_error 2 deg)
    o = dict(min_tri_angle=np.deg2rad(1.5), residual_type=T.ANGULAR_ERROR,
             max_error=np.deg2rad(2.0), confidence=0.9999, min_inlier_ratio=0.02,
             max_num_trials=10000, exhaustive_threshold=15)
    o.update(kw)
    return T.EstimateTriangulationOptions(**o)


def _compare(ctx, oracle, pb, opt, min_success=0.5):
    ok, xyz, mask, nt = T.EstimateTriangulationBatch(ctx, pb, opt)
    ok2, xyz2, mask2, nt2 = oracle.estimate_triangulation_batch(pb, opt)
    assert np.array_equal(ok, ok2) and ok.mean() >= min_success
    assert np.array_equal(nt, nt2)
    assert np.array_equal(mask, mask2)
    scale = np.maximum(1.0, np.abs(xyz2[ok]).max(axis=1, keepdims=True))
    assert (np.abs(xyz[ok] - xyz2[ok]) / scale).max() < 1e-8
    return ok, xyz, mask


@pytest.mark.parametrize("residual_type", [T.ANGULAR_ERROR, T.REPROJECTION_ERROR])
def test_triangulation_matches_oracle(ctx, oracle, residual_type):
    pb, _, bad = _tracks(12, 400, 8, seed=5)
    max_error = np.deg2rad(2.0) if residual_type == T.ANGULAR_ERROR else 4.0
    opt = _mapper_options(residual_type=residual_type, max_error=max_error)
    ok, xyz, mask = _compare(ctx, oracle, pb, opt, 0.9)
    # corrupted observations are rejected, clean ones kept (for the successful tracks)
    pt = np.repeat(np.arange(400), np.diff(pb.track_start))
    good = ok[pt]
    assert mask[good & ~bad].mean() > 0.97 and mask[good & bad].mean() < 0.15


def test_triangulation_recovers_points(ctx):
    sc = S.make_ba_scene(num_cams=10, num_points=300, obs_per_point=6, seed=9)
    order = np.argsort(sc["obs_pt"], kind="stable")
    ts = np.searchsorted(sc["obs_pt"][order], np.arange(301)).astype(np.int64)
    # noise-free lines through the true projections
    R = np.stack([S.quat_to_rotmat(q) for q in sc["qvecs_gt"]]) if "qvecs_gt" in sc else None
    pb = F.FilterProblem(sc["qvecs"], sc["tvecs"], np.zeros(10, np.int32), [1],
                         [[1000.0, 1000.0, 500.0, 500.0]], [(1000, 1000)], np.zeros((300, 3)), ts,
                         sc["obs_cam"][order], sc["obs_line"][order], np.zeros(len(order), np.uint8))
    ok, xyz, mask, nt = T.EstimateTriangulationBatch(ctx, pb, _mapper_options())
    assert ok.mean() > 0.9
    # the scene's poses / points are perturbed ground truth: triangulated points stay near them
    assert np.median(np.linalg.norm(xyz[ok] - sc["points"][ok], axis=1)) < 0.1


def test_triangulation_ragged_short_and_adaptive(ctx, oracle):
    pb, _, _ = _tracks(10, 300, 7, seed=11, varlen=True)     # lengths 2..7, some < 3
    opt = _mapper_options(exhaustive_threshold=0, min_num_trials=3)   # adaptive early stop
    ok, xyz, mask, nt = T.EstimateTriangulationBatch(ctx, pb, opt)
    ok2, xyz2, mask2, nt2 = oracle.estimate_triangulation_batch(pb, opt)
    lens = np.diff(pb.track_start)
    assert not ok[lens < 3].any() and (lens < 3).any()
    assert np.array_equal(ok, ok2) and np.array_equal(nt, nt2) and np.array_equal(mask, mask2)
    assert (nt[lens >= 3] < np.array([n * (n - 1) * (n - 2) // 6 for n in lens[lens >= 3]])).any()


def test_triangulation_full_size(ctx, oracle):
    """200k tracks x 10 views (the point set of BASELINE.json configs[3])."""
    pb, _, _ = _tracks(500, 200000, 10, seed=S.SCENE_SEED, outlier=0.05)
    opt = _mapper_options()
    ok, xyz, mask, nt = T.EstimateTriangulationBatch(ctx, pb, opt)
    # oracle on a sample of the tracks (it is a scalar loop)
    sel = np.arange(0, 200000, 97)
    keep = np.concatenate([np.arange(pb.track_start[t], pb.track_start[t + 1]) for t in sel])
    sub = F.FilterProblem(pb.qvecs, pb.tvecs, pb.image_camera, [1], [[1000.0, 1000.0, 500.0, 500.0]],
                          [(1000, 1000)], np.zeros((len(sel), 3)),
                          np.arange(len(sel) + 1, dtype=np.int64) * 10, pb.obs_image[keep],
                          pb.obs_line[keep], np.zeros(len(keep), np.uint8))
    ok2, xyz2, mask2, nt2 = oracle.estimate_triangulation_batch(sub, opt)
    assert np.array_equal(ok[sel], ok2) and np.array_equal(mask[keep], mask2)
    assert np.array_equal(nt[sel], nt2) and ok.mean() > 0.9
    scale = np.maximum(1.0, np.abs(xyz2[ok2]).max(axis=1, keepdims=True))
    assert (np.abs(xyz[sel][ok2] - xyz2[ok2]) / scale).max() < 1e-8
