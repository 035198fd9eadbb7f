"""Four-view initialisation with the candidate models scored on the GPU (SURVEY.md §8 f4:
FourView2dEstimator src/init/sfm2d.cc:302-444, PlanarOffsetEstimator src/init/initializer.cc:219-333
under LO-MSAC).  The device evaluates every track of every candidate model with the host
estimators' arithmetic and sums in track order, so the run must equal the host-only run bit for bit:
poses, inlier ratio, inlier counts and iteration counts (a single different score would change the
sequence of accepted models)."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import initializer as I, synthetic as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,n_al,n_out,seed,tilt", [
    (100, 50, 0, 1, 0.0), (100, 50, 10, 4, 0.0), (300, 120, 45, 13, 0.0), (60, 40, 20, 14, 0.0),
    (120, 60, 0, 7, 12.0), (257, 129, 30, 21, 5.0)])
def test_gpu_scoring_reproduces_the_host_run(ctx, n, n_al, n_out, seed, tilt):
    lines, aligned, gravity, gt = S.make_init_scene(n, n_al, n_out, seed=seed, tilt_deg=tilt)
    ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, gravity)
    ok2, poses2, ratio2, rep2 = I.initialize_reconstruction(lines, aligned, gravity, ctx=ctx)
    assert ok == ok2 and ratio == ratio2
    assert np.array_equal(poses, poses2)
    assert (rep.inliers_2d, rep.inliers_3d, rep.iterations_2d, rep.iterations_3d) == \
        (rep2.inliers_2d, rep2.inliers_3d, rep2.iterations_2d, rep2.iterations_3d)
    assert rep.mean_tri_angle_deg == rep2.mean_tri_angle_deg
    # one launch per minimal sample with at least one model, plus the local optimisations'
    assert rep2.gpu_launches >= rep2.iterations_3d


def test_config1_on_the_gpu(ctx):
    """BASELINE.json configs[0]: 2 000 tracks (1 000 aligned), 10 % outliers, max_error 0.005."""
    import time
    lines, aligned, gravity, gt = S.make_init_scene(2000, 1000, 200, seed=S.SCENE_SEED)
    opt = I.InitOptions(max_error=0.005)
    t0 = time.perf_counter()
    ok, poses, ratio, rep = I.initialize_reconstruction(lines, aligned, gravity, opt)
    t1 = time.perf_counter()
    ok2, poses2, ratio2, rep2 = I.initialize_reconstruction(lines, aligned, gravity, opt, ctx=ctx)
    t2 = time.perf_counter()
    assert ok and ok2 and ratio == ratio2 and np.array_equal(poses, poses2)
    assert (rep.iterations_2d, rep.iterations_3d) == (rep2.iterations_2d, rep2.iterations_3d)
    p = poses2.copy()
    p[:, :, 3] /= np.linalg.norm(p[1, :, 3])
    assert np.abs(p - gt).max() < 1e-4
    print(f"config 1: host {1e3 * (t1 - t0):.0f} ms, GPU-scored {1e3 * (t2 - t1):.0f} ms, "
          f"{rep2.gpu_launches} launches")
    assert (t2 - t1) < (t1 - t0)      # the point of the exercise


def test_contract_violations_on_the_gpu_path(ctx):
    lines, aligned, gravity, _ = S.make_init_scene(40, 20, 0, seed=9)
    bad = aligned.copy()
    bad[2, :] = 1 - bad[2, :]
    with pytest.raises(ValueError):
        I.initialize_reconstruction(lines, bad, gravity, ctx=ctx)
    idx = np.flatnonzero(aligned[0])[:4]        # too few tracks for a minimal sample
    ok, _, _, _ = I.initialize_reconstruction(lines[:, idx], aligned[:, idx], gravity, ctx=ctx)
    assert not ok
