"""Line lifting (producer side): privacy_preserving_sfm_b200/lifting.py (CPU).

``ImageToWorld`` of all 11 camera models against the REFERENCE'S OWN camera models
(src/base/camera_models.{h,cc} compiled from /root/reference into oracle/_ref/libref_cost.so):
bit-identical for the models whose undistortion is arithmetic only (every point runs the
reference's Newton iteration and stops on its own step), 1e-14 for the ones that call
atan / tan / sin / cos; WorldToImage(ImageToWorld(x)) == x; the lifted lines pass through their
points, are unit-normalised in (a, b), gravity-aligned where asked — and are what the path
consumes: the generating pose has zero line residual on them; the float blob of the database
(src/base/database.cc:55-74) round-trips as the reference's FeatureLinesFromBlob defines.

Skipped where neither oracle/_ref/libref_cost.so nor /root/reference exists."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import lifting as L

MODELS = [(0, [900.0, 500, 480]), (1, [1000.0, 990, 500, 480]), (2, [900.0, 500, 480, 0.05]),
          (3, [900.0, 500, 480, 0.05, -0.01]),
          (4, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003]),
          (5, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003]),
          (6, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.004, 0.01, -0.005, 0.002]),
          (7, [1000.0, 990, 500, 480, 0.3]), (7, [1000.0, 990, 500, 480, 0.005]),
          (8, [900.0, 500, 480, 0.03]), (9, [900.0, 500, 480, 0.03, -0.004]),
          (10, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.004, -0.002, 0.001, -0.001])]
EXACT = {0, 1, 2, 3, 4, 6}


@pytest.fixture(scope="module")
def ref():
    import oracle.reference as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_cost.so not built and /root/reference absent")
    return R


def _pixels(n, seed):
    rng = np.random.default_rng(seed)
    xy = rng.uniform([0, 0], [1000, 960], size=(n, 2))
    xy[0] = [500.0, 480.0]                                  # the principal point
    xy[1] = [500.0 + 1e-9, 480.0]
    xy[2] = [500.0 + 3.0, 480.0 - 2.0]                      # radius^2 < 1e-4 (FOV series branch)
    return xy


@pytest.mark.parametrize("model,params", MODELS)
def test_image_to_world_matches_the_reference(ref, model, params):
    xy = _pixels(400, seed=model)
    uv, uv_ref = L.ImageToWorld(model, params, xy), ref.image_to_world(model, params, xy)
    if model in EXACT:
        assert np.array_equal(uv.view(np.uint64), uv_ref.view(np.uint64))
    else:
        assert np.abs(uv - uv_ref).max() <= 1e-14 * max(1.0, np.abs(uv_ref).max())
    back = np.array([ref.world_to_image(model, params, u, v) for u, v in uv])
    fwd = L.WorldToImage(model, params, uv)                 # the forward model, same pins
    if model in EXACT:
        assert np.array_equal(fwd.view(np.uint64), back.view(np.uint64))
    else:
        assert np.abs(fwd - back).max() <= 1e-12 * 1000.0
    # the Newton stop is ||step||^2 < 1e-10; the reference's FOV model at omega^2 < 1e-4 is a pair
    # of truncated series (camera_models.h:1147-1161, :1187-1193), not exact inverses of each other
    series = model == 7 and params[4] ** 2 < 1e-4
    assert np.abs(back - xy).max() < (5e-3 if series else 1e-6)


def test_lifted_lines_are_what_the_path_consumes(oracle, ref):
    from privacy_preserving_sfm_b200 import synthetic as S
    model, params = 4, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003]
    rng = np.random.default_rng(3)
    sc = S.make_abs_pose_scene(n=300, inlier_ratio=1.0, seed=9)          # for a pose and points
    R, t, X = np.asarray(sc["R"]), np.asarray(sc["t"]).reshape(3), sc["points"]
    pc = X @ R.T + t
    front = pc[:, 2] > 0.1
    uv = pc[front, :2] / pc[front, 2:3]
    keep = np.abs(uv).max(axis=1) < 0.45                                   # inside the image
    uv, Xf = uv[keep], X[front][keep]
    assert len(uv) > 50
    keypoints = np.array([ref.world_to_image(model, params, u, v) for u, v in uv])
    gravity = np.array([0.05, 0.99, 0.1])
    gravity /= np.linalg.norm(gravity)
    aligned = L.select_aligned_features(len(uv), 0.4, rng)
    assert 0.4 <= aligned.mean() < 0.4 + 1.5 / len(uv)
    lines, flags = L.LiftLines(model, params, keypoints, rng.uniform(-1, 1, size=(len(uv), 3)),
                               aligned, gravity)
    assert np.array_equal(flags, aligned.astype(np.uint8))
    assert np.abs(np.hypot(lines[:, 0], lines[:, 1]) - 1).max() < 1e-15
    assert np.abs(lines[:, 0] * uv[:, 0] + lines[:, 1] * uv[:, 1] + lines[:, 2]).max() < 1e-7
    assert np.abs(lines[aligned] @ gravity).max() < 1e-12                  # gravity lies in the plane
    res = oracle.line_residuals(lines, Xf, S.model_from_pose(R, t))
    assert res.max() < 1e-13                                               # squared distances
    # without gravity nothing is aligned
    _, none = L.LiftLines(model, params, keypoints, rng.uniform(-1, 1, size=(len(uv), 3)), aligned,
                          np.full(3, np.nan))
    assert not none.any()
    # database blob: floats, renormalised on the way back (database.cc:64-74)
    blob = L.feature_lines_to_blob(lines, flags)
    assert len(blob) == 16 * len(lines)
    lines2, flags2 = L.feature_lines_from_blob(blob, rows=len(lines))
    assert np.array_equal(flags2, flags) and 0 < np.abs(lines2 - lines).max() < 1e-6
    f = lines.astype(np.float32).astype(np.float64)
    assert np.array_equal(lines2, f / np.sqrt(f[:, 0] ** 2 + f[:, 1] ** 2)[:, None])
    with pytest.raises(ValueError):
        L.feature_lines_from_blob(blob, rows=len(lines) + 1)


def test_has_bogus_params_matches_the_reference(ref):
    rng = np.random.default_rng(12)
    seen = set()
    for model, params in MODELS:
        n = len(params)
        for _ in range(60):
            p = np.array(params, np.float64)
            k = int(rng.integers(0, n))
            p[k] = p[k] * rng.choice([-1.0, 0.05, 0.5, 1.0, 3.0, 40.0]) + rng.choice([0.0, 0.0, 1e3, -2e3, 1.5])
            args = (1000, 960, 0.1, 10.0, 1.0)
            got = L.HasBogusParams(model, p, *args)
            assert got == ref.has_bogus_params(model, p, *args), (model, p.tolist())
            seen.add(got)
        # the parameter groups themselves (ParameterizeCameras uses the same index lists)
        focal, pp, extra = L._PARAM_GROUPS[model]
        assert sorted(focal + pp + extra) == list(range(n))
    assert seen == {True, False}
