"""Host-side logic of the mapper driver that needs no GPU: the local bundle
of IncrementalMapper::FindLocalBundle (src/sfm/incremental_mapper.cc:993-1160) against a literal
restatement of the reference's loops on a reconstruction set from the generating scene."""
import math

import numpy as np
import pytest

from privacy_preserving_sfm_b200 import mapper as M
from privacy_preserving_sfm_b200 import model_io as IO


def _state(num_images, num_points, registered, seed, visibility=0.5, rings=2):
    scene, gt = M.make_mapper_scene(num_images=num_images, num_points=num_points, seed=seed,
                                    visibility=visibility, rings=rings)
    rng = np.random.default_rng(seed)
    m = M.IncrementalMapper(None, scene)
    m.registered = list(registered)
    for i in range(num_images):
        m.qvec[i], m.tvec[i] = M._rotmat_to_quat(gt["R"][i]), gt["t"][i]
    m.obs_on = scene.visible & (rng.uniform(size=scene.visible.shape) < 0.85)
    views = m.obs_on[m.registered].sum(axis=0)
    m.has_point = (views >= 3) & (rng.uniform(size=num_points) < 0.9)
    m.points = np.where(m.has_point[:, None], gt["points"], np.nan)
    return m


def _find_local_bundle_literal(m, i, num_images, min_tri_angle_deg):
    """incremental_mapper.cc:993-1160 with the reconstruction read off the mapper's arrays; the
    unordered map is walked in registration order and the sort is stable (ties are unspecified
    in the reference)."""
    reg = list(m.registered)
    point_ids = [p for p in np.flatnonzero(m.obs_on[i] & m.has_point)]
    shared = {}
    for p in point_ids:
        for j in reg:                                        # track elements of point p
            if j != i and m.obs_on[j, p]:
                shared[j] = shared.get(j, 0) + 1
    overlapping = sorted(((j, shared[j]) for j in reg if j in shared), key=lambda t: -t[1])
    num_eff = min(num_images - 1, len(overlapping))
    if len(overlapping) == num_eff:
        return [j for j, _ in overlapping]
    a = math.radians(min_tri_angle_deg)
    n = float(len(point_ids))
    thresholds = [(a / 1.0, 0.6 * n), (a / 1.5, 0.6 * n), (a / 2.0, 0.5 * n), (a / 2.5, 0.4 * n),
                  (a / 3.0, 0.3 * n), (a / 4.0, 0.2 * n), (a / 5.0, 0.1 * n), (a / 6.0, 0.1 * n)]
    c1 = IO.projection_centers(m.qvec[[i]], m.tvec[[i]])[0]
    tri = [-1.0] * len(overlapping)
    used = [False] * len(overlapping)
    local = []
    for angle_thr, overlap_thr in thresholds:
        for k, (j, count) in enumerate(overlapping):
            if count < overlap_thr:
                break
            if used[k]:
                continue
            if tri[k] < 0.0:
                c2 = IO.projection_centers(m.qvec[[j]], m.tvec[[j]])[0]
                base2 = float(((c1 - c2) ** 2).sum())
                angles = []
                for p in point_ids:
                    r1 = float(((m.points[p] - c1) ** 2).sum())
                    r2 = float(((m.points[p] - c2) ** 2).sum())
                    den = 2.0 * math.sqrt(r1 * r2)
                    if den == 0.0:
                        angles.append(0.0)
                        continue
                    ang = abs(math.acos((r1 + r2 - base2) / den))
                    angles.append(min(ang, math.pi - ang))
                idx = max(0, min(len(angles) - 1, int(round(75 / 100 * (len(angles) - 1)))))
                tri[k] = sorted(angles)[idx]
            if tri[k] >= angle_thr:
                local.append(j)
                used[k] = True
                if len(local) >= num_eff:
                    break
        if len(local) >= num_eff:
            break
    for k, (j, _) in enumerate(overlapping):
        if len(local) >= num_eff:
            break
        if not used[k]:
            local.append(j)
            used[k] = True
    return local


@pytest.mark.parametrize("seed,min_angle", [(1, 6.0), (2, 6.0), (3, 25.0), (4, 60.0), (5, 2.0)])
def test_find_local_bundle_follows_the_reference(seed, min_angle):
    m = _state(14, 500, [0, 1, 2, 3, 5, 6, 8, 9, 11, 12, 13], seed)
    m.ba_local_min_tri_angle = min_angle
    differs = 0
    for i in m.registered:
        got = m.find_local_bundle(i)
        want = _find_local_bundle_literal(m, i, m.ba_local_num_images, min_angle)
        assert got == want, (seed, i)
        assert len(got) == 5 and i not in got and set(got) <= set(m.registered)
        # the plain "most overlapping" choice, for comparison
        shared = np.array([(m.obs_on[j] & m.obs_on[i] & m.has_point).sum() if j != i else -1
                           for j in m.registered])
        plain = [m.registered[k] for k in np.argsort(-shared, kind="stable")[:5]]
        differs += got != plain
    if min_angle >= 25.0:
        assert differs > 0                                   # the angle criterion does change the choice


def test_find_local_bundle_small_and_empty_cases():
    m = _state(8, 200, [0, 1, 2, 3], seed=9, visibility=0.8)
    assert sorted(m.find_local_bundle(0)) == [1, 2, 3]       # fewer overlapping images than asked: all
    m.has_point[:] = False
    assert m.find_local_bundle(0) == []


def test_scene_mean_focal_is_the_references_threshold_scale():
    """Scene.mean_focal turns the mapper's pixel thresholds into normalised ones (max_error =
    12 px / mean focal, sfm/incremental_mapper.cc:673-674): Camera::ImageToWorldThreshold of the
    reference's own camera models (base/camera_models.h:533-543) for all 11 models."""
    import oracle.reference as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_cost.so not built and /root/reference absent")
    from test_ref_lifting import MODELS
    lines = np.full((2, 1, 3), np.nan)
    for model, params in MODELS:
        sc = M.Scene(lines, [False], np.zeros((2, 3)), model, params, (1000, 960))
        assert 12.0 / sc.mean_focal == R.image_to_world_threshold(model, params, 12.0), model
