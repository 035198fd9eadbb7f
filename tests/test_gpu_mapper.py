"""Minimal incremental-mapper loop: four-view initialisation -> register images
with the P6L RANSAC + refinement -> batched triangulation -> global BA -> filters, on a synthetic
scene; the reconstruction must match the ground truth up to a similarity."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import mapper as M

pytestmark = pytest.mark.gpu


def test_mapper_loop_reconstructs_the_scene(ctx):
    scene, gt = M.make_mapper_scene(num_images=12, num_points=600, seed=1)
    m = M.IncrementalMapper(ctx, scene, ba_every=3)
    assert m.run([0, 1, 2, 3])
    assert len(m.registered) == 12
    assert m.has_point.sum() > 0.8 * (~scene.aligned).sum()
    rot_err, centre_err = M.pose_errors(m, gt)
    assert rot_err < 2e-3 and centre_err < 2e-3, (rot_err, centre_err, m.log)
    kinds = [e[0] for e in m.log]
    assert kinds[0] == "init" and "register" in kinds and "triangulate" in kinds and "global_ba" in kinds


def test_mapper_loop_with_partial_visibility(ctx):
    scene, gt = M.make_mapper_scene(num_images=16, num_points=900, seed=2, visibility=0.6,
                                    noise_px=0.5)
    m = M.IncrementalMapper(ctx, scene, ba_every=4)
    assert m.run([0, 1, 2, 3])
    assert len(m.registered) >= 15
    rot_err, centre_err = M.pose_errors(m, gt)
    assert rot_err < 5e-3 and centre_err < 5e-3, (rot_err, centre_err)
