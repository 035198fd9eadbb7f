"""Minimal incremental-mapper loop: four-view initialisation -> register images
with the P6L RANSAC + refinement -> batched triangulation -> global BA -> filters, on a synthetic
scene; the reconstruction must match the ground truth up to a similarity."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import mapper as M
from privacy_preserving_sfm_b200 import model_io

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


def test_mapper_controller_schedule_with_local_ba(ctx, tmp_path):
    """The controller's schedule (controllers/incremental_mapper.cc:484-510): local bundle
    adjustment after every registration (new image + the images sharing most points, gauge by
    constant pose / constant tvec[0], short tracks variable, views outside the bundle through
    constant poses, SOFT_L1), track completion, global adjustment when the model has grown by
    10 %; then the text model round trip (cameras / images / points3D.txt of this fork)."""
    scene, gt = M.make_mapper_scene(num_images=30, num_points=1500, seed=5, visibility=0.5,
                                    noise_px=0.5, rings=3)
    m = M.IncrementalMapper(ctx, scene, local_ba=True)
    assert m.run([0, 1, 2, 3])
    assert len(m.registered) >= 29
    rot_err, centre_err = M.pose_errors(m, gt)
    assert rot_err < 5e-3 and centre_err < 5e-3, (rot_err, centre_err)
    kinds = [e[0] for e in m.log]
    assert kinds.count("local_ba") >= 20 and 3 <= kinds.count("global_ba") < kinds.count("local_ba")
    lb = [e for e in m.log if e[0] == "local_ba"]
    assert all(e[4] <= e[3] * (1 + 1e-9) for e in lb)         # every local solve lowers the cost
    assert max(e[2] for e in lb) == 6                          # ba_local_num_images
    # text model round trip
    out = str(tmp_path / "model")
    m.write_text(out)
    model = M.IncrementalMapper.read_text(out)
    assert len(model["cameras"]) == 1 and model["cameras"][1][0] == "PINHOLE"
    assert sorted(model["images"]) == sorted(i + 1 for i in m.registered)
    assert len(model["points"]) == int(m.has_point.sum())
    for img_id, (q, t, cam_id, name, lines) in model["images"].items():
        i = img_id - 1
        assert np.array_equal(q, model_io.normalize_quaternion(m.qvec[i])) and np.array_equal(t, m.tvec[i])
        vis = np.flatnonzero(scene.visible[i])
        assert np.array_equal(lines[:, :3], scene.lines[i, vis])          # 17 digits: exact
        has = m.obs_on[i, vis] & m.has_point[vis]
        assert np.array_equal(lines[:, 4], np.where(has, vis + 1, -1))
    pid, (xyz, err, track) = next(iter(model["points"].items()))
    assert np.array_equal(xyz, m.points[pid - 1]) and len(track) >= 2
    for img_id, line_idx in track:                              # track elements point back
        assert model["images"][img_id][4][line_idx, 4] == pid
