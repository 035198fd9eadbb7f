#!/usr/bin/env python
"""BASELINE.json configs[4]: the incremental-mapper loop (register -> triangulate -> local /
global bundle adjustment -> filters) on a synthetic 1000-image scene, every operator on the GPU.

  python scripts/config5_mapper.py [images] [points] [visibility] [graph]
  torchrun --nproc-per-node N scripts/config5_mapper.py ...   (global BA sharded over N GPUs)

With a fourth argument `graph` the scene goes the way a matched image set does (SURVEY.md 8d,
config 5: "synthetic CorrespondenceGraph from ground-truth visibility"): pairwise matches of the
generated visibility -> correspondence_graph.CorrespondenceGraph -> tracks (connected components)
-> mapper.Scene, and the four initial images are searched and selected as
RegisterInitialLineImages does (mapper.select_initial_images, ten check images drawn with seed 0)
instead of being given.  (Building the graph of 1 000 images / 13 M correspondences takes ~30 s of host time and is
reported separately.  Added in the round's last session with under a GPU-minute left: run on a
B200 at 16 images / 1 000 points only — profiles/r02_s7_config5_graph_small.json: the search
finds the four initial images, all 16 images registered, 1.1e-3 rad / 7e-4 of the extent — the
host side is covered by tests/test_ref_correspondence_graph.py.)

Prints one JSON line: wall time of the loop, its split, registered images, final pose error
against the generating scene (after a similarity alignment)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import privacy_preserving_sfm_b200 as pp                      # noqa: E402
from privacy_preserving_sfm_b200 import mapper as M           # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
vis = float(sys.argv[3]) if len(sys.argv) > 3 else 0.03
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
ctx = pp.Context(local)
ba_ctx = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    ba_ctx = pp.Context(local)          # a second context with the communicator: every rank runs
    ba_ctx.comm_init_from_torch(dist)   # the same deterministic loop, global BA is collective
t0 = time.perf_counter()
scene, gt = M.make_mapper_scene(num_images=n_img, num_points=n_pts, seed=20201017,
                                visibility=vis, noise_px=0.5, rings=3)
t_scene = time.perf_counter() - t0
initial, graph_info = [0, 1, 2, 3], None
if len(sys.argv) > 4 and sys.argv[4] == "graph":
    from privacy_preserving_sfm_b200 import correspondence_graph as G
    t0 = time.perf_counter()
    graph, line_of, num_lines = G.graph_from_visibility(scene.visible)
    image_lines = [scene.lines[i, scene.visible[i]] for i in range(n_img)]
    image_aligned = [scene.aligned[scene.visible[i]] for i in range(n_img)]
    scene, tracks = M.Scene.from_correspondence_graph(graph, image_lines, image_aligned, scene.gravity,
                                                      scene.camera_model, scene.camera_params,
                                                      scene.camera_size)
    t_graph = time.perf_counter() - t0
    t0 = time.perf_counter()
    check = (np.random.default_rng(0).choice(n_img, min(10, n_img), replace=False) + 1).tolist()
    ok_init, image_set, _, ratio, tried = M.select_initial_images(
        graph, image_lines, image_aligned, scene.gravity, check, ctx=ctx)
    assert ok_init, "no initial image set found"
    initial = [i - 1 for i in image_set]
    graph_info = {"graph_and_tracks_s": t_graph, "initial_set_search_s": time.perf_counter() - t0,
                  "image_pairs": graph.NumImagePairs(), "tracks": len(tracks),
                  "initial_images": initial, "initial_inlier_ratio": ratio, "sets_tried": len(tried)}
m = M.IncrementalMapper(ctx, scene, local_ba=True, ba_ctx=ba_ctx)
t0 = time.perf_counter()
ok = m.run(initial)
wall = time.perf_counter() - t0
rot_err, centre_err = M.pose_errors(m, gt)
if rank == 0:
    kinds = [e[0] for e in m.log]
    print(json.dumps({
        "metric": "incremental mapper loop wall time (BASELINE.json configs[4])", "value": wall,
        "unit": "s", "higher_is_better": False, "n_gpus": world, "ok": bool(ok),
        "config": {"images": n_img, "points": n_pts, "observations": int(scene.visible.sum()),
                   "visibility": vis, "noise_px": 0.5},
        "registered_images": len(m.registered), "points3D": int(m.has_point.sum()),
        "images_per_s": len(m.registered) / wall,
        "split_s": {k: round(v, 3) for k, v in m.timing.items()},
        "calls": {k: kinds.count(k) for k in ("register", "triangulate", "local_ba", "global_ba")},
        "max_rotation_error_rad": rot_err, "max_centre_error_rel": centre_err,
        "scene_generation_s": t_scene, "from_correspondence_graph": graph_info,
        "note": "Python driver over the GPU operators; one shared PINHOLE camera, "
                + ("tracks = connected components of the correspondence graph, initial images searched"
                   if graph_info else "tracks given")}))
if world > 1:
    dist.destroy_process_group()
