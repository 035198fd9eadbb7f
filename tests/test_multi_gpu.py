"""Multi-GPU path.

CPU (gloo, world_size 2): the host-side sharding of a BA problem — every kept observation lands
on exactly one rank, camera blocks are identical on all ranks — and the rendezvous helper.
GPU (-m gpu, needs >= 2 devices): sharded BA over NCCL gives the single-GPU answer."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch
import torch.distributed as dist
from privacy_preserving_sfm_b200 import bundle_adjustment as ba, synthetic as S

backend = sys.argv[1]
dist.init_process_group(backend)
rank, world = dist.get_rank(), dist.get_world_size()
sc = S.make_ba_scene(num_cams=9, num_points=401, obs_per_point=5, seed=7)
flags = np.zeros(9, np.uint8); flags[0] = 1; flags[1] = 2; flags[8] = 1
pc = (np.arange(401) % 11 == 0).astype(np.uint8)
args = (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"], sc["obs_line"],
        [1], [sc["cam_params"]])
arr = ba.BaArrays(*args, pose_flags=flags, point_const=pc)
if backend == "gloo":
    local, pts, blocks, total = ba.shard_stats(arr, rank, world)
    t = torch.tensor([local, pts, blocks, total], dtype=torch.int64)
    s = t.clone(); dist.all_reduce(s)
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    one = ba.shard_stats(arr, 0, 1)
    assert s[0].item() == total == one[0], (s, total, one)        # observations partitioned
    assert s[1].item() == one[1]                                   # points partitioned
    assert mx[2].item() == mn[2].item() == one[2]                  # same camera blocks everywhere
    assert abs(local - total / world) < 0.1 * total                # balanced
    if rank == 0:
        print("GLOO_OK")
else:
    import privacy_preserving_sfm_b200 as pp
    torch.cuda.set_device(rank)
    ctx = pp.Context(rank)
    ctx.comm_init_from_torch(dist)
    assert (ctx.comm_rank(), ctx.comm_world_size()) == (rank, world)
    v = ctx.comm_allreduce_sum(np.array([1.0 + rank, 2.0]))
    assert np.allclose(v, [sum(1.0 + r for r in range(world)), 2.0 * world])
    kw = dict(max_num_iterations=15, gradient_tolerance=1e-3)
    ok, s = ba.solve_arrays(ctx, arr, ba.default_solver_options(**kw))
    np.save(os.path.join(sys.argv[2], f"q{{rank}}.npy"), arr.qvecs)
    np.save(os.path.join(sys.argv[2], f"t{{rank}}.npy"), arr.tvecs)
    np.save(os.path.join(sys.argv[2], f"X{{rank}}.npy"), arr.points)
    np.save(os.path.join(sys.argv[2], f"c{{rank}}.npy"),
            np.array([s.initial_cost, s.final_cost, s.num_successful_steps, s.num_unsuccessful_steps]))
    # intrinsics refinement, sharded: the intrinsics blocks are replicated like the pose blocks
    sn = S.make_ba_scene(num_cams=9, num_points=401, obs_per_point=5, seed=7, noise_px=2.0)
    prm = [[1000.0, 500, 500, 0.08], [1000.0, 500, 500, 0.05]]
    ai = ba.BaArrays(sn["qvecs"], sn["tvecs"], sn["points"], sn["obs_cam"], sn["obs_pt"],
                     sn["obs_line"], [2, 2], prm, image_camera=np.arange(9) % 2, pose_flags=flags,
                     point_const=pc)
    ok, si = ba.solve_arrays(ctx, ai, ba.default_solver_options(max_num_iterations=6,
                                                                refine_extra_params=1))
    np.save(os.path.join(sys.argv[2], f"iq{{rank}}.npy"), ai.qvecs)
    np.save(os.path.join(sys.argv[2], f"iX{{rank}}.npy"), ai.points)
    np.save(os.path.join(sys.argv[2], f"ip{{rank}}.npy"), ai.camera_params)
    np.save(os.path.join(sys.argv[2], f"ic{{rank}}.npy"),
            np.array([si.initial_cost, si.final_cost, si.num_successful_steps,
                      si.num_unsuccessful_steps]))
    dist.barrier()
    if rank == 0:
        print("NCCL_OK")
dist.destroy_process_group()
'''


_RANSAC_CASES = [
    # n, inlier ratio, min trials, max trials, scene seed, PRUNE_MIN
    (3000, 0.45, 2048, 10000, 1, 128),    # several waves, pruned second phases, adaptive abort
    (12000, 0.3, 2500, 2500, 7, 2048),    # fixed trial count
    (2000, 0.35, 0, 10000, 5, 128),       # abort inside the first wave
    (777, 0.5, 1100, 3000, 9, 128),       # ragged sizes
    (20000, 0.3, 6000, 6000, 11, 2048),   # more than one model block per rank
]

_RANSAC_WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np
import torch
import torch.distributed as dist
import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import RANSACOptions, synthetic as S

dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
torch.cuda.set_device(rank)
ctx = pp.Context(rank)
ctx.comm_init_from_torch(dist)
out = []
for n, ratio, tmin, tmax, seed, prune_min in {cases!r}:
    os.environ["PPSFM_RANSAC_PRUNE_MIN"] = str(prune_min)
    sc = S.make_abs_pose_scene(n=n, inlier_ratio=ratio, seed=seed)
    o = RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                      min_num_trials=tmin, max_num_trials=tmax)
    ctx.set_prng_seed(0)
    rep, mask = ctx.ransac_p6l_sharded(sc["lines"], sc["aligned"], sc["points"], o)
    tm = ctx.ransac_timing()
    out.append(dict(success=int(rep.success), num_trials=int(rep.num_trials),
                    num_inliers=int(rep.num_inliers), residual_sum=float(rep.residual_sum).hex(),
                    best=(int(rep.best_trial), int(rep.best_model_idx)),
                    model=[float(x).hex() for x in rep.model],
                    scored=int(rep.num_models_scored), mask=mask.tobytes().hex(),
                    prng=int(ctx.prng_peek()), pairs=int(tm.score_pairs)))
with open(os.path.join(sys.argv[2], f"ransac{{rank}}.json"), "w") as f:
    json.dump(out, f)
dist.barrier()
if rank == 0:
    print("RANSAC_NCCL_OK")
dist.destroy_process_group()
'''


def _launch(backend, nproc, tmp_path, extra=(), worker=None):
    script = tmp_path / "worker.py"
    script.write_text(worker if worker is not None else _WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script), backend, *extra]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def test_ba_sharding_partitions_the_problem_gloo_world2(tmp_path):
    r = _launch("gloo", 2, tmp_path)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_shard_stats_single_process():
    from privacy_preserving_sfm_b200 import bundle_adjustment as ba, synthetic as S
    sc = S.make_ba_scene(num_cams=5, num_points=50, obs_per_point=3, seed=1)
    arr = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                      sc["obs_line"], [1], [sc["cam_params"]])
    tot = ba.shard_stats(arr, 0, 1)
    assert tot == (150, 50, 5, 150)
    parts = [ba.shard_stats(arr, r, 3) for r in range(3)]
    assert sum(p[0] for p in parts) == 150 and sum(p[1] for p in parts) == 50
    assert all(p[2] == 5 and p[3] == 150 for p in parts)


def test_ransac_shard_accounting():
    """Host side of the sharded RANSAC call: the model blocks (512 models) are dealt round-robin,
    every model belongs to exactly one rank, shares differ by at most one block."""
    import privacy_preserving_sfm_b200 as pp
    for K in (0, 1, 511, 512, 513, 4000, 37712):
        for world in (1, 2, 3, 8):
            own = [pp.binding.ransac_shard_models(K, r, world) for r in range(world)]
            assert sum(own) == K
            assert max(own) - min(own) <= 512


@pytest.mark.gpu
def test_sharded_ransac_call_matches_single_gpu(tmp_path, ctx, monkeypatch):
    """ONE RANSAC call sharded over 2 GPUs: every rank returns the report, mask and
    generator state of the single-GPU call, bit for bit, and scores about half of the pairs."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import privacy_preserving_sfm_b200 as pp
    from privacy_preserving_sfm_b200 import RANSACOptions, synthetic as S
    out = tmp_path / "out"
    out.mkdir()
    r = _launch("nccl", 2, tmp_path, extra=(str(out),),
                worker=_RANSAC_WORKER.format(root=ROOT, cases=_RANSAC_CASES))
    assert r.returncode == 0 and "RANSAC_NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    got = [json.load(open(out / f"ransac{rank}.json")) for rank in range(2)]
    for i, (n, ratio, tmin, tmax, seed, prune_min) in enumerate(_RANSAC_CASES):
        monkeypatch.setenv("PPSFM_RANSAC_PRUNE_MIN", str(prune_min))
        sc = S.make_abs_pose_scene(n=n, inlier_ratio=ratio, seed=seed)
        o = RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                          min_num_trials=tmin, max_num_trials=tmax)
        ctx.set_prng_seed(0)
        rep, mask = ctx.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], o)
        pairs1 = ctx.ransac_timing().score_pairs
        for rank in range(2):
            g = got[rank][i]
            assert g["success"] == rep.success and g["num_trials"] == rep.num_trials, (i, rank)
            assert g["num_inliers"] == rep.num_inliers, (i, rank)
            assert g["residual_sum"] == float(rep.residual_sum).hex(), (i, rank)
            assert tuple(g["best"]) == (rep.best_trial, rep.best_model_idx), (i, rank)
            assert g["model"] == [float(x).hex() for x in rep.model], (i, rank)
            assert g["scored"] == rep.num_models_scored, (i, rank)
            assert g["mask"] == mask.tobytes().hex(), (i, rank)
            assert g["prng"] == ctx.prng_peek(), (i, rank)
        if n >= 12000:   # the work is really split: each rank evaluates about half of the pairs
            assert got[0][i]["pairs"] + got[1][i]["pairs"] <= 1.1 * pairs1
            assert max(got[0][i]["pairs"], got[1][i]["pairs"]) <= 0.75 * pairs1


@pytest.mark.gpu
def test_sharded_ba_matches_single_gpu(tmp_path, ctx):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from privacy_preserving_sfm_b200 import bundle_adjustment as ba, synthetic as S
    out = tmp_path / "out"
    out.mkdir()
    r = _launch("nccl", 2, tmp_path, extra=(str(out),))
    assert r.returncode == 0 and "NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    sc = S.make_ba_scene(num_cams=9, num_points=401, obs_per_point=5, seed=7)
    flags = np.zeros(9, np.uint8)
    flags[0], flags[1], flags[8] = 1, 2, 1
    pc = (np.arange(401) % 11 == 0).astype(np.uint8)
    arr = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                      sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags, point_const=pc)
    ok, s = ba.solve_arrays(ctx, arr, ba.default_solver_options(max_num_iterations=15,
                                                                gradient_tolerance=1e-3))
    for rank in range(2):
        c = np.load(out / f"c{rank}.npy")
        assert abs(c[1] - s.final_cost) <= 1e-9 * s.final_cost
        assert (int(c[2]), int(c[3])) == (s.num_successful_steps, s.num_unsuccessful_steps)
        assert np.abs(np.load(out / f"q{rank}.npy") - arr.qvecs).max() < 1e-8
        assert np.abs(np.load(out / f"t{rank}.npy") - arr.tvecs).max() < 1e-8
        assert np.abs(np.load(out / f"X{rank}.npy") - arr.points).max() < 1e-7
    # the sharded solve with intrinsics refinement (two variable SIMPLE_RADIAL cameras)
    sn = S.make_ba_scene(num_cams=9, num_points=401, obs_per_point=5, seed=7, noise_px=2.0)
    prm = [[1000.0, 500, 500, 0.08], [1000.0, 500, 500, 0.05]]
    ai = ba.BaArrays(sn["qvecs"], sn["tvecs"], sn["points"], sn["obs_cam"], sn["obs_pt"],
                     sn["obs_line"], [2, 2], prm, image_camera=np.arange(9) % 2, pose_flags=flags,
                     point_const=pc)
    ok, si = ba.solve_arrays(ctx, ai, ba.default_solver_options(max_num_iterations=6,
                                                                refine_extra_params=1))
    assert ai.camera_params[0, 3] != 0.08 and ai.camera_params[1, 3] != 0.05
    for rank in range(2):
        c = np.load(out / f"ic{rank}.npy")
        assert abs(c[0] - si.initial_cost) <= 1e-12 * si.initial_cost
        assert abs(c[1] - si.final_cost) <= 1e-8 * si.final_cost
        assert (int(c[2]), int(c[3])) == (si.num_successful_steps, si.num_unsuccessful_steps)
        pr = np.load(out / f"ip{rank}.npy")
        assert (np.abs(pr - ai.camera_params) / np.maximum(np.abs(ai.camera_params), 1e-3)).max() < 1e-6
        assert np.abs(np.load(out / f"iq{rank}.npy") - ai.qvecs).max() < 1e-7
        assert np.abs(np.load(out / f"iX{rank}.npy") - ai.points).max() < 1e-6
