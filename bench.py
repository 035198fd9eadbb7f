#!/usr/bin/env python
"""bench.py — headline benchmark of the privacy-preserving-SfM hot path on B200.

Primary metric (BASELINE.json configs[1]): absolute-pose P6L RANSAC hypotheses/s on 50 000
synthetic line<->point correspondences x 10 000 hypotheses per step (every hypothesis solved for
<= 8 poses, every pose scored on all correspondences).  A secondary object `ba` reports the
line-reprojection bundle adjustment (LM iterations/s) once that path is built.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One process per GPU (torchrun for N > 1).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CORR = 50000
N_HYP = 10000
MAX_ERROR = 12.0 / 1000.0
FLOP_PER_PAIR = 27.0   # 27 FP64 flop + 1 FP64 divide per (model, correspondence)
BYTES_PER_CORR = 48.0  # 6 doubles per correspondence per pass


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ba", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    NVML (nvidia_ml_py) is polled from a thread every 5 ms: the timed region of the RANSAC leg is
    tens of milliseconds, shorter than nvidia-smi's start-up, so a subprocess sampler misses it.
    The main thread sits in blocking C-ABI calls (GIL released) while the sampler runs.  If NVML
    cannot be loaded, one synchronous nvidia-smi query is taken at stop()."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40),
               ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.sm, self.mask, self.max_mhz = [], 0, None
        self.h = self.nv = self.th = None
        self.run = False

    def _handle(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv = pynvml
        try:   # NVML indices ignore CUDA_VISIBLE_DEVICES: go through the device UUID
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))

    def _loop(self):
        while self.run:
            try:
                self._sample()
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            self.h = self._handle()
            self.max_mhz = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.run = True
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception:
            self.h = None

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                  "-i", str(self.index)], capture_output=True, text=True,
                                 timeout=20).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            self.sm.append(float(f[0]))
            self.max_mhz = float(f[1])
            for (name, bit), v in zip(self.REASONS, f[2:6]):
                if v.lower().startswith("active"):
                    self.mask |= bit
            return "nvidia-smi (one sample after the timed region)"
        except Exception:
            return "unavailable"

    def stop(self):
        source = "nvml"
        if self.th is not None:
            self.run = False
            self.th.join(timeout=1)
        if not self.sm:
            source = self._smi_once()
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in self.REASONS if self.mask & bit],
                "samples": len(self.sm), "source": source}


def dist_setup(n_gpus):
    # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
    # and WARN, so the variable is removed unless PPSFM_NCCL_DEBUG asks for a level
    os.environ.pop("NCCL_DEBUG", None)
    if os.environ.get("PPSFM_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = os.environ["PPSFM_NCCL_DEBUG"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def make_scene():
    from privacy_preserving_sfm_b200 import synthetic as S
    return S.make_abs_pose_scene(n=N_CORR, inlier_ratio=0.30, noise_px=1.0, focal=1000.0,
                                 aligned_fraction=0.30, seed=S.SCENE_SEED)


def cpu_reference_run(sc, num_trials):
    """Oracle restatement of the reference's serial CPU loop (1 thread, as the reference)."""
    import oracle as O
    O.set_prng_seed(0)
    t0 = time.perf_counter()
    scored, rep = O.ransac_p6l_fixed_trials(sc["lines"], sc["aligned"], sc["points"], MAX_ERROR,
                                            num_trials)
    dt = time.perf_counter() - t0
    return num_trials / dt, dt, scored


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = make_scene()
    sample_trials = 1000  # bounded sample of the 10 000-hypothesis workload (~2 s per step)
    for _ in range(args.warmup):
        cpu_reference_run(sc, 100)
    total_t, total_h = 0.0, 0
    for _ in range(args.steps):
        hps, dt, _ = cpu_reference_run(sc, sample_trials)
        total_t += dt
        total_h += sample_trials
    value = total_h / total_t
    out = {
        "impl": "reference", "metric": "ransac_hypotheses_per_sec", "value": value,
        "unit": "hypotheses/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "absolute-pose P6L RANSAC, 50k lifted-line correspondences x 10k "
                               "hypotheses (BASELINE.json configs[1])",
                   "n_correspondences": N_CORR, "hypotheses_per_step": sample_trials},
        "cpu_baseline": {"value": value, "unit": "hypotheses/s", "cores": 1, "kind": "port",
                         "sample": f"{sample_trials} of {N_HYP} hypotheses per step on all "
                                   f"{N_CORR} correspondences; oracle restatement of the "
                                   "reference's serial RANSAC loop (Eigen/Ceres absent, the "
                                   "reference itself cannot be compiled here)"},
        "e2e": {"value": value, "unit": "hypotheses/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(out), flush=True)


def run_b200(args):
    rank, world, local, dist = dist_setup(args.gpus)
    import privacy_preserving_sfm_b200 as pp
    ctx = pp.Context(local)
    sc = make_scene()
    opts = pp.RANSACOptions(max_error=MAX_ERROR, min_inlier_ratio=0.25, confidence=0.99999,
                            dyn_num_trials_multiplier=3.0, min_num_trials=N_HYP,
                            max_num_trials=N_HYP)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---------------- kernel-side metric: correspondences resident in HBM -----------------
    corr = ctx.upload(sc["lines"], sc["aligned"], sc["points"])
    mask = np.zeros(N_CORR, dtype=np.uint8)
    for _ in range(args.warmup):
        ctx.set_prng_seed(0)
        ctx.ransac_p6l_resident(corr, opts, mask_out=mask)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    step_s, score_ms, solve_ms, exact_ms, pairs, launches, score_launches = [], 0.0, 0.0, 0.0, 0, 0, 0
    rep = None
    for _ in range(args.steps):
        ctx.bench_l2_flush()           # evict L2 between timed iterations (untimed)
        ctx.set_prng_seed(0)
        barrier()
        t0 = time.perf_counter()
        rep, _ = ctx.ransac_p6l_resident(corr, opts, mask_out=mask)   # blocking call
        step_s.append(time.perf_counter() - t0)
        tm = ctx.ransac_timing()
        score_ms += tm.score_ms
        solve_ms += tm.solve_ms
        exact_ms += tm.exact_ms
        pairs += tm.score_pairs
        launches += tm.kernel_launches
        score_launches += tm.score_launches
    clocks = sampler.stop()
    total = float(sum(step_s))
    if dist is not None:
        import torch
        t = torch.tensor([total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
    value = world * N_HYP * args.steps / total

    # ---------------- end-to-end metric: host buffers through the public API ---------------
    import torch
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = [pinned(sc["lines"]), pinned(sc["aligned"]), pinned(sc["points"])]
    hl, ha, hp = keep[0][1], keep[1][1], keep[2][1]
    for _ in range(max(1, args.warmup)):
        ctx.set_prng_seed(0)
        ctx.ransac_p6l(hl, ha, hp, opts)
    e2e_s = []
    for _ in range(args.steps):
        ctx.bench_l2_flush()
        ctx.set_prng_seed(0)
        barrier()
        t0 = time.perf_counter()
        rep_e, mask_e = ctx.ransac_p6l(hl, ha, hp, opts)
        e2e_s.append(time.perf_counter() - t0)
    e2e_total = float(sum(e2e_s))
    if dist is not None:
        t = torch.tensor([e2e_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    e2e_value = world * N_HYP * args.steps / e2e_total
    h2d = hl.nbytes + ha.nbytes + hp.nbytes + N_HYP * 6 * 4
    d2h = N_CORR + (N_HYP + 1) * 4 + int(rep.num_models_scored) * 4 + 256

    ba_out = None
    if not args.no_ba:
        if dist is not None:
            ctx.comm_init_from_torch(dist)      # NCCL communicator for the sharded BA
        ba_out = run_ba_b200(args, ctx, world, rank, dist)
    if rank != 0:
        return

    # ---------------- roofline of the dominant kernel (scoring) ----------------------------
    dfma_tips, dmuladd_tips = ctx.bench_fp64_peak()
    ffma_tips = ctx.bench_fp32_peak()
    hbm_peak, peak_kind = peaks()
    score_s = score_ms * 1e-3
    pairs_per_s = pairs / score_s
    models_per_step = pairs / args.steps / N_CORR
    passes = np.ceil(models_per_step / 256.0)  # one pass over all correspondences per 256 models
    roofline = {
        "kernel": "score_kernel",
        # CUDA-core FMA bound: neither HBM nor tensor pipe.  The kernel decides
        # all but a few pairs in a million in its float stage (13 packed FMAs per pair), so its
        # ceiling is the float FMA pipe; the FP64 numbers are kept for comparison.
        "bound": "fp32_fma",
        "achieved": pairs_per_s * (FLOP_PER_PAIR + 1) / 1e12,
        "peak": 2 * ffma_tips,
        "unit": "TFLOP/s (algorithmic 27 flop + 1 divide per (model, correspondence) pair, "
                "SURVEY.md 8d, against the measured packed-float FMA peak: the float stage of the "
                "filter executes 13 FMAs = 26 flop per pair, see DESIGN.md 2.4)",
        "frac": pairs_per_s * (FLOP_PER_PAIR + 1) / 1e12 / (2 * ffma_tips),
        "peak_source": "measured in this run (ppsfm_bench_fp32_peak, FFMA2 rate x 2 flop)",
        "fp64_peak_tflops": 2 * dfma_tips, "fp64_unfused_peak_tops": dmuladd_tips,
        "frac_of_fp64_peak": pairs_per_s * (FLOP_PER_PAIR + 1) / 1e12 / (2 * dfma_tips),
        "pairs_per_s": pairs_per_s,
        "launches_per_step": score_launches / max(1, args.steps),
        "avg_launch_ms": score_ms / max(1, score_launches),
        "hbm": {"achieved": passes * N_CORR * BYTES_PER_CORR / (score_s / args.steps) / 1e9,
                "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_kind,
                "note": "algorithmic bytes = passes x N x 48 B; compute-bound kernel"},
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch over all 10k hypotheses in
        # the committed ncu --set full capture (profiles/r01_s6_score_kernel_ncu.txt): the float
        # and double copies of the correspondence set are read from HBM about twice, the other
        # passes hit L2; scaled by the launches per step
        "traffic": 7.4e6 / max(1.0, score_launches / max(1, args.steps)),
    }

    out = {
        "metric": "ransac_hypotheses_per_sec", "value": value, "unit": "hypotheses/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "ms_min": 1e3 * min(step_s),
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "absolute-pose P6L RANSAC, 50k lifted-line correspondences x 10k "
                               "hypotheses per GPU per step (BASELINE.json configs[1])",
                   "n_correspondences": N_CORR, "hypotheses_per_step": N_HYP,
                   "models_scored_per_step": int(rep.num_models_scored),
                   "inlier_ratio": 0.30, "l2": "flushed between timed iterations",
                   "parallelism": f"{world} independent hypothesis batches (one per GPU), "
                                  "no data-path collective"},
        "e2e": {"value": e2e_value, "unit": "hypotheses/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_total / args.steps,
                "ms_min": 1e3 * min(e2e_s), "ms_median": 1e3 * float(np.median(e2e_s))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernel_ms_per_step": {"solve": solve_ms / args.steps, "score": score_ms / args.steps,
                               "exact": exact_ms / args.steps},
        "result": {"num_inliers": int(rep.num_inliers), "num_trials": int(rep.num_trials),
                   "best_trial": int(rep.best_trial)},
        "host_cores": os.cpu_count(),
    }
    if ba_out is not None:
        out["ba"] = ba_out
    if not args.no_cpu_baseline and world == 1:
        hps, dt, scored = cpu_reference_run(sc, 2000)
        out["cpu_baseline"] = {
            "value": hps, "unit": "hypotheses/s", "cores": 1, "kind": "port",
            "sample": f"2000 of {N_HYP} hypotheses ({scored} models) on all {N_CORR} "
                      f"correspondences, {dt:.1f} s; oracle restatement of the reference's serial "
                      "RANSAC loop (the reference needs Eigen/Ceres, absent here)"}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# Secondary metric: line-reprojection bundle adjustment, BASELINE.json configs[3]
# (500 cameras / 200k points / 2M observations), LM iterations per second.
# ------------------------------------------------------------------------------------------------
BA_CAMS, BA_POINTS, BA_OBS_PER_POINT, BA_ITERS = 500, 200000, 10, 10
BA_BYTES_PER_OBS = 216.0  # SURVEY.md §8(d), materialised Jacobian: 56 B read + 160 B written


def make_ba_problem():
    from privacy_preserving_sfm_b200 import synthetic as S
    sc = S.make_ba_scene(num_cams=BA_CAMS, num_points=BA_POINTS, obs_per_point=BA_OBS_PER_POINT,
                         seed=S.SCENE_SEED)
    flags = np.zeros(BA_CAMS, np.uint8)
    flags[0] = 1   # image 0: constant pose; image 1: constant tvec[0]
    flags[1] = 2   # (src/sfm/incremental_mapper.cc:907-926)
    return sc, flags


def run_ba_b200(args, ctx, world, rank, dist):
    from privacy_preserving_sfm_b200 import bundle_adjustment as ba
    sc, flags = make_ba_problem()
    arr_args = (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                sc["obs_line"], [1], [sc["cam_params"]])
    # global-BA options of the reference (src/controllers/incremental_mapper.cc:221-243) with a
    # fixed number of LM iterations so that every step does the same work
    opts = ba.default_solver_options(loss_type=0, max_num_iterations=BA_ITERS,
                                     gradient_tolerance=0.0, function_tolerance=0.0,
                                     parameter_tolerance=0.0)
    arrays = ba.BaArrays(*arr_args, pose_flags=flags)
    prob = ba.ResidentProblem(ctx, arrays, opts)
    for _ in range(max(1, args.warmup)):
        prob.reset()
        prob.run()
    times, iters, jac_s, jac_n, launches = [], 0, 0.0, 0, 0
    summ = None
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", 0)))
    sampler.start()
    for _ in range(args.steps):
        prob.reset()
        ctx.bench_l2_flush()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        ok, summ = prob.run()
        times.append(time.perf_counter() - t0)
        iters += summ.num_iterations
        jac_s += summ.jacobian_time_s
        jac_n += summ.jacobian_launches
        launches += summ.kernel_launches
    ba_clocks = sampler.stop()
    prob.free()
    total = float(sum(times))
    if dist is not None:
        import torch
        tt = torch.tensor([total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total = float(tt.item())
    value = iters / total
    # end to end: host arrays in, host arrays out (assembly + H2D + solve + D2H)
    e2e_t, e2e_it = [], 0
    nbytes_in = sum(a.nbytes for a in (arrays.qvecs, arrays.tvecs, arrays.points, arrays.obs_image,
                                       arrays.obs_point, arrays.obs_line))
    nbytes_out = arrays.qvecs.nbytes + arrays.tvecs.nbytes + arrays.points.nbytes
    # host buffers in pinned memory (as the contract asks of the end-to-end leg); the in-place
    # outputs (poses, points) are re-initialised before every solve, outside the timed region
    import torch
    def pinned(a, dtype):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).pin_memory()
        return t, t.numpy()
    keep = {k: pinned(sc[k], np.float64) for k in ("qvecs", "tvecs", "points", "obs_line")}
    keep.update({k: pinned(sc[k], np.int32) for k in ("obs_cam", "obs_pt")})
    for i in range(max(2, min(args.steps, 3)) + 1):
        for k in ("qvecs", "tvecs", "points"):
            keep[k][1][...] = sc[k]
        a2 = ba.BaArrays(keep["qvecs"][1], keep["tvecs"][1], keep["points"][1], keep["obs_cam"][1],
                         keep["obs_pt"][1], keep["obs_line"][1], [1], [sc["cam_params"]],
                         pose_flags=flags, copy=False)
        t0 = time.perf_counter()
        ok, s2 = ba.solve_arrays(ctx, a2, opts)
        dt = time.perf_counter() - t0
        if i > 0:   # first call is warm-up
            e2e_t.append(dt)
            e2e_it += s2.num_iterations
    hbm_peak, peak_kind = peaks()
    write_peak, read_peak = ctx.bench_hbm_rw_peak()
    K = len(sc["obs_cam"])
    K_rank = K / world   # points (with all their observations) are dealt round-robin to the ranks
    jac_avg_s = jac_s / max(1, jac_n)
    achieved = BA_BYTES_PER_OBS * K_rank / jac_avg_s / 1e9
    out = {
        "metric": "ba_lm_iterations_per_sec", "value": value, "unit": "LM iterations/s",
        "ms_per_iteration": 1e3 * total / max(1, iters), "steps": args.steps, "n_gpus": world,
        "scaling": "strong",
        "parallelism": ("single GPU" if world == 1 else
                        f"points sharded over {world} GPUs, NCCL all-reduce of the reduced camera "
                        "system per LM iteration, replicated dense solve"),
        "iterations_per_step": iters / max(1, args.steps), "dtype": "f64",
        "config": {"workload": "line-reprojection BA, 500 cams / 200k points / 2M observations "
                               "(BASELINE.json configs[3]), PINHOLE, TRIVIAL loss, gauge: cam 0 "
                               "constant, cam 1 tvec[0] constant", "cameras": BA_CAMS,
                   "points": BA_POINTS, "observations": int(K),
                   "lm_iterations_per_solve": BA_ITERS, "l2": "flushed between timed solves"},
        "e2e": {"value": e2e_it / sum(e2e_t), "unit": "LM iterations/s",
                "ms_per_solve": 1e3 * float(np.mean(e2e_t)), "h2d_bytes_per_step": int(nbytes_in),
                "d2h_bytes_per_step": int(nbytes_out)},
        "gpu_launches": int(launches), "clocks": ba_clocks,
        "roofline": {"kernel": "ba_linearize_kernel<true> (Jacobian build)", "bound": "hbm",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "peak_source": peak_kind,
                     "bytes_per_obs": BA_BYTES_PER_OBS, "avg_launch_ms": 1e3 * jac_avg_s,
                     # dram read 74 MB + write 260 MB per launch at N = 1 in the committed
                     # ncu --set full capture (profiles/r01_s2_ba_kernels_ncu.txt); the last
                     # ~100 MB of the 432 MB written are still dirty in the 126 MB L2 at kernel end
                     "traffic": 334e6 / world,
                     "note": "write-heavy kernel (160 of 216 B/obs are writes); measured in this "
                             "run: write-only HBM peak %.0f GB/s, read-only %.0f GB/s; the kernel "
                             "writes %.0f GB/s" % (write_peak, read_peak,
                                                   160.0 * K_rank / jac_avg_s / 1e9)},
        "phase_ms_per_iteration": {
            "jacobian_build": 1e3 * jac_avg_s,
            "reduced_system": 1e3 * summ.schur_time_s / max(1, summ.num_iterations),
            "cholesky_solve": 1e3 * summ.cholesky_time_s / max(1, summ.num_iterations),
            "backsubstitution_and_candidate_cost":
                1e3 * summ.backsub_time_s / max(1, summ.num_iterations)},
        "result": {"initial_cost": summ.initial_cost, "final_cost": summ.final_cost,
                   "successful_steps": summ.num_successful_steps,
                   "unsuccessful_steps": summ.num_unsuccessful_steps},
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = ba_cpu_reference_run(sc, flags, 2)
    return out


def ba_cpu_reference_run(sc, flags, iters):
    """Oracle restatement of the reference's Ceres path on all host threads (the reference uses
    hardware_concurrency when residuals >= 50000, src/optim/bundle_adjustment.cc:288-301)."""
    import oracle as O
    a = O.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                   sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags)
    t0 = time.perf_counter()
    ok, s = O.ba_solve(a, O.ba_default_options(max_num_iterations=iters, num_threads=-1))
    dt = time.perf_counter() - t0
    n = s.num_successful_steps + s.num_unsuccessful_steps
    return {"value": n / dt, "unit": "LM iterations/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} LM iterations of the same 500-camera problem, {dt:.1f} s; oracle "
                      "restatement (dense Schur + Cholesky) of the reference's Ceres path"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
