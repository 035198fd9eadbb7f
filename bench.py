#!/usr/bin/env python
"""bench.py — headline benchmark of the privacy-preserving-SfM hot path on B200.

Primary metric (BASELINE.json configs[1]): absolute-pose P6L RANSAC hypotheses/s; one call =
50 000 synthetic line<->point correspondences x 10 000 hypotheses (every hypothesis solved for
<= 8 poses, every pose scored on all correspondences); one step = CALLS_PER_STEP such calls on
distinct scenes per GPU (weak scaling: every rank registers its own images).  Secondary objects:
`ransac_sharded_call` (ONE call sharded over the GPUs, strong scaling),
`ransac_mapper_call` (the mapper's adaptive settings on 2 000 correspondences), `ba` (config 4:
500 cameras / 200 k points / 2 M observations, LM iterations/s, strong scaling over NCCL) and
`ba_config3` (100 cameras / 30 k points / 300 k observations, one GPU), `init_config1` (config 1:
the four-view initialisation on 2 000 tracks, host vs candidate models scored on the GPU).

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
ALG_FLOP_PER_PAIR = 28.0  # 27 FP64 flop + 1 FP64 divide per (model, correspondence)
EXEC_FLOP_PER_PAIR = 26.0  # what score_kernel executes: 13 FMAs per pair in its float stage
BYTES_PER_CORR = 48.0  # 6 doubles per correspondence per pass
CALLS_PER_STEP = 32    # a step = a batch of independent calls on distinct scenes (>= 1 s timed)


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
    # stdout carries ONE JSON line.  NCCL prints its INFO / VERSION log to stdout unless
    # NCCL_DEBUG_FILE is set, so the log is sent to stderr instead of being switched off: a
    # driver that sets NCCL_DEBUG=INFO still sees the communicator's rank count there.
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
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


def scene_seed(rank, call):
    """Scene of call `call` on rank `rank`; (0, 0) is the canonical config-2 scene of the parity
    tests (tests/test_gpu_parity_configs.py)."""
    from privacy_preserving_sfm_b200 import synthetic as S
    return S.SCENE_SEED + 1000 * rank + call


def make_scene(rank=0, call=0):
    from privacy_preserving_sfm_b200 import synthetic as S
    return S.make_abs_pose_scene(n=N_CORR, inlier_ratio=0.30, noise_px=1.0, focal=1000.0,
                                 aligned_fraction=0.30, seed=scene_seed(rank, call))


def workload_config(world):
    """The SAME dict on both arms (the driver compares them)."""
    return {"workload": "absolute-pose P6L RANSAC (BASELINE.json configs[1]): one call = 50k "
                        "lifted-line correspondences x 10k hypotheses, every model scored on all "
                        "correspondences; one step = %d such calls on distinct scenes per GPU"
                        % CALLS_PER_STEP,
            "n_correspondences": N_CORR, "hypotheses_per_call": N_HYP,
            "calls_per_step_per_gpu": CALLS_PER_STEP,
            "hypotheses_per_step": N_HYP * CALLS_PER_STEP * world,
            "inlier_ratio": 0.30, "max_error": MAX_ERROR,
            "l2": "flushed between timed steps; every call of a step reads its own scene",
            "parallelism": f"{world} rank(s), each registering its own scenes (distinct seeds); "
                           "no data-path collective"}


def cpu_arm():
    """The CPU implementation the baselines time: ("reference", module) when oracle/_ref/
    libref_p6l.so exists — the reference's OWN RANSAC loop / P6L / re3q3 / scoring sources,
    compiled in the build container by oracle/build_ref.sh against Eigen / glog stand-ins and
    shipped as a git-ignored file — else ("port", module): the oracle's restatement."""
    import oracle as O
    try:
        import oracle.reference as R
        if os.path.exists(R.LIB_PATH):
            R.lib()
            return "reference", R
    except (OSError, RuntimeError):
        pass
    return "port", O


CPU_ARM_TEXT = {
    "reference": "the reference's own sources (src/optim/ransac.h, src/estimators/absolute_pose.cc, "
                 "lib/re3q3, src/estimators/utils.cc) compiled by oracle/build_ref.sh against "
                 "stand-ins for Eigen / glog (absent in this image); serial loop, 1 thread, as the "
                 "reference runs it",
    "port": "oracle restatement of the reference's serial RANSAC loop, 1 thread like the "
            "reference (oracle/_ref/libref_p6l.so not present)",
}


def cpu_reference_run(sc, num_trials, seed=0):
    """`num_trials` trials of the reference's serial CPU loop (1 thread, as the reference) on
    the whole correspondence set.  min_inlier_ratio 0.01 keeps the constructor's cap
    (src/optim/ransac.h:144-156) far above num_trials; min = max = num_trials disables the
    dynamic abort, as in config 2."""
    import oracle as O
    kind, impl = cpu_arm()
    impl.set_prng_seed(seed)
    opt = O.make_options(MAX_ERROR, 0.01, 0.99999, 3.0, num_trials, num_trials)
    t0 = time.perf_counter()
    rep, _ = impl.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], opt)
    dt = time.perf_counter() - t0
    assert rep.num_trials == num_trials
    return num_trials / dt, dt, kind


def cpu_all_cores_run(sc, trials_per_thread):
    """For information: the same serial loop on every host core at once, each thread on its own
    slice of trials (its own thread-local generator) — the ceiling of an embarrassingly parallel
    CPU version the reference does not have.  Needs the reference build (its PRNG is
    thread_local; the oracle's is one global)."""
    import threading
    kind, impl = cpu_arm()
    if kind != "reference":
        return None
    cores = os.cpu_count() or 1
    done = []

    def work(tid):
        done.append(cpu_reference_run(sc, trials_per_thread, seed=tid + 1)[1])

    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return {"value": cores * trials_per_thread / dt, "unit": "hypotheses/s", "cores": cores,
            "sample": f"{trials_per_thread} trials per thread on all {N_CORR} correspondences, "
                      f"{dt:.1f} s; NOT a reference code path (its loop is serial)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = make_scene()
    # bounded sample of a step's 320 000 hypotheses: the first 1 000 trials of the step's first
    # call (~1.3 s on one core; the serial loop's cost per trial does not depend on the trial)
    sample_trials = 1000
    for _ in range(args.warmup):
        cpu_reference_run(sc, 100)
    total_t, total_h = 0.0, 0
    kind = "port"
    for _ in range(args.steps):
        hps, dt, kind = cpu_reference_run(sc, sample_trials)
        total_t += dt
        total_h += sample_trials
    value = total_h / total_t
    out = {
        "impl": "reference", "metric": "ransac_hypotheses_per_sec", "value": value,
        "unit": "hypotheses/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "hypotheses/s", "cores": 1, "kind": kind,
                         "sample": f"{sample_trials} of the {N_HYP * CALLS_PER_STEP} hypotheses "
                                   f"of a step (first call, all {N_CORR} correspondences) per "
                                   "timed step; " + CPU_ARM_TEXT[kind]},
        "e2e": {"value": value, "unit": "hypotheses/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count(),
    }
    allc = cpu_all_cores_run(sc, 200)
    if allc is not None:
        out["cpu_baseline"]["all_cores_for_information"] = allc
    print(json.dumps(out), flush=True)


def reduce_max(dist, x):
    if dist is None:
        return float(x)
    import torch
    t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def pinned(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).pin_memory()
    return t, t.numpy()


def ncu_traffic(kernel):
    """dram bytes per launch from this round's `ncu --set full` capture, if one is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel)
    return None


def run_b200(args):
    rank, world, local, dist = dist_setup(args.gpus)
    import privacy_preserving_sfm_b200 as pp
    ctx = pp.Context(local)
    if dist is not None:
        ctx.comm_init_from_torch(dist)      # NCCL communicator: sharded RANSAC call, sharded BA
    scenes = [make_scene(rank, c) for c in range(CALLS_PER_STEP)]
    opts = pp.RANSACOptions(max_error=MAX_ERROR, min_inlier_ratio=0.25, confidence=0.99999,
                            dyn_num_trials_multiplier=3.0, min_num_trials=N_HYP,
                            max_num_trials=N_HYP)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---------------- kernel-side metric: correspondences resident in HBM -----------------
    corrs = [ctx.upload(sc["lines"], sc["aligned"], sc["points"]) for sc in scenes]
    mask = np.zeros(N_CORR, dtype=np.uint8)
    for _ in range(args.warmup):
        for c, corr in enumerate(corrs):
            ctx.set_prng_seed(c)
            ctx.ransac_p6l_resident(corr, opts, mask_out=mask)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    step_s, score_ms, solve_ms, exact_ms, pairs, launches, score_launches = [], 0.0, 0.0, 0.0, 0, 0, 0
    rep0, models_scored = None, 0
    for _ in range(args.steps):
        ctx.bench_l2_flush()           # evict L2 between timed steps (untimed)
        barrier()
        t0 = time.perf_counter()
        for c, corr in enumerate(corrs):
            ctx.set_prng_seed(c)
            rep, _ = ctx.ransac_p6l_resident(corr, opts, mask_out=mask)   # blocking call
            tm = ctx.ransac_timing()
            score_ms += tm.score_ms
            solve_ms += tm.solve_ms
            exact_ms += tm.exact_ms
            pairs += tm.score_pairs
            launches += tm.kernel_launches
            score_launches += tm.score_launches
            if c == 0:
                rep0, models_scored = rep, int(rep.num_models_scored)
        step_s.append(time.perf_counter() - t0)
    clocks = sampler.stop()
    total = reduce_max(dist, sum(step_s))
    value = world * N_HYP * CALLS_PER_STEP * args.steps / total

    # ---------------- end-to-end metric: host buffers through the public API ---------------
    keep = [(pinned(sc["lines"]), pinned(sc["aligned"]), pinned(sc["points"])) for sc in scenes]
    for _ in range(max(1, min(args.warmup, 2))):
        for c, (hl, ha, hp) in enumerate(keep):
            ctx.set_prng_seed(c)
            ctx.ransac_p6l(hl[1], ha[1], hp[1], opts)
    e2e_s = []
    for _ in range(args.steps):
        ctx.bench_l2_flush()
        barrier()
        t0 = time.perf_counter()
        for c, (hl, ha, hp) in enumerate(keep):
            ctx.set_prng_seed(c)
            rep_e, mask_e = ctx.ransac_p6l(hl[1], ha[1], hp[1], opts)
        e2e_s.append(time.perf_counter() - t0)
    e2e_total = reduce_max(dist, sum(e2e_s))
    e2e_value = world * N_HYP * CALLS_PER_STEP * args.steps / e2e_total
    hl, ha, hp = keep[0]
    h2d = CALLS_PER_STEP * (hl[1].nbytes + ha[1].nbytes + hp[1].nbytes + N_HYP * 6 * 4)
    d2h = CALLS_PER_STEP * (N_CORR + (N_HYP + 1) * 4 + 8 * N_HYP * 4 + 256)
    del keep

    sharded = run_sharded_call(args, ctx, world, rank, dist, scenes[0] if rank == 0 else make_scene(),
                               opts)
    mapper = run_mapper_call(args, ctx, rank)
    init1 = run_init_config1(args, ctx) if rank == 0 else None
    for corr in corrs:
        corr.free()

    ba_out = ba3_out = ba_weak = None
    if not args.no_ba:
        ba_out = run_ba_b200(args, ctx, world, rank, dist, 500, 200000, "configs[3]", local)
        if world == 1:
            ba3_out = run_ba_b200(args, ctx, 1, 0, None, 100, 30000, "configs[2]", local)
        else:
            # weak scaling of the sharded solve: 200k points (2M observations) PER GPU, the same
            # 500 cameras — per-GPU work as at N = 1, plus the all-reduce of the reduced system
            ba_weak = run_ba_b200(args, ctx, world, rank, dist, 500, 200000 * world,
                                  "configs[3] x %d points" % world, local, weak=True)
    if rank != 0:
        return

    # ---------------- roofline of the dominant kernel (scoring) ----------------------------
    dfma_tips, dmuladd_tips = ctx.bench_fp64_peak()
    ffma_tips = ctx.bench_fp32_peak()
    hbm_peak, peak_kind = peaks()
    score_s = score_ms * 1e-3
    pairs_per_s = pairs / score_s
    calls = args.steps * CALLS_PER_STEP
    models_per_call = pairs / calls / N_CORR
    passes = np.ceil(models_per_call / 512.0)   # one pass over the set per block of 512 models
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    arith_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12     # FP32 lanes x 2 flop x clock
    achieved = pairs_per_s * EXEC_FLOP_PER_PAIR / 1e12
    traffic = ncu_traffic("score_kernel")
    roofline = {
        "kernel": "score_kernel",
        # CUDA-core FMA bound: neither HBM nor tensor pipe.  The kernel decides
        # all but ~1e-5 of the pairs in its float stage: 13 packed FMAs = 26 flop per pair.
        "bound": "fp32_fma",
        "achieved": achieved, "peak": 2 * ffma_tips,
        "unit": "TFLOP/s",
        "unit_note": "EXECUTED flop: 26 per (model, correspondence) pair actually evaluated, pruned "
                     "pairs not counted",
        "bound_note": "CUDA-core packed-float FMA pipe (SURVEY.md 8d: the scoring is ALU-bound, "
                      "neither HBM nor tensor pipe); the HBM view of the same kernel is under 'hbm'",
        "frac": achieved / (2 * ffma_tips),
        "peak_source": "measured in this run (ppsfm_bench_fp32_peak: packed FFMA2 issue rate x 2 "
                       "flop); not in MEASURED_PEAKS.json",
        "arithmetic_fp32_peak": arith_peak,
        "frac_of_arithmetic_fp32_peak": achieved / arith_peak,
        "algorithmic": {"flop_per_pair": ALG_FLOP_PER_PAIR,
                        "tflops": pairs_per_s * ALG_FLOP_PER_PAIR / 1e12,
                        "frac_of_fp64_peak": pairs_per_s * ALG_FLOP_PER_PAIR / 1e12 / (2 * dfma_tips),
                        "fp64_peak_tflops": 2 * dfma_tips,
                        "fp64_unfused_peak_tops": dmuladd_tips,
                        "note": "SURVEY.md 8d's figure (27 flop + 1 divide in FP64); above 1 "
                                "because the float filter stage replaces the FP64 evaluation"},
        "pairs_per_s": pairs_per_s,
        "launches_per_call": score_launches / max(1, calls),
        "avg_launch_ms": score_ms / max(1, score_launches),
        "hbm": {"achieved": passes * N_CORR * BYTES_PER_CORR / (score_s / calls) / 1e9,
                "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_kind,
                "note": "algorithmic bytes = passes x N x 48 B; compute-bound kernel"},
        "traffic": traffic["bytes_per_launch"] if traffic else None,
        "traffic_source": traffic["source"] if traffic else
                          "no ncu --set full capture of this build committed",
    }

    cfg = workload_config(world)
    out = {
        "metric": "ransac_hypotheses_per_sec", "value": value, "unit": "hypotheses/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "ms_per_call": 1e3 * total / calls,
        "ms_min_step": 1e3 * min(step_s),
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "timing": "host clock around the blocking C-ABI calls of a step (>= device time: "
                  "sampling, replay and result copies included), max over ranks",
        "e2e": {"value": e2e_value, "unit": "hypotheses/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_total / args.steps,
                "ms_per_call": 1e3 * e2e_total / calls},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernel_ms_per_call": {"solve": solve_ms / calls, "score": score_ms / calls,
                               "exact": exact_ms / calls},
        "result": {"scene": "call 0 of rank 0 (the parity-tested config-2 scene)",
                   "num_inliers": int(rep0.num_inliers), "num_trials": int(rep0.num_trials),
                   "best_trial": int(rep0.best_trial), "models_scored": models_scored},
        "host_cores": os.cpu_count(),
        "ransac_sharded_call": sharded,
        "ransac_mapper_call": mapper,
        "init_config1": init1,
    }
    if ba_out is not None:
        out["ba"] = ba_out
    if ba3_out is not None:
        out["ba_config3"] = ba3_out
    if ba_weak is not None:
        out["ba_weak_scaling"] = ba_weak
    if not args.no_cpu_baseline and world == 1:
        hps, dt, kind = cpu_reference_run(scenes[0], 2000)
        out["cpu_baseline"] = {
            "value": hps, "unit": "hypotheses/s", "cores": 1, "kind": kind,
            "sample": f"2000 of the {N_HYP} hypotheses of one call on all {N_CORR} "
                      f"correspondences, {dt:.1f} s; " + CPU_ARM_TEXT[kind]}
        allc = cpu_all_cores_run(scenes[0], 200)
        if allc is not None:
            out["cpu_baseline"]["all_cores_for_information"] = allc
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# ONE call sharded over the GPUs (RANSAC row): strong scaling of a single
# EstimateAbsolutePoseFromLines-sized call, parity-checked against the single-GPU call.
# ------------------------------------------------------------------------------------------------
def run_sharded_call(args, ctx, world, rank, dist, sc, opts):
    corr = ctx.upload(sc["lines"], sc["aligned"], sc["points"])
    n_calls = max(args.steps, 1) * 4
    # single-GPU answer of the same call (untimed), on every rank
    ctx.set_prng_seed(0)
    ref, ref_mask = ctx.ransac_p6l_resident(corr, opts)
    ref_peek = ctx.prng_peek()
    for _ in range(3):
        ctx.set_prng_seed(0)
        rep, mask = ctx.ransac_p6l_resident_sharded(corr, opts)
    same = (rep.num_trials == ref.num_trials and rep.num_inliers == ref.num_inliers and
            (rep.best_trial, rep.best_model_idx) == (ref.best_trial, ref.best_model_idx) and
            list(rep.model) == list(ref.model) and rep.residual_sum == ref.residual_sum and
            np.array_equal(mask, ref_mask) and ctx.prng_peek() == ref_peek)
    ts, comm_ms, pairs = [], 0.0, 0
    for _ in range(n_calls):
        ctx.bench_l2_flush()
        ctx.set_prng_seed(0)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        ctx.ransac_p6l_resident_sharded(corr, opts)
        ts.append(time.perf_counter() - t0)
        tm = ctx.ransac_timing()
        comm_ms += tm.comm_ms
        pairs += tm.score_pairs
    total = reduce_max(dist, sum(ts))
    hl, ha, hp = pinned(sc["lines"]), pinned(sc["aligned"]), pinned(sc["points"])
    for _ in range(2):
        ctx.set_prng_seed(0)
        ctx.ransac_p6l_sharded(hl[1], ha[1], hp[1], opts)
    te = []
    for _ in range(n_calls):
        ctx.bench_l2_flush()
        ctx.set_prng_seed(0)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        ctx.ransac_p6l_sharded(hl[1], ha[1], hp[1], opts)
        te.append(time.perf_counter() - t0)
    e2e_total = reduce_max(dist, sum(te))
    corr.free()
    ok = same
    if dist is not None:
        import torch
        t = torch.tensor([1.0 if same else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item() > 0.5)
    return {"metric": "ransac_hypotheses_per_sec (one call)", "value": N_HYP * n_calls / total,
            "unit": "hypotheses/s", "n_gpus": world, "scaling": "strong",
            "ms_per_call": 1e3 * total / n_calls, "ms_min": 1e3 * min(ts), "calls_timed": n_calls,
            "e2e": {"value": N_HYP * n_calls / e2e_total, "unit": "hypotheses/s",
                    "ms_per_call": 1e3 * e2e_total / n_calls},
            "allreduce_ms_per_call": comm_ms / n_calls,
            "pairs_scored_per_call_this_rank": pairs / n_calls,
            "parity_vs_single_gpu_call": ok,
            "parallelism": ("single GPU" if world == 1 else
                            f"ONE call: every rank samples and solves all hypotheses, scores 1/{world} "
                            "of each wave's models; NCCL all-reduce of the 32-bit counts per wave; "
                            "every rank replays the reference loop"),
            "config": "50k correspondences x 10k hypotheses, seed 0, L2 flushed before every call"}


# ------------------------------------------------------------------------------------------------
# The mapper's call shape (src/sfm/incremental_mapper.cc:673-681): a few thousand correspondences,
# min 100 / max 10 000 trials, adaptive abort.  Latency bound: the solve kernel is the cost.
# ------------------------------------------------------------------------------------------------
def run_init_config1(args, ctx):
    """BASELINE.json configs[0]: init::initialize_reconstruction on 2 000 lifted-line
    tracks (1 000 aligned), 10 % outliers, max_error 0.005: the host run (the reference's CPU path,
    restated) against the run whose candidate models are scored on the GPU; the two must agree
    bit for bit."""
    from privacy_preserving_sfm_b200 import initializer as I, synthetic as S
    lines, aligned, gravity, gt = S.make_init_scene(2000, 1000, 200, seed=S.SCENE_SEED)
    opt = I.InitOptions(max_error=0.005)
    I.initialize_reconstruction(lines[:, :200], aligned[:, :200], gravity, opt, ctx=ctx)  # warm-up
    t0 = time.perf_counter()
    okg, pg, rg, repg = I.initialize_reconstruction(lines, aligned, gravity, opt, ctx=ctx)
    t1 = time.perf_counter()
    out = {"metric": "four-view initialisation (config 1), wall time per run", "unit": "ms",
           "gpu_scored_ms": 1e3 * (t1 - t0), "gpu_launches": int(repg.gpu_launches),
           "iterations": [int(repg.iterations_2d), int(repg.iterations_3d)],
           "inlier_ratio": rg,
           "max_pose_error": float(np.abs(_normalise_init(pg) - gt).max()) if okg else None,
           "config": {"tracks": 2000, "aligned": 1000, "outliers": 200, "max_error": 0.005}}
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        okh, ph, rh, reph = I.initialize_reconstruction(lines, aligned, gravity, opt)
        out["host_ms"] = 1e3 * (time.perf_counter() - t0)
        out["identical_to_host_run"] = bool(okh == okg and rh == rg and np.array_equal(ph, pg))
        assert out["identical_to_host_run"], "GPU-scored initialisation differs from the host run"
    return out


def _normalise_init(poses):
    p = poses.copy()
    p[:, :, 3] /= np.linalg.norm(p[1, :, 3])
    return p


def run_mapper_call(args, ctx, rank):
    import privacy_preserving_sfm_b200 as pp
    from privacy_preserving_sfm_b200 import synthetic as S
    n = 2000
    sc = S.make_abs_pose_scene(n=n, inlier_ratio=0.5, noise_px=1.0, focal=1000.0,
                               aligned_fraction=0.30, seed=S.SCENE_SEED + 7)
    o = pp.RANSACOptions(max_error=MAX_ERROR, min_inlier_ratio=0.25, confidence=0.99999,
                         dyn_num_trials_multiplier=3.0, min_num_trials=100, max_num_trials=10000)
    hl, ha, hp = pinned(sc["lines"]), pinned(sc["aligned"]), pinned(sc["points"])
    for _ in range(5):
        ctx.set_prng_seed(0)
        rep, mask = ctx.ransac_p6l(hl[1], ha[1], hp[1], o)
    n_calls = max(50, 10 * args.steps)
    ts, solve_ms, score_ms = [], 0.0, 0.0
    for i in range(n_calls):
        ctx.set_prng_seed(i)
        t0 = time.perf_counter()
        rep_i, _ = ctx.ransac_p6l(hl[1], ha[1], hp[1], o)
        ts.append(time.perf_counter() - t0)
        tm = ctx.ransac_timing()
        solve_ms += tm.solve_ms
        score_ms += tm.score_ms
    # the same call on a clean scene (95 % inliers, 500 correspondences): what the mapper driver
    # issues on synthetic data; most good models tie at the full inlier count
    sc2 = S.make_abs_pose_scene(n=500, inlier_ratio=0.95, noise_px=0.5, focal=1000.0,
                                aligned_fraction=0.40, seed=S.SCENE_SEED + 9)
    tc2 = []
    for i in range(5 + 40):
        ctx.set_prng_seed(i)
        t0 = time.perf_counter()
        ctx.ransac_p6l(sc2["lines"], sc2["aligned"], sc2["points"], o)
        if i >= 5:
            tc2.append(time.perf_counter() - t0)
    clean = {"n_correspondences": 500, "inlier_ratio": 0.95, "ms_per_call": 1e3 * sum(tc2) / len(tc2),
             "kernel_launches_per_call": int(ctx.ransac_timing().kernel_launches)}
    out = {"metric": "EstimateAbsolutePoseFromLines-shaped calls per second (host buffers in, "
                     "report + mask out)", "value": n_calls / sum(ts), "unit": "calls/s",
           "clean_scene_call": clean,
           "ms_per_call": 1e3 * sum(ts) / n_calls, "ms_median": 1e3 * float(np.median(ts)),
           "config": {"n_correspondences": n, "inlier_ratio": 0.5, "min_num_trials": 100,
                      "max_num_trials": 10000, "confidence": 0.99999,
                      "num_trials_seed0": int(rep.num_trials)},
           "kernel_ms_per_call": {"solve": solve_ms / n_calls, "score": score_ms / n_calls}}
    if rank == 0 and not args.no_cpu_baseline:
        import oracle as O
        oo = O.make_options(o.max_error, o.min_inlier_ratio, o.confidence,
                            o.dyn_num_trials_multiplier, o.min_num_trials, o.max_num_trials)
        kind, impl = cpu_arm()
        tc = []
        for i in range(20):
            impl.set_prng_seed(i)
            t0 = time.perf_counter()
            orep, omask = impl.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], oo)
            tc.append(time.perf_counter() - t0)
        out["cpu_baseline"] = {"value": len(tc) / sum(tc), "unit": "calls/s", "cores": 1,
                               "kind": kind, "ms_per_call": 1e3 * sum(tc) / len(tc),
                               "sample": "20 whole calls (seeds 0..19) of the serial loop; "
                                         + CPU_ARM_TEXT[kind]}
    return out


# ------------------------------------------------------------------------------------------------
# Line-reprojection bundle adjustment: BASELINE.json configs[3] (500 cameras / 200k points / 2M
# observations; strong scaling over NCCL) and configs[2] (100 / 30k / 300k; one GPU).
# ------------------------------------------------------------------------------------------------
BA_OBS_PER_POINT, BA_ITERS = 10, 10
BA_BYTES_PER_OBS = 216.0  # SURVEY.md §8(d), materialised Jacobian: 56 B read + 160 B written


def make_ba_problem(cams, points):
    from privacy_preserving_sfm_b200 import synthetic as S
    sc = S.make_ba_scene(num_cams=cams, num_points=points, obs_per_point=BA_OBS_PER_POINT,
                         seed=S.SCENE_SEED)
    flags = np.zeros(cams, np.uint8)
    flags[0] = 1   # image 0: constant pose; image 1: constant tvec[0]
    flags[1] = 2   # (src/sfm/incremental_mapper.cc:907-926)
    return sc, flags


def run_ba_b200(args, ctx, world, rank, dist, cams, points, config_name, local, weak=False):
    import privacy_preserving_sfm_b200 as pp
    from privacy_preserving_sfm_b200 import bundle_adjustment as ba
    sc, flags = make_ba_problem(cams, points)
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
    phase = {"schur": 0.0, "chol": 0.0, "backsub": 0.0}
    summ = None
    sampler = ClockSampler(local)
    sampler.start()
    solves = max(4 if weak else args.steps, 4 if cams >= 500 else 40)
    if cams < 500:
        solves = max(solves, 100)       # a config-3 solve takes ~7 ms: >= 0.5 s timed
    for _ in range(solves):
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
        phase["schur"] += summ.schur_time_s
        phase["chol"] += summ.cholesky_time_s
        phase["backsub"] += summ.backsub_time_s
    ba_clocks = sampler.stop()
    prob.download()
    prob.free()
    total = reduce_max(dist, sum(times))
    value = iters / total
    # sharded solve == single-GPU solve (rank 0 repeats it alone on a second context)
    parity = None
    if world > 1 and not weak:
        parity = {}
        if rank == 0:
            ctx1 = pp.Context(local)
            a1 = ba.BaArrays(*arr_args, pose_flags=flags)
            ok1, s1 = ba.solve_arrays(ctx1, a1, opts)
            ctx1.close()
            dq = float(np.abs(a1.qvecs - arrays.qvecs).max())
            dt = float(np.abs(a1.tvecs - arrays.tvecs).max())
            dX = float(np.abs(a1.points - arrays.points).max())
            dc = abs(s1.final_cost - summ.final_cost) / s1.final_cost
            parity = {"final_cost_rel_diff": dc, "max_abs_qvec_diff": dq, "max_abs_tvec_diff": dt,
                      "max_abs_point_diff": dX,
                      "same_step_sequence": (s1.num_successful_steps, s1.num_unsuccessful_steps) ==
                                            (summ.num_successful_steps, summ.num_unsuccessful_steps),
                      "ok": bool(dc <= 1e-9 and dq <= 1e-8 and dt <= 1e-8 and dX <= 1e-7)}
            assert parity["ok"], ("sharded BA differs from the single-GPU solve", parity)
    # end to end: host arrays in, host arrays out (assembly + H2D + solve + D2H)
    e2e_t, e2e_it = [], 0
    nbytes_in = sum(a.nbytes for a in (arrays.qvecs, arrays.tvecs, arrays.points, arrays.obs_image,
                                       arrays.obs_point, arrays.obs_line))
    nbytes_out = arrays.qvecs.nbytes + arrays.tvecs.nbytes + arrays.points.nbytes
    # host buffers in pinned memory (as the contract asks of the end-to-end leg); the in-place
    # outputs (poses, points) are re-initialised before every solve, outside the timed region
    keep = {k: pinned(sc[k], np.float64) for k in ("qvecs", "tvecs", "points", "obs_line")}
    keep.update({k: pinned(sc[k], np.int32) for k in ("obs_cam", "obs_pt")})
    for i in range((2 if weak else max(3, min(args.steps, 6))) + 1):
        for k in ("qvecs", "tvecs", "points"):
            keep[k][1][...] = sc[k]
        a2 = ba.BaArrays(keep["qvecs"][1], keep["tvecs"][1], keep["points"][1], keep["obs_cam"][1],
                         keep["obs_pt"][1], keep["obs_line"][1], [1], [sc["cam_params"]],
                         pose_flags=flags, copy=False)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        ok, s2 = ba.solve_arrays(ctx, a2, opts)
        dt = time.perf_counter() - t0
        if i > 0:   # first call is warm-up
            e2e_t.append(dt)
            e2e_it += s2.num_iterations
    e2e_total = reduce_max(dist, sum(e2e_t))
    hbm_peak, peak_kind = peaks()
    write_peak, read_peak = ctx.bench_hbm_rw_peak()
    K = len(sc["obs_cam"])
    K_rank = K / world   # points (with all their observations) are dealt round-robin to the ranks
    jac_avg_s = jac_s / max(1, jac_n)
    achieved = BA_BYTES_PER_OBS * K_rank / jac_avg_s / 1e9
    traffic = ncu_traffic("ba_linearize_kernel") if (world == 1 and cams >= 500) else None
    out = {
        "metric": "ba_lm_iterations_per_sec", "value": value, "unit": "LM iterations/s",
        "ms_per_iteration": 1e3 * total / max(1, iters), "steps": solves, "n_gpus": world,
        "scaling": "weak" if weak else "strong",
        "observation_iterations_per_s": value * len(sc["obs_cam"]),
        "parallelism": ("single GPU" if world == 1 else
                        f"points sharded over {world} GPUs, NCCL all-reduce of the reduced camera "
                        "system per LM iteration, replicated dense solve"),
        "iterations_per_step": iters / max(1, solves), "dtype": "f64",
        "config": {"workload": "line-reprojection BA, %d cams / %dk points / %d observations "
                               "(BASELINE.json %s), PINHOLE, TRIVIAL loss, gauge: cam 0 "
                               "constant, cam 1 tvec[0] constant" % (cams, points // 1000, K,
                                                                      config_name),
                   "cameras": cams, "points": points, "observations": int(K),
                   "lm_iterations_per_solve": BA_ITERS, "l2": "flushed between timed solves"},
        "e2e": {"value": e2e_it / e2e_total, "unit": "LM iterations/s",
                "ms_per_solve": 1e3 * e2e_total / len(e2e_t), "h2d_bytes_per_step": int(nbytes_in),
                "d2h_bytes_per_step": int(nbytes_out)},
        "gpu_launches": int(launches), "clocks": ba_clocks,
        "roofline": {"kernel": "ba_linearize_kernel<true> (Jacobian build)", "bound": "hbm",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "peak_source": peak_kind,
                     "bytes_per_obs": BA_BYTES_PER_OBS, "avg_launch_ms": 1e3 * jac_avg_s,
                     "traffic": traffic["bytes_per_launch"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None,
                     "note": "write-heavy kernel (160 of 216 B/obs are writes); measured in this "
                             "run: write-only HBM peak %.0f GB/s, read-only %.0f GB/s; the kernel "
                             "writes %.0f GB/s" % (write_peak, read_peak,
                                                   160.0 * K_rank / jac_avg_s / 1e9)},
        "phase_ms_per_iteration": {
            "jacobian_build": 1e3 * jac_avg_s,
            "reduced_system_incl_allreduce": 1e3 * phase["schur"] / max(1, iters),
            "cholesky_solve": 1e3 * phase["chol"] / max(1, iters),
            "backsubstitution_and_candidate_cost": 1e3 * phase["backsub"] / max(1, iters)},
        "result": {"initial_cost": summ.initial_cost, "final_cost": summ.final_cost,
                   "successful_steps": summ.num_successful_steps,
                   "unsuccessful_steps": summ.num_unsuccessful_steps},
    }
    if parity is not None:
        out["parity_vs_single_gpu"] = parity
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        out["cpu_baseline"] = ba_cpu_reference_run(sc, flags, 10)
    return out


def ba_cpu_reference_run(sc, flags, iters):
    """Oracle restatement of the reference's Ceres path on all host threads (the reference uses
    hardware_concurrency when residuals >= 50000, src/optim/bundle_adjustment.cc:288-301)."""
    import oracle as O
    a = O.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                   sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags)
    t0 = time.perf_counter()
    ok, s = O.ba_solve(a, O.ba_default_options(max_num_iterations=iters, num_threads=-1,
                                               gradient_tolerance=0.0, function_tolerance=0.0,
                                               parameter_tolerance=0.0))
    dt = time.perf_counter() - t0
    n = s.num_successful_steps + s.num_unsuccessful_steps
    return {"value": n / dt, "unit": "LM iterations/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} LM iterations of the same problem, {dt:.1f} s, all host threads; "
                      "oracle port with a DENSE Schur complement + dense Cholesky — NOT Ceres "
                      "SPARSE_SCHUR, which the reference would use here and which is expected to "
                      "be several times faster than this port; a reported baseline only"}


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
