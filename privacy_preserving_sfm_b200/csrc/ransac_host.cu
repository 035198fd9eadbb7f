// ransac_host.cu — host side of the absolute-pose RANSAC path and the C-ABI entry points.
//
// RANSAC<P6LEstimator, InlierSupportMeasurer, RandomSampler>::Estimate
// (src/optim/ransac.h:144-278) is a serial loop with loop-carried state.  Here hypotheses are
// sampled on the host (the sampler must reproduce libstdc++'s mt19937 +
// uniform_int_distribution stream, src/util/random.h:88-128), solved and scored on the GPU in
// waves, and the loop-carried logic (best-so-far, dyn_max_num_trials, abort inside the model
// loop) is replayed literally on the host over per-model supports.
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

#include "comm.h"
#include "common.h"
#include "ransac_kernels.h"
#include "ransac_replay.h"

namespace ppsfm {
namespace {

using ppsfm::ComputeNumTrials;  // ransac_replay.h

// RandomSampler (src/optim/random_sampler.cc:40-62): persistent permutation, partial
// Fisher-Yates of the first 6 slots with RandomInteger<uint32_t>(i, n-1).
struct HostSampler {
  std::vector<uint32_t> idxs;
  void Initialize(size_t n) {
    idxs.resize(n);
    std::iota(idxs.begin(), idxs.end(), 0u);
  }
  void Sample(std::mt19937& prng, uint32_t* out) {
    const uint32_t last = static_cast<uint32_t>(idxs.size() - 1);
    for (uint32_t i = 0; i < 6; ++i) {
      std::uniform_int_distribution<uint32_t> distribution(i, last);
      const uint32_t j = distribution(prng);
      std::swap(idxs[i], idxs[j]);
    }
    for (int i = 0; i < 6; ++i) out[i] = idxs[i];
  }
  // Advance the generator exactly as `trials` calls of Sample() would, without a permutation.
  static void Skip(std::mt19937& prng, size_t n, size_t trials) {
    const uint32_t last = static_cast<uint32_t>(n - 1);
    for (size_t t = 0; t < trials; ++t)
      for (uint32_t i = 0; i < 6; ++i) {
        std::uniform_int_distribution<uint32_t> distribution(i, last);
        (void)distribution(prng);
      }
  }
};

float EventMs(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

}  // namespace

// Chooses the segment count so that the LIVE part of the scoring grid (model blocks below K;
// the blocks above exit at once) is just under a whole number of waves of the resident CTAs
// (148 SMs x kScoreCtasPerSm): K is only known on the device, so the live block count is
// estimated from the expected models per trial (3.9 of 8 for P6L, or what the call has seen so
// far), and the grid is kept 4 % under the wave boundary — a grid of 4.01 waves costs 5.
static void ChooseSegments(const ppsfm_ctx* ctx, int n, int kcap, double models_per_trial,
                           int* num_segs, int* seg_len, int shard_world = 1) {
  const int resident = ctx->num_sms * ppsfm::kScoreCtasPerSm;
  // (a sharded call scores 1 / shard_world of the wave's models on this GPU)
  const double live_models = std::max(1.0, (kcap / 8.0) * models_per_trial / shard_world);
  const int live_blocks = (int)std::ceil(live_models / ppsfm::kScoreModelsPerCta);
  const int forced = ppsfm::tune_int("PPSFM_SCORE_SEGS", 0);
  int best_segs = 1, best_len = std::max(256, ((n + 127) / 128) * 128);
  if (forced > 0) {
    best_len = std::max(256, (((n + forced - 1) / forced + 127) / 128) * 128);
    best_segs = (n + best_len - 1) / best_len;
  } else {
    // the largest grid of at most 4 waves (segments no shorter than 512 correspondences)
    const int budget = (int)(resident * 4 * 0.96);
    for (int segs = std::max(1, budget / live_blocks); segs >= 1; --segs) {
      const int len = std::max(512, (((n + segs - 1) / segs + 127) / 128) * 128);
      const int real = (n + len - 1) / len;
      if ((long long)real * live_blocks <= budget || segs == 1) {
        best_segs = real;
        best_len = len;
        break;
      }
    }
  }
  *num_segs = std::max(1, best_segs);
  *seg_len = best_len;
}

// One wave of the trial loop: trials [t_begin, t_end), its buffers (a slot of ctx->wave) and the
// generator state before its samples were drawn (needed to rewind after an abort).
struct Wave {
  size_t t_begin = 0, t_end = 0;
  int H = 0, kcap = 0, slot = 0;
  int n_first = 0;  // > 0: scored in two phases with exact pruning (launch_score)
  std::mt19937 prng_at_start;
};

// Runs trials [0, ...) of the RANSAC loop on a resident correspondence set.
//
// The loop is a software pipeline over waves of trials.  issue(w) draws the wave's samples on the
// host and queues copy + solve on the high-priority stream and the scoring kernel on the main
// stream; consume(w) waits for the wave's counts and replays the reference's sequential loop
// over them.  Up to kWaveSlots waves are in flight, so the host work of the next waves (sampling,
// replay) and their latency-bound solve kernels run under the scoring kernel of the current one.
// Waves are issued ahead only over trials the loop is certain to reach (below min_num_trials, or
// below the current dynamic bound once a best model exists), so a call that stops inside its
// first wave pays nothing for the pipeline.
//
// sharded = true (RANSAC row): the call is COLLECTIVE over the context's
// communicator.  Every rank holds the whole correspondence set and the same generator state,
// draws the same samples and solves every hypothesis of a wave (the solve kernel is latency
// bound: its time does not depend on the number of hypotheses, and having all models on every
// rank means no model ever has to travel), but scores only its share of the wave's models (model
// blocks interleaved over the ranks); one NCCL sum all-reduce of the 32-bit counts per wave gives
// every rank every count, the pruning bound is raised to the best count over all ranks, and every
// rank replays the same loop over the same counts -> the same report, mask and generator state
// as the single-GPU call, on every rank.
int RansacResident(ppsfm_ctx* ctx, const ppsfm_corr* corr, const ppsfm_ransac_options* opt_in,
                   ppsfm_ransac_report* report, uint8_t* inlier_mask, bool sharded = false) {
  if (!ctx || !corr || !opt_in || !report) return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  ppsfm::ScoreShard shard;
  if (sharded && ctx->world > 1) {
    shard.world = ctx->world;
    shard.rank = ctx->rank;
  }
  ppsfm_ransac_options opt = *opt_in;
  // RANSACOptions::Check, src/optim/ransac.h:68-75
  if (!(opt.max_error > 0) || opt.min_inlier_ratio < 0 || opt.min_inlier_ratio > 1 ||
      opt.confidence < 0 || opt.confidence > 1 || opt.min_num_trials > opt.max_num_trials)
    return fail(ctx, PPSFM_ERR_INVALID, "RANSACOptions::Check failed");
  // constructor cap, src/optim/ransac.h:149-155
  {
    const size_t kNumSamples = 100000;
    const size_t dyn = ComputeNumTrials(static_cast<size_t>(opt.min_inlier_ratio * kNumSamples),
                                        kNumSamples, opt.confidence,
                                        opt.dyn_num_trials_multiplier);
    opt.max_num_trials = std::min<uint64_t>(opt.max_num_trials, dyn);
  }
  std::memset(report, 0, sizeof(*report));
  report->best_trial = -1;
  report->best_model_idx = -1;
  report->residual_sum = std::numeric_limits<double>::max();
  ctx->timing = ppsfm_ransac_timing{};
  const size_t n = corr->n;
  if (n < 6) return PPSFM_OK;  // src/optim/ransac.h:189-191
  if (n > 0x7fffffffull) return fail(ctx, PPSFM_ERR_INVALID, "too many correspondences");

  const double max_residual = opt.max_error * opt.max_error;
  const size_t max_num_trials = opt.max_num_trials;
  ppsfm::ReplayState rs;        // best support so far, dynamic trial bound (ransac_replay.h)
  rs.dyn_max_num_trials = max_num_trials;
  double best_model[12] = {0};
  bool finished = false;
  size_t reported_trials = max_num_trials;
  uint64_t scored = 0;
  float total_ms = 0.f;

  HostSampler sampler;
  sampler.Initialize(n);
  cudaStream_t st = ctx->stream;      // scoring kernels, result copies
  cudaStream_t hi = ctx->stream_hi;   // exact (tie / final) kernels; the waves' sample copies and
                                      // solve kernels go to their slot's own high-priority stream
  // the correspondence set may still be in flight on the main stream (upload + pack)
  PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev_sync, st));
  PPSFM_CUDA(ctx, cudaStreamWaitEvent(hi, ctx->ev_sync, 0));
  for (auto& sl : ctx->wave) PPSFM_CUDA(ctx, cudaStreamWaitEvent(sl.solve_stream, ctx->ev_sync, 0));

  const bool trace = ppsfm::tune_int("PPSFM_RANSAC_TRACE", 0) != 0;
  const auto host_t0 = std::chrono::steady_clock::now();
  auto host_ms = [&]() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0)
        .count();
  };
  if (trace) PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));  // GPU time base
  // exact pruning (launch_score): best count of the waves scored so far, kept on the device
  const bool kPrune = ppsfm::tune_int("PPSFM_RANSAC_PRUNE", 1) != 0;
  // smallest second phase worth its three extra launches (tests lower it to reach the path)
  const size_t kPruneMin = (size_t)std::max(128, ppsfm::tune_int("PPSFM_RANSAC_PRUNE_MIN", 2048));
  PPSFM_CUDA(ctx, ctx->d_best_lb.reserve(sizeof(unsigned)));
  PPSFM_CUDA(ctx, cudaMemsetAsync(ctx->d_best_lb.p, 0, sizeof(unsigned), st));
  // On every exit path the streams are drained: speculative waves may still be running.
  struct Drain {
    ppsfm_ctx* c;
    ~Drain() {
      for (auto& sl : c->wave) cudaStreamSynchronize(sl.solve_stream);
      cudaStreamSynchronize(c->stream);
      cudaStreamSynchronize(c->stream_hi);
      cudaStreamSynchronize(c->stream_copy);
    }
  } drain{ctx};

  // ---- wave schedule.  A "plan" is the range the reference loop is currently bound to reach
  // (at least min_num_trials and a floor that fills the GPU, at most the dynamic bound); it is
  // cut into waves (next_wave_range) so that the pipeline has something to overlap.
  const size_t kWaveFloor = 1024;
  const size_t kWaveCap = 1u << 17;
  // (a call sharded over >= 4 GPUs keeps every plan whole: its scoring is too short to hide the
  // next wave's solve, so a second wave would only add a solve latency and a collective)
  const size_t kChunks =
      (size_t)std::max(1, ppsfm::tune_int("PPSFM_RANSAC_CHUNKS", shard.world >= 4 ? 1 : 4));
  const size_t kFirst = (size_t)std::max(1, ppsfm::tune_int("PPSFM_RANSAC_FIRST", 3));
  const size_t kGrowth = (size_t)std::max(1, ppsfm::tune_int("PPSFM_RANSAC_GROWTH", 100));
  // Hypotheses per warp of the solve kernel: as few as keep all warps resident at once (the
  // kernel needs 254 registers: 8 warps per SM; 6 leaves room beside a scoring CTA).  Measured on
  // the bench workload: 8 lanes 0.28 ms per wave, 32 lanes 0.31 ms.
  const int kSolveLanes = ppsfm::tune_int("PPSFM_SOLVE_LANES", 0);
  // CTA size of the solve kernels that run under the previous wave's scoring: 256 threads x 255
  // registers fill an SM, so the solve sits on ~26 SMs and leaves the others to the scoring
  // kernel alone; spread out in 64-thread CTAs it shared ~100 SMs with it and cost the scoring
  // more than it gained (bench step 1.61 ms against 1.70 ms)
  const int kSolveThreadsLate = ppsfm::tune_int("PPSFM_SOLVE_THREADS_LATE", 256);
  size_t num_issued = 0;   // waves issued so far
  double models_per_trial = 3.9;  // P6L: 3.8 +- 0.1 on generic data; updated from consumed waves
  // the first wave of a call is solved eight lanes per hypothesis (p6l_octet.cuh): its latency is
  // the one thing the pipeline cannot hide; PPSFM_SOLVE_OCTET = 0 never, 2 every wave
  const int kSolveOctetMode = ppsfm::tune_int("PPSFM_SOLVE_OCTET", 1);
  auto solve_lanes = [&](int H) {
    if (kSolveLanes > 0) return kSolveLanes;
    if (kSolveOctetMode >= 2 || (kSolveOctetMode == 1 && num_issued == 0)) return ppsfm::kSolveOctet;
    // later waves solve under the scoring of the previous one, where their latency is hidden
    // anyway: full warps, so that the solve occupies as few SMs as possible (kSolveThreadsLate)
    if (num_issued > 0) return 32;
    int lanes = 8;
    while (lanes < 32 && (H + lanes - 1) / lanes > ctx->num_sms * 6) lanes *= 2;
    return lanes;
  };
  size_t t_issue = 0;      // first trial not yet issued
  size_t plan_end = 0, plan_chunk = 0;
  // A plan is cut into a first wave of a third of its trials and a second wave with the rest: the
  // first wave's sampling and solve latency is the one thing nothing can hide, and its scoring has
  // to cover the solve latency of the second.  Measured on the bench workload (scripts/
  // plan_sweep.sh; first fraction 1/N, growth factor of the following waves): 1/3 + rest 1.74 ms
  // per call, 1/8 x2 1.79, 1/10 x3 1.81, 1/5 x4 1.91, whole plan 1.95.  PPSFM_RANSAC_CHUNKS=1 keeps
  // the plan whole; PPSFM_RANSAC_FIRST / PPSFM_RANSAC_GROWTH change the shape.
  auto next_wave_range = [&](size_t* t_end_out) {
    if (t_issue >= plan_end) {
      size_t want_end = std::max<size_t>(opt.min_num_trials, t_issue + kWaveFloor);
      if (rs.dyn_max_num_trials != std::numeric_limits<size_t>::max())
        want_end = std::max(want_end, std::min(rs.dyn_max_num_trials + 1, max_num_trials));
      plan_end = std::min(max_num_trials, want_end);
      const size_t span = plan_end - t_issue;
      plan_chunk = kChunks <= 1 ? kWaveCap
                                : std::min(kWaveCap, std::max(kWaveFloor, (span / kFirst + 255) / 256 * 256));
    }
    size_t t_end = std::min(plan_end, t_issue + plan_chunk);
    if (plan_end - t_end < plan_chunk) t_end = std::min(plan_end, t_issue + kWaveCap);
    plan_chunk = std::min(kWaveCap, plan_chunk * kGrowth);
    *t_end_out = t_end;
  };

  auto issue = [&](Wave& w) -> int {
    auto& sl = ctx->wave[w.slot];
    cudaStream_t hs_stream = sl.solve_stream;
    const int H = w.H, kcap = w.kcap;
    w.prng_at_start = ctx->prng;
    const double trace_t0 = trace ? host_ms() : 0.0;
    // ---- sample on the host (A7)
    PPSFM_CUDA(ctx, sl.h_samples.reserve(sizeof(uint32_t) * 6 * (size_t)H));
    uint32_t* hs = sl.h_samples.as<uint32_t>();
    for (int t = 0; t < H; ++t) sampler.Sample(ctx->prng, hs + 6 * (size_t)t);
    // ---- device buffers
    PPSFM_CUDA(ctx, sl.d_samples.reserve(sizeof(uint32_t) * 6 * (size_t)H));
    PPSFM_CUDA(ctx, sl.d_models.reserve(sizeof(double) * 96 * (size_t)H));
    PPSFM_CUDA(ctx, sl.d_num_models.reserve(sizeof(int) * (size_t)H));
    PPSFM_CUDA(ctx, sl.d_off.reserve(sizeof(int) * ((size_t)H + 1)));
    // Two-phase scoring with exact pruning for every wave after the first of the call: all
    // models on the first n_first correspondences, then only those that can still reach the best
    // count of the earlier waves.  A model with inlier ratio rho is dropped if
    // rho n_first + (n - n_first) < best; with n_first = n (1 - 0.8 r), r the inlier ratio the
    // call expects of its best model (min_inlier_ratio, or the best ratio seen so far), that is
    // every model with rho < r / 5 or so — the bulk of the hypotheses.
    ppsfm::ScorePrune prune;
    int n_first = (int)n;
    if (kPrune) {
      prune.best_lb = ctx->d_best_lb.as<unsigned>();
      double r = opt.min_inlier_ratio;
      if (rs.have_best) r = std::max(r, (double)rs.best_inliers / (double)n);
      const size_t cut = ((size_t)((double)n * (1.0 - 0.8 * std::min(1.0, r))) + 127) / 128 * 128;
      if (num_issued > 0 && r > 0.0 && cut >= 2 * kPruneMin && cut + kPruneMin <= n) n_first = (int)cut;
    }
    int num_segs, seg_len;
    ChooseSegments(ctx, n_first, kcap, models_per_trial, &num_segs, &seg_len, shard.world);
    int part_segs = num_segs;
    if (n_first < (int)n) {
      prune.n_first = n_first;
      // (few models survive: the second phase is sized as if one in eight did)
      ChooseSegments(ctx, (int)n - n_first, kcap, 1.0, &prune.num_segs2, &prune.seg_len2,
                     shard.world);
      part_segs = std::max(part_segs, prune.num_segs2);
      PPSFM_CUDA(ctx, sl.d_list.reserve(sizeof(int) * ((size_t)kcap + 1)));
      prune.list = sl.d_list.as<int>() + 1;
      prune.list_count = sl.d_list.as<int>();
    }
    PPSFM_CUDA(ctx, sl.d_part_cnt.reserve(sizeof(unsigned) * (size_t)part_segs * kcap));
    PPSFM_CUDA(ctx, sl.d_cnt.reserve(sizeof(unsigned) * (size_t)kcap));
    PPSFM_CUDA(ctx, sl.h_off.reserve(sizeof(int) * ((size_t)H + 2)));  // + survivor count
    PPSFM_CUDA(ctx, sl.h_cnt.reserve(sizeof(unsigned) * (size_t)kcap));
    w.n_first = prune.n_first;
    // ---- copy + solve on the slot's high-priority stream
    PPSFM_CUDA(ctx, cudaMemcpyAsync(sl.d_samples.p, hs, sizeof(uint32_t) * 6 * (size_t)H,
                                    cudaMemcpyHostToDevice, hs_stream));
    PPSFM_CUDA(ctx, cudaEventRecord(sl.ev[0], hs_stream));
    launch_p6l_solve(corr->corr6, corr->aligned, sl.d_samples.as<uint32_t>(), H,
                     sl.d_models.as<double>(), sl.d_num_models.as<int>(), hs_stream,
                     solve_lanes(H), num_issued > 0 ? kSolveThreadsLate : 64);
    launch_model_offsets(sl.d_num_models.as<int>(), H, sl.d_off.as<int>(), hs_stream);
    PPSFM_CUDA(ctx, cudaEventRecord(sl.ev[1], hs_stream));
    // ---- score on the main stream, results to the host
    PPSFM_CUDA(ctx, cudaStreamWaitEvent(st, sl.ev[1], 0));
    PPSFM_CUDA(ctx, cudaEventRecord(sl.ev[2], st));
    launch_score(corr->corr6, corr->corr6f, corr->bounds, (int)n, sl.d_models.as<double>(),
                 sl.d_off.as<int>(), H, num_segs, seg_len, max_residual, kcap,
                 sl.d_part_cnt.as<unsigned>(), sl.d_cnt.as<unsigned>(), st, prune, shard);
    PPSFM_CUDA(ctx, cudaEventRecord(sl.ev[3], st));
    if (shard.world > 1) {
      // the wave's one exchange: every rank gets every count; then the pruning bound of the
      // following waves becomes the best count over all ranks
      const int rc = ppsfm::CommAllReduceU32(ctx, sl.d_cnt.as<unsigned>(), (size_t)kcap, st);
      if (rc != PPSFM_OK) return rc;
      if (kPrune)
        launch_raise_best_lb(sl.d_cnt.as<unsigned>(), sl.d_off.as<int>(), H,
                             ctx->d_best_lb.as<unsigned>(), st);
      PPSFM_CUDA(ctx, cudaEventRecord(sl.ev[5], st));
      ctx->timing.kernel_launches += 2;
    }
    // results to the host on the copy stream: the next wave's scoring kernel follows directly
    cudaStream_t cp = ctx->stream_copy;
    PPSFM_CUDA(ctx, cudaStreamWaitEvent(cp, shard.world > 1 ? sl.ev[5] : sl.ev[3], 0));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(sl.h_off.p, sl.d_off.p, sizeof(int) * ((size_t)H + 1),
                                    cudaMemcpyDeviceToHost, cp));
    if (prune.n_first > 0)
      PPSFM_CUDA(ctx, cudaMemcpyAsync(sl.h_off.as<int>() + H + 1, prune.list_count, sizeof(int),
                                      cudaMemcpyDeviceToHost, cp));
    // (K is not known on the host yet: all kcap counts travel, 32 B per trial)
    PPSFM_CUDA(ctx, cudaMemcpyAsync(sl.h_cnt.p, sl.d_cnt.p, sizeof(unsigned) * (size_t)kcap,
                                    cudaMemcpyDeviceToHost, cp));
    PPSFM_CUDA(ctx, cudaEventRecord(sl.ev[4], cp));
    ctx->timing.kernel_launches += prune.n_first > 0 ? 7 : 4;
    ctx->timing.score_launches += prune.n_first > 0 ? 2 : 1;
    if (trace)
      fprintf(stderr, "[ransac] issue   trials %zu..%zu host %.3f -> %.3f ms\n", w.t_begin, w.t_end,
              trace_t0, host_ms());
    return PPSFM_OK;
  };

  // ---- index-order support (+ mask) of the best model, src/optim/ransac.h:251-275.  Launched as
  // soon as a wave has changed the best model, on the high-priority stream, so that it runs under
  // the scoring of the following waves; if the best model is still the same at the end of the loop
  // the result is simply picked up.
  const bool mask_wanted = inlier_mask != nullptr;
  bool final_valid = false;      // a launch for the CURRENT best model is in flight / done
  bool final_has_mask = false;
  auto launch_final = [&]() -> int {
    PPSFM_CUDA(ctx, ctx->d_fmodel.reserve(sizeof(double) * 12));
    PPSFM_CUDA(ctx, ctx->d_frbuf.reserve(sizeof(double) * n));
    PPSFM_CUDA(ctx, ctx->d_fmask.reserve(n));
    PPSFM_CUDA(ctx, ctx->h_fmask.reserve(n));
    PPSFM_CUDA(ctx, ctx->d_fcnt.reserve(sizeof(unsigned long long)));
    PPSFM_CUDA(ctx, ctx->d_fsum.reserve(sizeof(double)));
    PPSFM_CUDA(ctx, ctx->h_fres.reserve(16));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_fmodel.p, best_model, sizeof(best_model),
                                    cudaMemcpyHostToDevice, hi));
    PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev_final[0], hi));
    final_has_mask = mask_wanted && rs.best_inliers >= 6;
    launch_exact(corr->corr6, (int)n, ctx->d_fmodel.as<double>(), 1, max_residual,
                 ctx->d_frbuf.as<double>(), final_has_mask ? ctx->d_fmask.as<uint8_t>() : nullptr,
                 ctx->d_fcnt.as<unsigned long long>(), ctx->d_fsum.as<double>(), hi);
    PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev_final[1], hi));
    ctx->timing.kernel_launches += 2;
    if (final_has_mask)
      PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_fmask.p, ctx->d_fmask.p, n, cudaMemcpyDeviceToHost, hi));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_fres.p, ctx->d_fsum.p, sizeof(double),
                                    cudaMemcpyDeviceToHost, hi));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_fres.as<char>() + 8, ctx->d_fcnt.p,
                                    sizeof(unsigned long long), cudaMemcpyDeviceToHost, hi));
    PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev_final[2], hi));
    final_valid = true;
    return PPSFM_OK;
  };

  auto consume = [&](Wave& w) -> int {
    auto& sl = ctx->wave[w.slot];
    const int H = w.H;
    const size_t t_begin = w.t_begin, t_end = w.t_end;
    const double trace_t0 = trace ? host_ms() : 0.0;
    PPSFM_CUDA(ctx, cudaEventSynchronize(sl.ev[4]));
    if (trace)
      fprintf(stderr,
              "[ransac] consume trials %zu..%zu host wait %.3f -> %.3f ms | gpu solve %.3f..%.3f "
              "score %.3f..%.3f results %.3f\n",
              w.t_begin, w.t_end, trace_t0, host_ms(), EventMs(ctx->ev[0], sl.ev[0]),
              EventMs(ctx->ev[0], sl.ev[1]), EventMs(ctx->ev[0], sl.ev[2]),
              EventMs(ctx->ev[0], sl.ev[3]), EventMs(ctx->ev[0], sl.ev[4]));
    int* h_off = sl.h_off.as<int>();
    unsigned* h_cnt = sl.h_cnt.as<unsigned>();
    const int K = h_off[H];
    if (H >= 256) models_per_trial = std::max(0.25, (double)K / H) * 1.02;
    ctx->timing.solve_ms += EventMs(sl.ev[0], sl.ev[1]);
    ctx->timing.score_ms += EventMs(sl.ev[2], sl.ev[3]);
    total_ms += EventMs(sl.ev[0], sl.ev[1]) + EventMs(sl.ev[2], sl.ev[3]);
    if (shard.world > 1) {
      ctx->timing.comm_ms += EventMs(sl.ev[3], sl.ev[5]);
      total_ms += EventMs(sl.ev[3], sl.ev[5]);
    }
    // pairs the scoring kernel actually evaluated (dropped models skip the second phase)
    const uint64_t K_own = ppsfm_ransac_shard_models((uint64_t)K, shard.rank, shard.world);
    ctx->timing.score_pairs += w.n_first > 0 ? K_own * w.n_first +
                                                   (uint64_t)h_off[H + 1] * (n - w.n_first)
                                             : K_own * n;

    // ---- pass 1 (counts only): models that beat or tie the running best count.  Only a TIE
    // needs residual sums (InlierSupportMeasurer::Compare, support_measurement.cc:52-60), and
    // those must be index-order sums to match the reference bit for bit.
    std::vector<int> cand;
    const bool has_tie = ppsfm::replay_candidates(h_cnt, K, rs, &cand);
    auto model_src = [&](int k) -> size_t {
      const int t = int(std::upper_bound(h_off, h_off + H + 1, k) - h_off) - 1;
      return (size_t)t * 96 + (size_t)(k - h_off[t]) * 12;
    };
    const int E = (int)cand.size();
    std::vector<double> cand_sum;
    if (has_tie) {
      // exact (index-order) supports for every candidate of this wave (+ the carried best)
      const bool carry = rs.have_best && !rs.best_sum_known;
      const int EE = E + (carry ? 1 : 0);
      cand_sum.resize(EE);
      PPSFM_CUDA(ctx, ctx->h_esum.reserve(sizeof(double) * (size_t)EE));
      PPSFM_CUDA(ctx, ctx->h_ecnt.reserve(sizeof(unsigned long long) * (size_t)EE));
      // Batches of up to 64 MB of residuals.  With a high inlier ratio (the mapper's registration
      // calls) most good models TIE at the full inlier count, so a wave can hold thousands of
      // candidates: they are gathered by one kernel from an index list, not copied one by one.
      const int kBatch = (int)std::min<size_t>(8192, std::max<size_t>(32, (64u << 20) / (8 * n)));
      const int cap = std::min(kBatch, EE);
      PPSFM_CUDA(ctx, ctx->d_emodels.reserve(sizeof(double) * 12 * cap));
      PPSFM_CUDA(ctx, ctx->d_rbuf.reserve(sizeof(double) * (size_t)cap * n));
      PPSFM_CUDA(ctx, ctx->d_ecnt.reserve(sizeof(unsigned long long) * cap));
      PPSFM_CUDA(ctx, ctx->d_esum.reserve(sizeof(double) * cap));
      PPSFM_CUDA(ctx, ctx->d_eidx.reserve(sizeof(long long) * cap));
      PPSFM_CUDA(ctx, ctx->h_eidx.reserve(sizeof(long long) * (size_t)EE));
      long long* h_idx = ctx->h_eidx.as<long long>();
      for (int e = 0; e < E; ++e) h_idx[e] = (long long)model_src(cand[e]);
      for (int e0 = 0; e0 < EE; e0 += kBatch) {
        const int ne = std::min(kBatch, EE - e0);
        const int ng = std::min(ne, E - e0);  // gathered from the wave; the last one may be the carried best
        if (ng > 0) {
          PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_eidx.p, h_idx + e0, sizeof(long long) * ng,
                                          cudaMemcpyHostToDevice, hi));
          launch_gather_models(sl.d_models.as<double>(), ctx->d_eidx.as<long long>(), ng,
                               ctx->d_emodels.as<double>(), hi);
          ++ctx->timing.kernel_launches;
        }
        if (ne > std::max(ng, 0))
          PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_emodels.as<double>() + (size_t)std::max(ng, 0) * 12,
                                          best_model, sizeof(double) * 12, cudaMemcpyHostToDevice,
                                          hi));
        PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[3], hi));
        launch_exact(corr->corr6, (int)n, ctx->d_emodels.as<double>(), ne, max_residual,
                     ctx->d_rbuf.as<double>(), nullptr, ctx->d_ecnt.as<unsigned long long>(),
                     ctx->d_esum.as<double>(), hi);
        PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[4], hi));
        ctx->timing.kernel_launches += 2;
        PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_esum.as<double>() + e0, ctx->d_esum.p,
                                        sizeof(double) * ne, cudaMemcpyDeviceToHost, hi));
        PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_ecnt.as<unsigned long long>() + e0, ctx->d_ecnt.p,
                                        sizeof(unsigned long long) * ne, cudaMemcpyDeviceToHost,
                                        hi));
        PPSFM_CUDA(ctx, cudaStreamSynchronize(hi));
        const float ems = EventMs(ctx->ev[3], ctx->ev[4]);
        ctx->timing.exact_ms += ems;
        total_ms += ems;
      }
      for (int e = 0; e < EE; ++e) {
        cand_sum[e] = ctx->h_esum.as<double>()[e];
        const size_t want = e < E ? (size_t)h_cnt[cand[e]] : rs.best_inliers;
        if (ctx->h_ecnt.as<unsigned long long>()[e] != want)
          return fail(ctx, PPSFM_ERR_CUDA, "internal: exact/segmented inlier counts differ");
      }
      if (carry) {
        rs.best_sum = cand_sum[E];
        rs.best_sum_known = true;
      }
    }

    // ---- pass 2: replay of src/optim/ransac.h:213-249 over this wave (ransac_replay.h).
    // An abort at trial t means: samples were drawn for trials 0..t, and the loop reports
    // num_trials = t + 2 (the `if (abort) { num_trials += 1; break; }` at the top of the next
    // iteration) unless t + 1 already equals max_num_trials.
    ppsfm::ReplayParams rp;
    rp.t_begin = t_begin;
    rp.t_end = t_end;
    rp.num_samples = n;
    rp.min_num_trials = opt.min_num_trials;
    rp.confidence = opt.confidence;
    rp.multiplier = opt.dyn_num_trials_multiplier;
    const ppsfm::ReplayOutcome ro =
        ppsfm::replay_wave(h_off, H, h_cnt, cand, has_tie ? cand_sum.data() : nullptr, rp, &rs);
    scored += ro.scored;
    const int best_k = ro.best_k;
    if (best_k >= 0) {
      report->best_trial = rs.best_trial;
      report->best_model_idx = rs.best_model_idx;
    }
    if (ro.abort_model >= 0) {
      const size_t t_abort = ro.t_abort;
      reported_trials = (t_abort + 1 < max_num_trials) ? t_abort + 2 : max_num_trials;
      // the reference drew samples for trials 0..t_abort only: rewind the generator to the
      // start of this wave (later waves may have been sampled ahead) and skip forward
      ctx->prng = w.prng_at_start;
      HostSampler::Skip(ctx->prng, n, t_abort + 1 - t_begin);
      finished = true;
    }
    if (best_k >= 0) {  // the best model changed in this wave: bring its 12 doubles to the host
      PPSFM_CUDA(ctx, cudaMemcpyAsync(best_model, sl.d_models.as<double>() + model_src(best_k),
                                      sizeof(best_model), cudaMemcpyDeviceToHost, hi));
      PPSFM_CUDA(ctx, cudaStreamSynchronize(hi));
      const int rc = launch_final();
      if (rc != PPSFM_OK) return rc;
    }
    return PPSFM_OK;
  };

  // ---- the pipeline
  constexpr int kSlots = ppsfm_ctx::kWaveSlots;
  Wave waves[kSlots];
  int head = 0, in_flight = 0;  // waves[head] is the oldest wave in flight
  auto certain_to_reach = [&](size_t t) {
    return t < opt.min_num_trials || (rs.have_best && t <= rs.dyn_max_num_trials);
  };
  while (!finished) {
    // issue: always when nothing is in flight, ahead only over trials certain to be reached
    while (in_flight < kSlots && t_issue < max_num_trials &&
           (in_flight == 0 || certain_to_reach(t_issue))) {
      Wave& w = waves[(head + in_flight) % kSlots];
      w.slot = (head + in_flight) % kSlots;
      w.t_begin = t_issue;
      next_wave_range(&w.t_end);
      w.H = static_cast<int>(w.t_end - w.t_begin);
      w.kcap = 8 * w.H;
      const int rc = issue(w);
      if (rc != PPSFM_OK) return rc;
      t_issue = w.t_end;
      ++in_flight;
      ++num_issued;
    }
    if (in_flight == 0) break;  // all trials done
    const int rc = consume(waves[head]);
    if (rc != PPSFM_OK) return rc;
    head = (head + 1) % kSlots;
    --in_flight;
  }

  report->num_trials = reported_trials;
  report->num_inliers = rs.best_inliers;
  report->num_models_scored = scored;
  std::memcpy(report->model, best_model, sizeof(best_model));

  // Support + inlier mask of the best model in reference (index) order
  // (src/optim/ransac.h:251-275: the reference also rescans the best model once more).
  if (rs.have_best) {
    if (!final_valid) {
      const int rc = launch_final();
      if (rc != PPSFM_OK) return rc;
    }
    const double trace_t0 = trace ? host_ms() : 0.0;
    PPSFM_CUDA(ctx, cudaEventSynchronize(ctx->ev_final[2]));
    if (trace)
      fprintf(stderr, "[ransac] final   host wait %.3f -> %.3f ms | gpu exact %.3f..%.3f\n",
              trace_t0, host_ms(), EventMs(ctx->ev[0], ctx->ev_final[0]),
              EventMs(ctx->ev[0], ctx->ev_final[1]));
    unsigned long long fcnt;
    std::memcpy(&fcnt, ctx->h_fres.as<char>() + 8, sizeof(fcnt));
    if (fcnt != rs.best_inliers)
      return fail(ctx, PPSFM_ERR_CUDA, "internal: exact/segmented inlier counts differ");
    std::memcpy(&rs.best_sum, ctx->h_fres.p, sizeof(double));
    if (final_has_mask) std::memcpy(inlier_mask, ctx->h_fmask.p, n);
    const float ems = EventMs(ctx->ev_final[0], ctx->ev_final[1]);
    ctx->timing.exact_ms += ems;
    total_ms += ems;
  }
  report->residual_sum = rs.best_sum;
  ctx->timing.total_ms = total_ms;
  PPSFM_CUDA(ctx, cudaGetLastError());
  if (rs.best_inliers < 6) return PPSFM_OK;  // src/optim/ransac.h:255-259
  report->success = 1;
  return PPSFM_OK;
}

// Copies a host correspondence set to HBM and packs it into 48-byte records.  With
// `use_ctx_buffers` the device storage is the context's growable scratch (no cudaMalloc /
// cudaFree on the per-call path); otherwise the set owns its allocation (resident handles).
int UploadCorr(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned, const double* points,
               size_t n, bool use_ctx_buffers, ppsfm_corr** out) {
  if (!ctx || !out || (n > 0 && (!lines || !points)))
    return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  ppsfm_corr* c = new ppsfm_corr();
  c->n = n;
  c->owns = !use_ctx_buffers;
  cudaStream_t st = ctx->stream;
  if (n > 0) {
    cudaError_t e;
    if (use_ctx_buffers) {
      e = ctx->d_corr6.reserve(sizeof(double) * 6 * n);
      if (e == cudaSuccess) e = ctx->d_aligned.reserve(n);
      if (e == cudaSuccess) e = ctx->d_bounds.reserve(4 * sizeof(double));
      if (e == cudaSuccess) e = ctx->d_corr6f.reserve(sizeof(float) * 12 * ((n + 1) / 2));
      c->corr6f = ctx->d_corr6f.as<float>();
      c->corr6 = ctx->d_corr6.as<double>();
      c->aligned = ctx->d_aligned.as<uint8_t>();
      c->bounds = ctx->d_bounds.as<double>();
    } else {
      e = cudaMalloc(&c->corr6, sizeof(double) * 6 * n);
      if (e == cudaSuccess) e = cudaMalloc(&c->aligned, n);
      if (e == cudaSuccess) e = cudaMalloc(&c->bounds, 4 * sizeof(double));
      if (e == cudaSuccess) e = cudaMalloc(&c->corr6f, sizeof(float) * 12 * ((n + 1) / 2));
    }
    if (e != cudaSuccess) {
      if (c->owns) {
        if (c->corr6) cudaFree(c->corr6);
        if (c->aligned) cudaFree(c->aligned);
        if (c->bounds) cudaFree(c->bounds);
      }
      delete c;
      return fail(ctx, PPSFM_ERR_CUDA, "device allocation: %s", cudaGetErrorString(e));
    }
    PPSFM_CUDA(ctx, ctx->d_tmp_corr.reserve(sizeof(double) * 6 * n));
    double* tl = ctx->d_tmp_corr.as<double>();
    double* tp = tl + 3 * n;
    PPSFM_CUDA(ctx, cudaMemcpyAsync(tl, lines, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(tp, points, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
    if (aligned) {
      PPSFM_CUDA(ctx, cudaMemcpyAsync(c->aligned, aligned, n, cudaMemcpyHostToDevice, st));
    } else {
      PPSFM_CUDA(ctx, cudaMemsetAsync(c->aligned, 0, n, st));
    }
    launch_pack_corr(tl, tp, n, c->corr6, c->corr6f, c->bounds, st);
    // no synchronisation here: later work is queued on the same stream; the host buffers must
    // stay valid until the call that consumes the set returns (all entry points are blocking)
  }
  *out = c;
  return PPSFM_OK;
}

void FreeCorr(ppsfm_corr* c) {
  if (!c) return;
  if (c->owns) {
    if (c->corr6) cudaFree(c->corr6);
    if (c->aligned) cudaFree(c->aligned);
    if (c->bounds) cudaFree(c->bounds);
    if (c->corr6f) cudaFree(c->corr6f);
  }
  delete c;
}

// Eigen::Quaterniond(Matrix3d) as used by RotationMatrixToQuaternion (src/base/pose.cc:41-44).
void RotationMatrixToQuaternion(const double* R /*col-major*/, double* q) {
  auto at = [&](int r, int c) { return R[3 * c + r]; };
  double t = at(0, 0) + at(1, 1) + at(2, 2);
  double w, v[3];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    w = 0.5 * t;
    t = 0.5 / t;
    v[0] = (at(2, 1) - at(1, 2)) * t;
    v[1] = (at(0, 2) - at(2, 0)) * t;
    v[2] = (at(1, 0) - at(0, 1)) * t;
  } else {
    int i = 0;
    if (at(1, 1) > at(0, 0)) i = 1;
    if (at(2, 2) > at(i, i)) i = 2;
    const int j = (i + 1) % 3;
    const int k = (j + 1) % 3;
    t = std::sqrt(at(i, i) - at(j, j) - at(k, k) + 1.0);
    v[i] = 0.5 * t;
    t = 0.5 / t;
    w = (at(k, j) - at(j, k)) * t;
    v[j] = (at(j, i) + at(i, j)) * t;
    v[k] = (at(k, i) + at(i, k)) * t;
  }
  q[0] = w;
  q[1] = v[0];
  q[2] = v[1];
  q[3] = v[2];
}

}  // namespace ppsfm

// ============================================================================================
// C-ABI
// ============================================================================================
using namespace ppsfm;

extern "C" {

const char* ppsfm_version(void) { return "ppsfm_b200 0.1 (sm_100a)"; }

int ppsfm_ctx_create(int device, ppsfm_ctx** out) {
  if (!out) return PPSFM_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return PPSFM_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return PPSFM_ERR_CUDA;
  ppsfm_ctx* ctx = new ppsfm_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return PPSFM_ERR_CUDA;
  }
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return PPSFM_ERR_CUDA;
  }
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  {
    int lo = 0, hi = 0;  // numerically lowest value = greatest priority
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&ctx->stream_hi, cudaStreamNonBlocking, hi) != cudaSuccess) {
      cudaStreamDestroy(ctx->stream);
      delete ctx;
      return PPSFM_ERR_CUDA;
    }
    cudaStreamCreateWithPriority(&ctx->stream_copy, cudaStreamNonBlocking, hi);
    cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming);
    for (auto& ev : ctx->ev_final) cudaEventCreate(&ev);
    for (auto& sl : ctx->wave) {
      for (auto& ev : sl.ev) cudaEventCreate(&ev);
      cudaStreamCreateWithPriority(&sl.solve_stream, cudaStreamNonBlocking, hi);
    }
  }
  {  // keep freed stream-ordered allocations cached in the device pool (BA scratch reuse)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long threshold = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
  }
  *out = ctx;
  return PPSFM_OK;
}

void ppsfm_comm_destroy(ppsfm_ctx* ctx);  // comm.cu

void ppsfm_ctx_destroy(ppsfm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->stream_hi) cudaStreamSynchronize(ctx->stream_hi);
  ppsfm_comm_destroy(ctx);
  ppsfm::DevBuf* dbufs[] = {&ctx->d_samples, &ctx->d_models, &ctx->d_num_models, &ctx->d_cmodels,
                            &ctx->d_msrc, &ctx->d_K, &ctx->d_part_cnt, &ctx->d_part_sum,
                            &ctx->d_cnt, &ctx->d_sum, &ctx->d_eidx, &ctx->d_emodels, &ctx->d_rbuf,
                            &ctx->d_esum, &ctx->d_ecnt, &ctx->d_mask, &ctx->d_tmp_corr,
                            &ctx->d_tmp_aligned, &ctx->d_corr6, &ctx->d_aligned, &ctx->d_bounds,
                            &ctx->d_corr6f};
  for (auto* b : dbufs) b->release();
  ppsfm::PinBuf* pbufs[] = {&ctx->h_samples, &ctx->h_num_models, &ctx->h_cnt, &ctx->h_sum,
                            &ctx->h_eidx, &ctx->h_emodels, &ctx->h_esum, &ctx->h_ecnt,
                            &ctx->h_mask, &ctx->h_K, &ctx->h_stage};
  for (auto* b : pbufs) b->release();
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& sl : ctx->wave) {
    ppsfm::DevBuf* d[] = {&sl.d_samples, &sl.d_models, &sl.d_num_models, &sl.d_off,
                          &sl.d_part_cnt, &sl.d_cnt, &sl.d_list};
    for (auto* b : d) b->release();
    ppsfm::PinBuf* h[] = {&sl.h_samples, &sl.h_off, &sl.h_cnt};
    for (auto* b : h) b->release();
    for (auto& ev : sl.ev)
      if (ev) cudaEventDestroy(ev);
    if (sl.solve_stream) cudaStreamDestroy(sl.solve_stream);
  }
  if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
  for (auto& ev : ctx->ev_final)
    if (ev) cudaEventDestroy(ev);
  {
    ppsfm::DevBuf* d[] = {&ctx->d_fmodel, &ctx->d_frbuf, &ctx->d_fmask, &ctx->d_fcnt, &ctx->d_fsum};
    for (auto* b : d) b->release();
    ctx->h_fmask.release();
    ctx->d_best_lb.release();
    ctx->h_fres.release();
  }
  if (ctx->stream_hi) cudaStreamDestroy(ctx->stream_hi);
  if (ctx->stream_copy) cudaStreamDestroy(ctx->stream_copy);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* ppsfm_last_error(const ppsfm_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

// Host-only self-test (no context, no GPU): the event-driven replay RansacResident uses against
// the literal model-by-model loop, on `rounds` random waves (empty trials, ties, carried best
// models, aborts at every position).  Returns the number of waves on which they differ.
int ppsfm_selftest_replay(uint32_t seed, int rounds) {
  std::mt19937 g(seed);
  auto uni = [&](int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(g); };
  int mismatches = 0;
  for (int r = 0; r < rounds; ++r) {
    const int H = uni(1, 300);
    const size_t N = (size_t)uni(50, 4000);
    std::vector<int> off(H + 1, 0);
    const int empty_pct = uni(0, 60);
    for (int t = 0; t < H; ++t) off[t + 1] = off[t] + (uni(0, 99) < empty_pct ? 0 : uni(1, 8));
    const int K = off[H];
    const int levels = uni(1, 40);  // few distinct counts -> ties
    const unsigned top = (unsigned)uni(1, (int)N);
    std::vector<unsigned> cnt(K);
    for (int k = 0; k < K; ++k) cnt[k] = (unsigned)((uint64_t)uni(0, levels) * top / levels);
    ppsfm::ReplayParams p;
    p.t_begin = uni(0, 1) ? 0 : (size_t)uni(1, 5000);
    p.t_end = p.t_begin + H;
    p.num_samples = N;
    p.min_num_trials = uni(0, 2) == 0 ? 0 : (size_t)uni(0, (int)p.t_end + 50);
    p.confidence = uni(0, 2) == 0 ? 0.9 : (uni(0, 1) ? 0.99 : 0.99999);
    p.multiplier = uni(0, 1) ? 3.0 : 1.0;
    ppsfm::ReplayState s0;
    s0.dyn_max_num_trials = (size_t)uni(1, 20000);
    if (uni(0, 1)) {  // a best model carried over from earlier waves
      s0.have_best = true;
      s0.best_inliers = (size_t)uni(0, (int)N);
      s0.best_sum = uni(0, 1000) * 1e-3;
      s0.best_sum_known = uni(0, 1) != 0;
      s0.dyn_max_num_trials = std::min<size_t>(
          s0.dyn_max_num_trials, ComputeNumTrials(s0.best_inliers, N, p.confidence, p.multiplier));
    }
    std::vector<int> cand;
    const bool has_tie = ppsfm::replay_candidates(cnt.data(), K, s0, &cand);
    std::vector<double> sums(cand.size());
    for (auto& v : sums) v = uni(0, 5) * 0.25;  // equal sums happen too
    if (has_tie && s0.have_best) s0.best_sum_known = true;  // the tie pass makes it exact
    const double* cs = has_tie ? sums.data() : nullptr;
    ppsfm::ReplayState a = s0, b = s0;
    const ppsfm::ReplayOutcome oa = ppsfm::replay_wave_literal(off.data(), H, cnt.data(), cand, cs, p, &a);
    const ppsfm::ReplayOutcome ob = ppsfm::replay_wave(off.data(), H, cnt.data(), cand, cs, p, &b);
    const bool same = oa.best_k == ob.best_k && oa.abort_model == ob.abort_model &&
                      (oa.abort_model < 0 || oa.t_abort == ob.t_abort) && oa.scored == ob.scored &&
                      a.have_best == b.have_best && a.best_inliers == b.best_inliers &&
                      a.best_sum == b.best_sum && a.best_sum_known == b.best_sum_known &&
                      a.dyn_max_num_trials == b.dyn_max_num_trials &&
                      a.best_trial == b.best_trial && a.best_model_idx == b.best_model_idx;
    if (!same) ++mismatches;
  }
  return mismatches;
}

void ppsfm_set_prng_seed(ppsfm_ctx* ctx, uint32_t seed) {
  if (ctx) ctx->prng = std::mt19937(seed);
}

uint32_t ppsfm_prng_peek(const ppsfm_ctx* ctx) {
  if (!ctx) return 0;
  std::mt19937 copy = ctx->prng;
  return static_cast<uint32_t>(copy());
}

void ppsfm_ransac_options_default(ppsfm_ransac_options* opt) {
  if (!opt) return;
  opt->max_error = 0.0;
  opt->min_inlier_ratio = 0.1;
  opt->confidence = 0.99;
  opt->dyn_num_trials_multiplier = 3.0;
  opt->min_num_trials = 0;
  opt->max_num_trials = std::numeric_limits<uint64_t>::max();
}

uint64_t ppsfm_compute_num_trials(uint64_t num_inliers, uint64_t num_samples, double confidence,
                                  double num_trials_multiplier) {
  return ComputeNumTrials(num_inliers, num_samples, confidence, num_trials_multiplier);
}

int ppsfm_sample_table(ppsfm_ctx* ctx, size_t n, size_t num_trials, uint32_t* table_out) {
  if (!ctx || !table_out || n < 6) return fail(ctx, PPSFM_ERR_INVALID, "bad sample_table args");
  HostSampler s;
  s.Initialize(n);
  for (size_t t = 0; t < num_trials; ++t) s.Sample(ctx->prng, table_out + 6 * t);
  return PPSFM_OK;
}

int ppsfm_corr_upload(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                      const double* points, size_t n, ppsfm_corr** out) {
  if (ctx) cudaSetDevice(ctx->device);
  int rc = UploadCorr(ctx, lines, aligned, points, n, false, out);
  if (rc == PPSFM_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    rc = fail(ctx, PPSFM_ERR_CUDA, "upload failed");
  return rc;
}

void ppsfm_corr_free(ppsfm_ctx* ctx, ppsfm_corr* corr) {
  if (ctx) cudaSetDevice(ctx->device);
  FreeCorr(corr);
}

int ppsfm_image_to_world_threshold(int camera_model, const double* camera_params, double threshold,
                                   double* out) {
  if (!camera_params || !out) return PPSFM_ERR_INVALID;
  // focal_length_idxs of src/base/camera_models.h:263-420
  double mean_focal_length = 0;
  switch (camera_model) {
    case 0: case 2: case 3: case 8: case 9:  // SIMPLE_PINHOLE, SIMPLE_RADIAL, RADIAL, *_FISHEYE: {0}
      mean_focal_length += camera_params[0];
      mean_focal_length /= 1;
      break;
    case 1: case 4: case 5: case 6: case 7: case 10:  // PINHOLE, OPENCV*, FOV, THIN_PRISM: {0, 1}
      mean_focal_length += camera_params[0];
      mean_focal_length += camera_params[1];
      mean_focal_length /= 2;
      break;
    default:
      return PPSFM_ERR_INVALID;
  }
  *out = threshold / mean_focal_length;
  return PPSFM_OK;
}

void ppsfm_rotation_matrix_to_quaternion(const double* R, double* qvec) {
  RotationMatrixToQuaternion(R, qvec);
}

int ppsfm_ransac_p6l_resident(ppsfm_ctx* ctx, const ppsfm_corr* corr,
                              const ppsfm_ransac_options* options, ppsfm_ransac_report* report,
                              uint8_t* inlier_mask) {
  if (ctx) cudaSetDevice(ctx->device);
  return RansacResident(ctx, corr, options, report, inlier_mask);
}

int ppsfm_ransac_p6l(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                     const double* points, size_t n, const ppsfm_ransac_options* options,
                     ppsfm_ransac_report* report, uint8_t* inlier_mask) {
  if (ctx) cudaSetDevice(ctx->device);
  ppsfm_corr* corr = nullptr;
  int rc = UploadCorr(ctx, lines, aligned, points, n, true, &corr);
  if (rc != PPSFM_OK) return rc;
  rc = RansacResident(ctx, corr, options, report, inlier_mask);
  if (ctx) cudaStreamSynchronize(ctx->stream);
  FreeCorr(corr);
  return rc;
}

// Host-only: how many of a wave's K compact models rank `rank` of `world` scores in a sharded
// call (blocks of kScoreModelsPerCta models, block b belongs to rank b % world).
uint64_t ppsfm_ransac_shard_models(uint64_t num_models, int rank, int world) {
  if (world <= 1) return num_models;
  uint64_t own = 0;
  const uint64_t B = (uint64_t)ppsfm::kScoreModelsPerCta;
  for (uint64_t b = (uint64_t)rank; b * B < num_models; b += (uint64_t)world)
    own += std::min<uint64_t>(B, num_models - b * B);
  return own;
}

int ppsfm_ransac_p6l_resident_sharded(ppsfm_ctx* ctx, const ppsfm_corr* corr,
                                      const ppsfm_ransac_options* options,
                                      ppsfm_ransac_report* report, uint8_t* inlier_mask) {
  if (ctx) cudaSetDevice(ctx->device);
  return RansacResident(ctx, corr, options, report, inlier_mask, true);
}

int ppsfm_ransac_p6l_sharded(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                             const double* points, size_t n, const ppsfm_ransac_options* options,
                             ppsfm_ransac_report* report, uint8_t* inlier_mask) {
  if (ctx) cudaSetDevice(ctx->device);
  ppsfm_corr* corr = nullptr;
  int rc = UploadCorr(ctx, lines, aligned, points, n, true, &corr);
  if (rc != PPSFM_OK) return rc;
  rc = RansacResident(ctx, corr, options, report, inlier_mask, true);
  if (ctx) cudaStreamSynchronize(ctx->stream);
  FreeCorr(corr);
  return rc;
}

int ppsfm_estimate_absolute_pose_from_lines(ppsfm_ctx* ctx, const double* lines,
                                            const uint8_t* aligned, const double* points,
                                            size_t n, const ppsfm_ransac_options* options,
                                            double* qvec, double* tvec, uint64_t* num_inliers,
                                            uint8_t* inlier_mask, ppsfm_ransac_report* report_out) {
  // src/estimators/pose.cc:52-94
  if (!ctx || !qvec || !tvec || !num_inliers)
    return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  std::vector<uint8_t> mask(n, 0);
  ppsfm_ransac_report report;
  int rc = ppsfm_ransac_p6l(ctx, lines, aligned, points, n, options, &report, mask.data());
  if (rc != PPSFM_OK) return rc;
  if (report_out) *report_out = report;
  *num_inliers = report.num_inliers;
  if (inlier_mask) std::memcpy(inlier_mask, mask.data(), n);
  if (*num_inliers == 0) return PPSFM_NO_SOLUTION;
  // reference: `inlier_mask->at(i)` on an empty mask (success == false, 1..5 inliers) would
  // throw; we report "no solution" instead.
  if (!report.success) return PPSFM_NO_SOLUTION;
  size_t num_aligned_inliers = 0;
  for (size_t i = 0; i < n; ++i)
    if (mask[i] && aligned && aligned[i]) num_aligned_inliers += 1;
  if (num_aligned_inliers > *num_inliers * 0.9) return PPSFM_NO_SOLUTION;
  RotationMatrixToQuaternion(report.model, qvec);
  tvec[0] = report.model[9];
  tvec[1] = report.model[10];
  tvec[2] = report.model[11];
  for (int i = 0; i < 4; ++i)
    if (std::isnan(qvec[i])) return PPSFM_NO_SOLUTION;
  for (int i = 0; i < 3; ++i)
    if (std::isnan(tvec[i])) return PPSFM_NO_SOLUTION;
  return PPSFM_OK;
}

int ppsfm_p6l_solve_batch(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                          const double* points, size_t n, const uint32_t* sample_idx,
                          size_t num_samples, double* models_out, int32_t* num_models_out) {
  if (!ctx || !sample_idx || !models_out || !num_models_out)
    return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  for (size_t i = 0; i < 6 * num_samples; ++i)
    if (sample_idx[i] >= n) return fail(ctx, PPSFM_ERR_INVALID, "sample index out of range");
  cudaSetDevice(ctx->device);
  ppsfm_corr* corr = nullptr;
  int rc = UploadCorr(ctx, lines, aligned, points, n, true, &corr);
  if (rc != PPSFM_OK) return rc;
  cudaStream_t st = ctx->stream;
  const size_t H = num_samples;
  auto body = [&]() -> int {
    PPSFM_CUDA(ctx, ctx->d_samples.reserve(sizeof(uint32_t) * 6 * H));
    PPSFM_CUDA(ctx, ctx->d_models.reserve(sizeof(double) * 96 * H));
    PPSFM_CUDA(ctx, ctx->d_num_models.reserve(sizeof(int) * H));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_samples.p, sample_idx, sizeof(uint32_t) * 6 * H,
                                    cudaMemcpyHostToDevice, st));
    PPSFM_CUDA(ctx, cudaMemsetAsync(ctx->d_models.p, 0, sizeof(double) * 96 * H, st));
    // (PPSFM_SOLVE_OCTET >= 2: through the eight-lanes-per-hypothesis kernel — the parity tests
    // run both and require identical bits)
    launch_p6l_solve(corr->corr6, corr->aligned, ctx->d_samples.as<uint32_t>(), (int)H,
                     ctx->d_models.as<double>(), ctx->d_num_models.as<int>(), st,
                     ppsfm::tune_int("PPSFM_SOLVE_OCTET", 1) >= 2 ? ppsfm::kSolveOctet : 32);
    PPSFM_CUDA(ctx, cudaMemcpyAsync(models_out, ctx->d_models.p, sizeof(double) * 96 * H,
                                    cudaMemcpyDeviceToHost, st));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(num_models_out, ctx->d_num_models.p, sizeof(int) * H,
                                    cudaMemcpyDeviceToHost, st));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(st));
    PPSFM_CUDA(ctx, cudaGetLastError());
    return PPSFM_OK;
  };
  rc = H > 0 ? body() : PPSFM_OK;
  FreeCorr(corr);
  return rc;
}

// Test hook: inlier counts of `num_models` models through the RANSAC scoring kernel (the filtered
// count-only path with its reference fallback), so that tests can aim at the filter's edge cases
// directly.  Same result as num_inliers_out of ppsfm_line_residuals, by construction.
int ppsfm_score_models(ppsfm_ctx* ctx, const double* lines, const double* points, size_t n,
                       const double* models, size_t num_models, double max_residual,
                       uint32_t* counts_out) {
  if (!ctx || !models || !counts_out) return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  if (n == 0 || num_models == 0) {
    for (size_t k = 0; k < num_models; ++k) counts_out[k] = 0;
    return PPSFM_OK;
  }
  if (n > 0x7fffffffull || num_models > (1u << 24)) return fail(ctx, PPSFM_ERR_INVALID, "too large");
  cudaSetDevice(ctx->device);
  ppsfm_corr* corr = nullptr;
  int rc = UploadCorr(ctx, lines, nullptr, points, n, true, &corr);
  if (rc != PPSFM_OK) return rc;
  cudaStream_t st = ctx->stream;
  const int H = (int)num_models, kcap = 8 * H;
  auto body = [&]() -> int {
    // one model per trial: slot 0 of every 8-model group, offsets = 0, 1, 2, ...
    std::vector<double> hm((size_t)H * 96, 0.0);
    std::vector<int> hoff(H + 1);
    for (int k = 0; k < H; ++k) {
      std::copy(models + 12 * (size_t)k, models + 12 * (size_t)k + 12, hm.begin() + 96 * (size_t)k);
      hoff[k] = k;
    }
    hoff[H] = H;
    int num_segs, seg_len;
    ChooseSegments(ctx, (int)n, kcap, 1.0, &num_segs, &seg_len);  // one model per trial here
    PPSFM_CUDA(ctx, ctx->d_models.reserve(sizeof(double) * hm.size()));
    PPSFM_CUDA(ctx, ctx->d_msrc.reserve(sizeof(int) * ((size_t)H + 1)));
    PPSFM_CUDA(ctx, ctx->d_part_cnt.reserve(sizeof(unsigned) * (size_t)num_segs * kcap));
    PPSFM_CUDA(ctx, ctx->d_cnt.reserve(sizeof(unsigned) * (size_t)kcap));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_models.p, hm.data(), sizeof(double) * hm.size(),
                                    cudaMemcpyHostToDevice, st));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_msrc.p, hoff.data(), sizeof(int) * hoff.size(),
                                    cudaMemcpyHostToDevice, st));
    launch_score(corr->corr6, corr->corr6f, corr->bounds, (int)n, ctx->d_models.as<double>(),
                 ctx->d_msrc.as<int>(), H, num_segs, seg_len, max_residual, kcap,
                 ctx->d_part_cnt.as<unsigned>(), ctx->d_cnt.as<unsigned>(), st);
    PPSFM_CUDA(ctx, cudaMemcpyAsync(counts_out, ctx->d_cnt.p, sizeof(unsigned) * (size_t)H,
                                    cudaMemcpyDeviceToHost, st));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(st));
    PPSFM_CUDA(ctx, cudaGetLastError());
    return PPSFM_OK;
  };
  rc = body();
  FreeCorr(corr);
  return rc;
}

int ppsfm_line_residuals(ppsfm_ctx* ctx, const double* lines, const double* points, size_t n,
                         const double* models, size_t num_models, double max_residual,
                         double* residuals_out, uint64_t* num_inliers_out,
                         double* residual_sum_out) {
  if (!ctx || !models) return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  if (n == 0 || num_models == 0) {
    for (size_t k = 0; k < num_models; ++k) {
      if (num_inliers_out) num_inliers_out[k] = 0;
      if (residual_sum_out) residual_sum_out[k] = 0.0;
    }
    return PPSFM_OK;
  }
  cudaSetDevice(ctx->device);
  ppsfm_corr* corr = nullptr;
  int rc = UploadCorr(ctx, lines, nullptr, points, n, true, &corr);
  if (rc != PPSFM_OK) return rc;
  cudaStream_t st = ctx->stream;
  auto body = [&]() -> int {
    const size_t batch = std::max<size_t>(1, std::min<size_t>(num_models, (64u << 20) / (8 * n)));
    PPSFM_CUDA(ctx, ctx->d_emodels.reserve(sizeof(double) * 12 * batch));
    PPSFM_CUDA(ctx, ctx->d_rbuf.reserve(sizeof(double) * batch * n));
    PPSFM_CUDA(ctx, ctx->d_ecnt.reserve(sizeof(unsigned long long) * batch));
    PPSFM_CUDA(ctx, ctx->d_esum.reserve(sizeof(double) * batch));
    std::vector<unsigned long long> cnt(batch);
    for (size_t k0 = 0; k0 < num_models; k0 += batch) {
      const size_t nb = std::min(batch, num_models - k0);
      PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_emodels.p, models + 12 * k0, sizeof(double) * 12 * nb,
                                      cudaMemcpyHostToDevice, st));
      launch_exact(corr->corr6, (int)n, ctx->d_emodels.as<double>(), (int)nb, max_residual,
                   ctx->d_rbuf.as<double>(), nullptr, ctx->d_ecnt.as<unsigned long long>(),
                   ctx->d_esum.as<double>(), st);
      if (residuals_out)
        PPSFM_CUDA(ctx, cudaMemcpyAsync(residuals_out + k0 * n, ctx->d_rbuf.p,
                                        sizeof(double) * nb * n, cudaMemcpyDeviceToHost, st));
      if (residual_sum_out)
        PPSFM_CUDA(ctx, cudaMemcpyAsync(residual_sum_out + k0, ctx->d_esum.p, sizeof(double) * nb,
                                        cudaMemcpyDeviceToHost, st));
      PPSFM_CUDA(ctx, cudaMemcpyAsync(cnt.data(), ctx->d_ecnt.p, sizeof(unsigned long long) * nb,
                                      cudaMemcpyDeviceToHost, st));
      PPSFM_CUDA(ctx, cudaStreamSynchronize(st));
      if (num_inliers_out)
        for (size_t k = 0; k < nb; ++k) num_inliers_out[k0 + k] = cnt[k];
    }
    PPSFM_CUDA(ctx, cudaGetLastError());
    return PPSFM_OK;
  };
  rc = body();
  FreeCorr(corr);
  return rc;
}

void ppsfm_get_ransac_timing(const ppsfm_ctx* ctx, ppsfm_ransac_timing* out) {
  if (ctx && out) *out = ctx->timing;
}

}  // extern "C"
