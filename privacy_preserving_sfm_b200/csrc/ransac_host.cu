// ransac_host.cu — host side of the absolute-pose RANSAC path and the C-ABI entry points.
//
// RANSAC<P6LEstimator, InlierSupportMeasurer, RandomSampler>::Estimate
// (src/optim/ransac.h:144-278) is a serial loop with loop-carried state.  Here hypotheses are
// sampled on the host (the sampler must reproduce libstdc++'s mt19937 +
// uniform_int_distribution stream, src/util/random.h:88-128), solved and scored on the GPU in
// waves, and the loop-carried logic (best-so-far, dyn_max_num_trials, abort inside the model
// loop) is replayed literally on the host over per-model supports.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

#include "common.h"
#include "ransac_kernels.h"

namespace ppsfm {
namespace {

// RANSAC<P6LEstimator>::ComputeNumTrials, src/optim/ransac.h:158-176 (kMinNumSamples = 6)
size_t ComputeNumTrials(size_t num_inliers, size_t num_samples, double confidence,
                        double multiplier) {
  const double inlier_ratio = num_inliers / static_cast<double>(num_samples);
  const double nom = 1 - confidence;
  if (nom <= 0) return std::numeric_limits<size_t>::max();
  const double denom = 1 - std::pow(inlier_ratio, 6);
  if (denom <= 0) return 1;
  return static_cast<size_t>(std::ceil(std::log(nom) / std::log(denom) * multiplier));
}

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

struct Support {
  size_t num_inliers = 0;
  double residual_sum = std::numeric_limits<double>::max();  // support_measurement.h:51-52
};

float EventMs(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

}  // namespace

// Chooses the segment count so that the scoring grid is a few waves of 148 x 4 resident CTAs.
static void ChooseSegments(const ppsfm_ctx* ctx, int n, int kcap_blocks, int* num_segs,
                           int* seg_len) {
  const int resident = ctx->num_sms * 4;
  int target_blocks = resident * 4;
  int segs = (target_blocks + kcap_blocks - 1) / std::max(1, kcap_blocks);
  // model blocks beyond K exit immediately; on average half the capacity is live.
  segs = std::max(1, std::min(segs * 2, 64));
  segs = std::max(1, std::min(ppsfm::tune_int("PPSFM_SCORE_SEGS", segs), 256));
  int len = (n + segs - 1) / segs;
  len = std::max(256, ((len + 127) / 128) * 128);
  segs = (n + len - 1) / len;
  *num_segs = std::max(1, segs);
  *seg_len = len;
}

// Runs trials [0, ...) of the RANSAC loop on a resident correspondence set.
int RansacResident(ppsfm_ctx* ctx, const ppsfm_corr* corr, const ppsfm_ransac_options* opt_in,
                   ppsfm_ransac_report* report, uint8_t* inlier_mask) {
  if (!ctx || !corr || !opt_in || !report) return fail(ctx, PPSFM_ERR_INVALID, "null argument");
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
  size_t dyn_max_num_trials = max_num_trials;
  Support best;
  bool have_best = false;       // false while `best` is the initial {0, DBL_MAX}
  bool best_sum_known = true;   // residual_sum of `best` is an exact index-order sum
  double best_model[12] = {0};
  bool abort = false;
  bool finished = false;
  size_t reported_trials = max_num_trials;
  uint64_t scored = 0;

  HostSampler sampler;
  sampler.Initialize(n);
  cudaStream_t st = ctx->stream;

  // Wave schedule: the first wave covers at least min_num_trials (and a floor that fills the
  // GPU); later waves run up to the current dyn_max_num_trials.
  const size_t kWaveFloor = 1024;
  const size_t kWaveCap = 1u << 17;
  size_t t_begin = 0;
  float total_ms = 0.f;
  while (!finished && t_begin < max_num_trials) {
    size_t want_end = std::max<size_t>(opt.min_num_trials, t_begin + kWaveFloor);
    if (dyn_max_num_trials != std::numeric_limits<size_t>::max())
      want_end = std::max(want_end, std::min(dyn_max_num_trials + 1, max_num_trials));
    size_t t_end = std::min(max_num_trials, want_end);
    t_end = std::min(t_end, t_begin + kWaveCap);
    const int H = static_cast<int>(t_end - t_begin);
    const int kcap = 8 * H;

    // ---- sample on the host (A7), keep a PRNG snapshot for the rewind after an abort
    const std::mt19937 prng_at_wave_start = ctx->prng;
    PPSFM_CUDA(ctx, ctx->h_samples.reserve(sizeof(uint32_t) * 6 * (size_t)H));
    uint32_t* hs = ctx->h_samples.as<uint32_t>();
    for (int t = 0; t < H; ++t) sampler.Sample(ctx->prng, hs + 6 * (size_t)t);

    // ---- device buffers
    PPSFM_CUDA(ctx, ctx->d_samples.reserve(sizeof(uint32_t) * 6 * (size_t)H));
    PPSFM_CUDA(ctx, ctx->d_models.reserve(sizeof(double) * 96 * (size_t)H));
    PPSFM_CUDA(ctx, ctx->d_num_models.reserve(sizeof(int) * (size_t)H));
    PPSFM_CUDA(ctx, ctx->d_msrc.reserve(sizeof(int) * ((size_t)H + 1)));
    int num_segs, seg_len;
    ChooseSegments(ctx, (int)n, (kcap + 255) / 256, &num_segs, &seg_len);  // 256 models per CTA
    PPSFM_CUDA(ctx, ctx->d_part_cnt.reserve(sizeof(unsigned) * (size_t)num_segs * kcap));
    PPSFM_CUDA(ctx, ctx->d_cnt.reserve(sizeof(unsigned) * (size_t)kcap));
    PPSFM_CUDA(ctx, ctx->h_num_models.reserve(sizeof(int) * ((size_t)H + 1)));
    PPSFM_CUDA(ctx, ctx->h_cnt.reserve(sizeof(unsigned) * (size_t)kcap));

    // ---- solve + score on the GPU
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_samples.p, hs, sizeof(uint32_t) * 6 * (size_t)H,
                                    cudaMemcpyHostToDevice, st));
    PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    launch_p6l_solve(corr->corr6, corr->aligned, ctx->d_samples.as<uint32_t>(), H,
                     ctx->d_models.as<double>(), ctx->d_num_models.as<int>(), st);
    launch_model_offsets(ctx->d_num_models.as<int>(), H, ctx->d_msrc.as<int>(), st);
    PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    launch_score(corr->corr6, corr->corr6f, corr->bounds, (int)n, ctx->d_models.as<double>(), ctx->d_msrc.as<int>(), H,
                 num_segs, seg_len, max_residual, kcap, ctx->d_part_cnt.as<unsigned>(),
                 ctx->d_cnt.as<unsigned>(), st);
    PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    ctx->timing.kernel_launches += 4;
    ctx->timing.score_launches += 1;
    int* h_off = ctx->h_num_models.as<int>();
    PPSFM_CUDA(ctx, cudaMemcpyAsync(h_off, ctx->d_msrc.p, sizeof(int) * ((size_t)H + 1),
                                    cudaMemcpyDeviceToHost, st));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(st));
    const int K = h_off[H];
    unsigned* h_cnt = ctx->h_cnt.as<unsigned>();
    if (K > 0) {
      PPSFM_CUDA(ctx, cudaMemcpyAsync(h_cnt, ctx->d_cnt.p, sizeof(unsigned) * (size_t)K,
                                      cudaMemcpyDeviceToHost, st));
      PPSFM_CUDA(ctx, cudaStreamSynchronize(st));
    }
    ctx->timing.solve_ms += EventMs(ctx->ev[0], ctx->ev[1]);
    ctx->timing.score_ms += EventMs(ctx->ev[1], ctx->ev[2]);
    total_ms += EventMs(ctx->ev[0], ctx->ev[2]);
    ctx->timing.score_pairs += (uint64_t)K * n;

    // ---- pass 1 (counts only): models that beat or tie the running best count.  Only a TIE
    // needs residual sums (InlierSupportMeasurer::Compare, support_measurement.cc:52-60), and
    // those must be index-order sums to match the reference bit for bit.
    std::vector<int> cand;
    bool has_tie = false;
    {
      size_t b = best.num_inliers;
      bool have = have_best;
      for (int k = 0; k < K; ++k) {
        const size_t c = h_cnt[k];
        if (!have || c > b) {
          cand.push_back(k);
          b = c;
          have = true;
        } else if (c == b) {
          cand.push_back(k);
          has_tie = true;
        }
      }
    }
    auto model_src = [&](int k) -> size_t {
      const int t = int(std::upper_bound(h_off, h_off + H + 1, k) - h_off) - 1;
      return (size_t)t * 96 + (size_t)(k - h_off[t]) * 12;
    };
    const int E = (int)cand.size();
    std::vector<double> cand_sum;
    if (has_tie) {
      // exact (index-order) supports for every candidate of this wave (+ the carried best)
      const bool carry = have_best && !best_sum_known;
      const int EE = E + (carry ? 1 : 0);
      cand_sum.resize(EE);
      PPSFM_CUDA(ctx, ctx->h_esum.reserve(sizeof(double) * (size_t)EE));
      PPSFM_CUDA(ctx, ctx->h_ecnt.reserve(sizeof(unsigned long long) * (size_t)EE));
      const int kBatch = 32;
      PPSFM_CUDA(ctx, ctx->d_emodels.reserve(sizeof(double) * 12 * kBatch));
      PPSFM_CUDA(ctx, ctx->d_rbuf.reserve(sizeof(double) * (size_t)kBatch * n));
      PPSFM_CUDA(ctx, ctx->d_ecnt.reserve(sizeof(unsigned long long) * kBatch));
      PPSFM_CUDA(ctx, ctx->d_esum.reserve(sizeof(double) * kBatch));
      for (int e0 = 0; e0 < EE; e0 += kBatch) {
        const int ne = std::min(kBatch, EE - e0);
        for (int e = e0; e < e0 + ne; ++e) {
          double* dst = ctx->d_emodels.as<double>() + (size_t)(e - e0) * 12;
          if (e < E)
            PPSFM_CUDA(ctx, cudaMemcpyAsync(dst, ctx->d_models.as<double>() + model_src(cand[e]),
                                            sizeof(double) * 12, cudaMemcpyDeviceToDevice, st));
          else
            PPSFM_CUDA(ctx, cudaMemcpyAsync(dst, best_model, sizeof(double) * 12,
                                            cudaMemcpyHostToDevice, st));
        }
        PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
        launch_exact(corr->corr6, (int)n, ctx->d_emodels.as<double>(), ne, max_residual,
                     ctx->d_rbuf.as<double>(), nullptr, ctx->d_ecnt.as<unsigned long long>(),
                     ctx->d_esum.as<double>(), st);
        PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
        ctx->timing.kernel_launches += 2;
        PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_esum.as<double>() + e0, ctx->d_esum.p,
                                        sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
        PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_ecnt.as<unsigned long long>() + e0, ctx->d_ecnt.p,
                                        sizeof(unsigned long long) * ne, cudaMemcpyDeviceToHost,
                                        st));
        PPSFM_CUDA(ctx, cudaStreamSynchronize(st));
        const float ems = EventMs(ctx->ev[3], ctx->ev[4]);
        ctx->timing.exact_ms += ems;
        total_ms += ems;
      }
      for (int e = 0; e < EE; ++e) {
        cand_sum[e] = ctx->h_esum.as<double>()[e];
        const size_t want = e < E ? (size_t)h_cnt[cand[e]] : best.num_inliers;
        if (ctx->h_ecnt.as<unsigned long long>()[e] != want)
          return fail(ctx, PPSFM_ERR_CUDA, "internal: exact/segmented inlier counts differ");
      }
      if (carry) {
        best.residual_sum = cand_sum[E];
        best_sum_known = true;
      }
    }

    // ---- pass 2: literal replay of src/optim/ransac.h:213-249 over this wave.
    // `abort` set at trial t means: samples were drawn for trials 0..t, and the loop reports
    // num_trials = t + 2 (the `if (abort) { num_trials += 1; break; }` at the top of the next
    // iteration) unless t + 1 already equals max_num_trials.
    size_t ci = 0;       // cursor into cand
    int best_k = -1;     // compact id of the best model if it was set in this wave
    for (size_t trial = t_begin; trial < t_end && !abort; ++trial) {
      const int lt = (int)(trial - t_begin);
      const int k0 = h_off[lt], k1 = h_off[lt + 1];
      for (int k = k0; k < k1; ++k) {
        ++scored;
        while (ci < cand.size() && cand[ci] < k) ++ci;
        if (ci < cand.size() && cand[ci] == k) {
          const size_t c = h_cnt[k];
          bool better;
          if (!have_best) {
            better = true;  // {c, sum} vs the initial {0, DBL_MAX}: more inliers, or 0 < DBL_MAX
          } else if (c > best.num_inliers) {
            better = true;
          } else if (c == best.num_inliers) {
            // tie: both sums are index-order exact here (has_tie forced the exact pass)
            better = cand_sum[ci] < best.residual_sum;
          } else {
            better = false;
          }
          if (better) {
            have_best = true;
            best.num_inliers = c;
            if (has_tie) {
              best.residual_sum = cand_sum[ci];
              best_sum_known = true;
            } else {
              best_sum_known = false;
            }
            best_k = k;
            report->best_trial = (int64_t)trial;
            report->best_model_idx = k - k0;
            dyn_max_num_trials = ComputeNumTrials(best.num_inliers, n, opt.confidence,
                                                  opt.dyn_num_trials_multiplier);
          }
        }
        if (trial >= dyn_max_num_trials && trial >= opt.min_num_trials) {
          abort = true;
          const size_t t_abort = trial;
          reported_trials = (t_abort + 1 < max_num_trials) ? t_abort + 2 : max_num_trials;
          // the reference drew samples for trials 0..t_abort only: rewind the generator
          ctx->prng = prng_at_wave_start;
          HostSampler::Skip(ctx->prng, n, t_abort + 1 - t_begin);
          finished = true;
          break;
        }
      }
    }
    if (best_k >= 0) {  // the best model changed in this wave: bring its 12 doubles to the host
      PPSFM_CUDA(ctx, cudaMemcpyAsync(best_model, ctx->d_models.as<double>() + model_src(best_k),
                                      sizeof(best_model), cudaMemcpyDeviceToHost, st));
      PPSFM_CUDA(ctx, cudaStreamSynchronize(st));
    }
    t_begin = t_end;
  }

  report->num_trials = reported_trials;
  report->num_inliers = best.num_inliers;
  report->num_models_scored = scored;
  std::memcpy(report->model, best_model, sizeof(best_model));

  // Support + inlier mask of the best model in reference (index) order
  // (src/optim/ransac.h:251-275: the reference also rescans the best model once more).
  if (have_best) {
    PPSFM_CUDA(ctx, ctx->d_emodels.reserve(sizeof(double) * 12));
    PPSFM_CUDA(ctx, ctx->d_rbuf.reserve(sizeof(double) * n));
    PPSFM_CUDA(ctx, ctx->d_mask.reserve(n));
    PPSFM_CUDA(ctx, ctx->h_mask.reserve(n));
    PPSFM_CUDA(ctx, ctx->d_ecnt.reserve(sizeof(unsigned long long)));
    PPSFM_CUDA(ctx, ctx->d_esum.reserve(sizeof(double)));
    PPSFM_CUDA(ctx, ctx->h_esum.reserve(sizeof(double)));
    PPSFM_CUDA(ctx, ctx->h_ecnt.reserve(sizeof(unsigned long long)));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->d_emodels.p, best_model, sizeof(best_model),
                                    cudaMemcpyHostToDevice, st));
    PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    const bool want_mask = inlier_mask != nullptr && best.num_inliers >= 6;
    launch_exact(corr->corr6, (int)n, ctx->d_emodels.as<double>(), 1, max_residual,
                 ctx->d_rbuf.as<double>(), want_mask ? ctx->d_mask.as<uint8_t>() : nullptr,
                 ctx->d_ecnt.as<unsigned long long>(), ctx->d_esum.as<double>(), st);
    PPSFM_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
    ctx->timing.kernel_launches += 2;
    if (want_mask)
      PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_mask.p, ctx->d_mask.p, n, cudaMemcpyDeviceToHost, st));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_esum.p, ctx->d_esum.p, sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(ctx->h_ecnt.p, ctx->d_ecnt.p, sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost, st));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(st));
    if (ctx->h_ecnt.as<unsigned long long>()[0] != best.num_inliers)
      return fail(ctx, PPSFM_ERR_CUDA, "internal: exact/segmented inlier counts differ");
    best.residual_sum = ctx->h_esum.as<double>()[0];
    if (want_mask) std::memcpy(inlier_mask, ctx->h_mask.p, n);
    const float ems = EventMs(ctx->ev[3], ctx->ev[4]);
    ctx->timing.exact_ms += ems;
    total_ms += ems;
  }
  report->residual_sum = best.residual_sum;
  ctx->timing.total_ms = total_ms;
  PPSFM_CUDA(ctx, cudaGetLastError());
  if (best.num_inliers < 6) return PPSFM_OK;  // src/optim/ransac.h:255-259
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
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* ppsfm_last_error(const ppsfm_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

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
    launch_p6l_solve(corr->corr6, corr->aligned, ctx->d_samples.as<uint32_t>(), (int)H,
                     ctx->d_models.as<double>(), ctx->d_num_models.as<int>(), st);
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
    ChooseSegments(ctx, (int)n, (kcap + 255) / 256, &num_segs, &seg_len);
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
