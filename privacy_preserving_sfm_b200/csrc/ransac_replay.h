// ransac_replay.h — host-only replay of the reference's sequential RANSAC loop over the inlier
// counts of one wave of trials (src/optim/ransac.h:213-249).  No CUDA in here: the functions are
// exercised on the CPU by ppsfm_selftest_replay (tests/test_abi.py), which checks the
// event-driven replay used by RansacResident against the literal model-by-model loop.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace ppsfm {

// RANSAC<P6LEstimator>::ComputeNumTrials, src/optim/ransac.h:158-176 (kMinNumSamples = 6)
inline size_t ComputeNumTrials(size_t num_inliers, size_t num_samples, double confidence,
                               double multiplier) {
  const double inlier_ratio = num_inliers / static_cast<double>(num_samples);
  const double nom = 1 - confidence;
  if (nom <= 0) return std::numeric_limits<size_t>::max();
  const double denom = 1 - std::pow(inlier_ratio, 6);
  if (denom <= 0) return 1;
  return static_cast<size_t>(std::ceil(std::log(nom) / std::log(denom) * multiplier));
}

// State the loop carries from trial to trial (and the replay from wave to wave).
struct ReplayState {
  bool have_best = false;       // false while the best support is the initial {0, DBL_MAX}
  size_t best_inliers = 0;
  double best_sum = std::numeric_limits<double>::max();  // support_measurement.h:51-52
  bool best_sum_known = true;   // best_sum is an exact index-order sum
  size_t dyn_max_num_trials = 0;
  int64_t best_trial = -1;
  int best_model_idx = -1;
};

struct ReplayParams {
  size_t t_begin = 0, t_end = 0;   // trials of the wave
  size_t num_samples = 0;          // N, for ComputeNumTrials
  size_t min_num_trials = 0;
  double confidence = 0.0, multiplier = 0.0;
};

struct ReplayOutcome {
  int best_k = -1;       // compact id of the best model if the wave set it
  int abort_model = -1;  // compact id of the model after which the loop aborted (-1: no abort)
  size_t t_abort = 0;    // its trial
  uint64_t scored = 0;   // models the reference loop visited in this wave
};

// Pass 1: models that beat or tie the running best count, in order.  Only a TIE needs residual
// sums (InlierSupportMeasurer::Compare, support_measurement.cc:52-60).
inline bool replay_candidates(const unsigned* cnt, int K, const ReplayState& st,
                              std::vector<int>* cand) {
  cand->clear();
  bool has_tie = false;
  size_t b = st.best_inliers;
  bool have = st.have_best;
  for (int k = 0; k < K; ++k) {
    const size_t c = cnt[k];
    if (!have || c > b) {
      cand->push_back(k);
      b = c;
      have = true;
    } else if (c == b) {
      cand->push_back(k);
      has_tie = true;
    }
  }
  return has_tie;
}

namespace replay_detail {
// visit of candidate cand[ci] = model k of local trial lt: the body of ransac.h:228-241
inline void visit_candidate(int k, int lt, size_t ci, const int* off, const unsigned* cnt,
                            const double* cand_sum, const ReplayParams& p, ReplayState* st,
                            ReplayOutcome* out) {
  const size_t c = cnt[k];
  bool better;
  if (!st->have_best) {
    better = true;  // {c, sum} vs the initial {0, DBL_MAX}: more inliers, or 0 < DBL_MAX
  } else if (c > st->best_inliers) {
    better = true;
  } else if (c == st->best_inliers) {
    better = cand_sum[ci] < st->best_sum;  // tie: both sums are index-order exact here
  } else {
    better = false;
  }
  if (!better) return;
  st->have_best = true;
  st->best_inliers = c;
  if (cand_sum) {
    st->best_sum = cand_sum[ci];
    st->best_sum_known = true;
  } else {
    st->best_sum_known = false;
  }
  out->best_k = k;
  st->best_trial = (int64_t)(p.t_begin + lt);
  st->best_model_idx = k - off[lt];
  st->dyn_max_num_trials = ComputeNumTrials(st->best_inliers, p.num_samples, p.confidence,
                                            p.multiplier);
}
}  // namespace replay_detail

// Pass 2, literal: every model of every trial, as the reference walks them.
// off[H + 1]: exclusive scan of the models per trial; cand / cand_sum from pass 1 (cand_sum is
// null unless pass 1 found a tie).
inline ReplayOutcome replay_wave_literal(const int* off, int H, const unsigned* cnt,
                                         const std::vector<int>& cand, const double* cand_sum,
                                         const ReplayParams& p, ReplayState* st) {
  ReplayOutcome out;
  size_t ci = 0;
  for (size_t trial = p.t_begin; trial < p.t_end; ++trial) {
    const int lt = (int)(trial - p.t_begin);
    for (int k = off[lt]; k < off[lt + 1]; ++k) {
      ++out.scored;
      while (ci < cand.size() && cand[ci] < k) ++ci;
      if (ci < cand.size() && cand[ci] == k)
        replay_detail::visit_candidate(k, lt, ci, off, cnt, cand_sum, p, st, &out);
      if (trial >= st->dyn_max_num_trials && trial >= p.min_num_trials) {
        out.abort_model = k;
        out.t_abort = trial;
        return out;
      }
    }
  }
  (void)H;
  return out;
}

// Pass 2, event-driven (what RansacResident runs).  Only two kinds of visits change state: a
// candidate (may become the best model and lower dyn_max_num_trials) and the first model at which
// the abort test `trial >= dyn_max_num_trials && trial >= min_num_trials` holds (it is checked
// after every model, so an empty trial cannot abort).  For a given bound the abort can only
// trigger at the first model of the first non-empty trial at or above it, i.e. at compact index
// off[bound - t_begin]; the replay jumps from one such visit to the next.
inline ReplayOutcome replay_wave(const int* off, int H, const unsigned* cnt,
                                 const std::vector<int>& cand, const double* cand_sum,
                                 const ReplayParams& p, ReplayState* st) {
  ReplayOutcome out;
  const int K = off[H];
  auto local_trial = [&](int k) { return int(std::upper_bound(off, off + H + 1, k) - off) - 1; };
  size_t ci = 0;
  int k_done = 0;  // models [0, k_done) of the wave have been visited
  while (true) {
    const size_t bound = std::max(st->dyn_max_num_trials, p.min_num_trials);
    int abort_k = K;
    if (bound < p.t_end) abort_k = std::max(k_done, off[bound > p.t_begin ? bound - p.t_begin : 0]);
    while (ci < cand.size() && cand[ci] < k_done) ++ci;
    const int next_cand = ci < cand.size() ? cand[ci] : K;
    if (next_cand >= K && abort_k >= K) break;
    int k;
    if (next_cand <= abort_k) {  // the candidate is visited first (or is the aborting model)
      k = next_cand;
      const int lt = local_trial(k);
      replay_detail::visit_candidate(k, lt, ci, off, cnt, cand_sum, p, st, &out);
      ++ci;
      k_done = k + 1;
      const size_t trial = p.t_begin + lt;
      if (!(trial >= st->dyn_max_num_trials && trial >= p.min_num_trials)) continue;
    } else {
      k = abort_k;
    }
    out.abort_model = k;
    out.t_abort = p.t_begin + local_trial(k);
    out.scored = (uint64_t)k + 1;
    return out;
  }
  out.scored = (uint64_t)K;
  return out;
}

}  // namespace ppsfm
