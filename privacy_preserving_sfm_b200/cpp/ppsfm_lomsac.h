// ppsfm_lomsac.h — locally optimised MSAC over the RansacLib "Solver" concept.
//
// Drop-in for ransac_lib::LocallyOptimizedMSAC (lib/RansacLib/RansacLib/ransac.h:127-271; options
// :47-96, statistics :98-105), UniformSampling (sampling.h:46-135) and NumRequiredIterations
// (utils.h:110-132): same template parameters, option / statistics member names and Solver
// interface (min_sample_size, non_minimal_sample_size, num_data, MinimalSolver, NonMinimalSolver,
// EvaluateModelOnPoint, LeastSquares — lib/RansacLib/README.md:45-97), and the same draws from
// libstdc++'s std::mt19937 / std::uniform_int_distribution<int>, so a run is bit-identical to the
// reference driver on the same solver (checked in tests/test_init.py against the real header,
// compiled from /root/reference into oracle/_ref).
//
// Structure: the minimal-sample stream depends only on the seed (the local optimisation re-seeds a
// private generator on every call, ransac.h:357-358), which is what allows minimal solves and
// model scoring to be evaluated ahead in batches.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <random>
#include <vector>

namespace ppsfm {

struct RansacOptions {  // ransac.h:47-61
  uint32_t min_num_iterations_ = 100u;
  uint32_t max_num_iterations_ = 10000u;
  double success_probability_ = 0.9999;
  double squared_inlier_threshold_ = 1.0;
  unsigned int random_seed_ = 0u;
};

struct LORansacOptions : RansacOptions {  // ransac.h:65-96 (Lebeda et al., BMVC 2012, table 1)
  int num_lo_steps_ = 10;
  double threshold_multiplier_ = std::sqrt(2.0);
  int num_lsq_iterations_ = 4;
  int min_sample_multiplicator_ = 7;
  int non_min_sample_multiplier_ = 3;
  uint32_t lo_starting_iterations_ = 50u;
  bool final_least_squares_ = false;
};

struct RansacStatistics {  // ransac.h:98-105
  uint32_t num_iterations = 0;
  int best_num_inliers = 0;
  double best_model_score = std::numeric_limits<double>::max();
  double inlier_ratio = 0.0;
  std::vector<int> inlier_indices;
  int number_lo_iterations = 0;
};

namespace lomsac_detail {

inline void shuffle_prefix_all(std::mt19937* rng, std::vector<int>* v) {  // utils.h:48-57
  const int n = static_cast<int>(v->size());
  for (int i = 0; i + 1 < n; ++i) {
    std::uniform_int_distribution<int> pick(i, n - 1);
    std::swap((*v)[i], (*v)[pick(*rng)]);
  }
}

inline void shuffle_and_resize(int target, std::mt19937* rng, std::vector<int>* v) {
  shuffle_prefix_all(rng, v);  // utils.h:71-75
  v->resize(target);
}

// utils.h:110-132
inline uint32_t num_required_iterations(double inlier_ratio, double prob_missing, int sample_size,
                                        uint32_t min_it, uint32_t max_it) {
  if (inlier_ratio <= 0.0) return max_it;
  if (inlier_ratio >= 1.0) return min_it;
  const double p_bad = 1.0 - std::pow(inlier_ratio, static_cast<double>(sample_size));
  const double iters = std::ceil(std::log(prob_missing) / std::log(p_bad) + 0.5);
  return std::max(min_it, std::min(static_cast<uint32_t>(iters), max_it));
}

// sampling.h:46-135: rejection sampling when n / (n - k) < e, otherwise a full shuffle
class MinimalSampler {
 public:
  MinimalSampler(unsigned seed, int num_data, int sample_size)
      : n_(num_data), k_(sample_size), any_(0, num_data - 1) {
    rng_.seed(seed);
    by_rejection_ = static_cast<double>(n_) / static_cast<double>(n_ - k_) < M_E;
  }
  void draw(std::vector<int>* out) {
    std::vector<int>& s = *out;
    if (by_rejection_) {
      s.resize(k_);
      for (int i = 0; i < k_; ++i) {
        bool again = true;
        while (again) {
          s[i] = any_(rng_);
          again = std::find(s.begin(), s.begin() + i, s[i]) != s.begin() + i;
        }
      }
      return;
    }
    s.resize(n_);
    std::iota(s.begin(), s.end(), 0);
    if (k_ == n_) return;
    shuffle_prefix_all(&rng_, &s);
    s.resize(k_);
  }

 private:
  int n_, k_;
  bool by_rejection_ = true;
  std::mt19937 rng_;
  std::uniform_int_distribution<int> any_;
};

}  // namespace lomsac_detail

namespace lomsac_detail {
// Optional members of a Solver (not part of the RansacLib concept): ScoreModels scores all the
// models of one minimal sample at once (e.g. on a GPU) and returns false if it cannot; Materialize
// completes a model that MinimalSolver left "lazy" before it is stored or refined.
template <class S, class MV>
auto batch_score(const S& s, const MV& models, int n, double thr, double* out, int)
    -> decltype(s.ScoreModels(models, n, thr, out)) {
  return s.ScoreModels(models, n, thr, out);
}
template <class S, class MV>
bool batch_score(const S&, const MV&, int, double, double*, long) {
  return false;
}
template <class S, class M>
auto materialize(const S& s, M* m, int) -> decltype(s.Materialize(m), void()) {
  s.Materialize(m);
}
template <class S, class M>
void materialize(const S&, M*, long) {}
}  // namespace lomsac_detail

template <class Model, class ModelVector, class Solver>
class LocallyOptimizedMSAC {
 public:
  // ransac.h:134-271.  Returns the number of inliers of *best_model.
  int EstimateModel(const LORansacOptions& opt, const Solver& solver, Model* best_model,
                    RansacStatistics* statistics) const {
    RansacStatistics& st = *statistics;
    st = RansacStatistics();
    const int k = solver.min_sample_size(), n = solver.num_data();
    if (k > n || k <= 0) return 0;
    lomsac_detail::MinimalSampler sampler(opt.random_seed_, n, k);
    uint32_t max_it = std::max(opt.max_num_iterations_, opt.min_num_iterations_);
    const double thr = opt.squared_inlier_threshold_;
    const double kInf = std::numeric_limits<double>::max();
    Model best_minimal;
    double best_minimal_score = kInf;
    std::vector<int> sample(k);
    ModelVector models;
    std::vector<double> batch_scores;

    auto refresh = [&]() {  // inliers of the best model and the adaptive iteration bound
      st.best_num_inliers = Inliers(solver, *best_model, thr, &st.inlier_indices);
      st.inlier_ratio = static_cast<double>(st.best_num_inliers) / static_cast<double>(n);
    };
    auto refresh_and_rebound = [&]() {
      refresh();
      max_it = lomsac_detail::num_required_iterations(st.inlier_ratio,
                                                      1.0 - opt.success_probability_, k,
                                                      opt.min_num_iterations_,
                                                      opt.max_num_iterations_);
    };

    for (st.num_iterations = 0u; st.num_iterations < max_it; ++st.num_iterations) {
      const bool lo_start = st.num_iterations == opt.lo_starting_iterations_;
      if (lo_start && best_minimal_score < kInf) {  // first local optimisation, on the best so far
        ++st.number_lo_iterations;
        LocalOptimization(opt, solver, best_model, &st.best_model_score);
        refresh_and_rebound();
      }
      sampler.draw(&sample);
      const int num_models = solver.MinimalSolver(sample, &models);
      if (num_models <= 0) continue;
      double local_score = kInf;
      int local_id = 0;
      batch_scores.resize(num_models);
      const bool batched =
          lomsac_detail::batch_score(solver, models, num_models, thr, batch_scores.data(), 0);
      for (int m = 0; m < num_models; ++m) {
        if (!batched) lomsac_detail::materialize(solver, &models[m], 0);
        const double s = batched ? batch_scores[m] : Score(solver, models[m], thr);
        if (s < local_score) {
          local_score = s;
          local_id = m;
        }
      }
      if (!(local_score < best_minimal_score || lo_start)) continue;
      const bool improved = local_score < best_minimal_score;
      if (improved) {
        best_minimal_score = local_score;
        best_minimal = models[local_id];
        lomsac_detail::materialize(solver, &best_minimal, 0);
        KeepBetter(best_minimal_score, best_minimal, &st.best_model_score, best_model);
      }
      const bool run_lo =
          st.num_iterations >= opt.lo_starting_iterations_ && best_minimal_score < kInf;
      if (!improved && !run_lo) continue;
      if (run_lo) {
        ++st.number_lo_iterations;
        double score = best_minimal_score;
        LocalOptimization(opt, solver, &best_minimal, &score);
        KeepBetter(score, best_minimal, &st.best_model_score, best_model);
      }
      refresh_and_rebound();
    }
    if (st.num_iterations <= opt.lo_starting_iterations_ && st.best_model_score < kInf) {
      ++st.number_lo_iterations;
      LocalOptimization(opt, solver, best_model, &st.best_model_score);
      refresh();
    }
    if (opt.final_least_squares_) {
      Model refined = *best_model;
      solver.LeastSquares(st.inlier_indices, &refined);
      const double score = Score(solver, refined, thr);
      if (score < st.best_model_score) {
        st.best_model_score = score;
        *best_model = refined;
        refresh();
      }
    }
    return st.best_num_inliers;
  }

 protected:
  // MSAC (top-hat) score, ransac.h:291-305
  static double Score(const Solver& solver, const Model& model, double thr) {
    const int n = solver.num_data();
    double score = 0.0;
    for (int i = 0; i < n; ++i) score += std::min(solver.EvaluateModelOnPoint(model, i), thr);
    return score;
  }
  // strict '<', ransac.h:307-332
  static int Inliers(const Solver& solver, const Model& model, double thr, std::vector<int>* out) {
    const int n = solver.num_data();
    out->clear();
    for (int i = 0; i < n; ++i)
      if (solver.EvaluateModelOnPoint(model, i) < thr) out->push_back(i);
    return static_cast<int>(out->size());
  }
  static void KeepBetter(double score, const Model& m, double* best_score, Model* best) {
    if (score < *best_score) {
      *best_score = score;
      *best = m;
    }
  }
  // ransac.h:408-419
  static void LeastSquaresFit(const LORansacOptions& opt, double thr, const Solver& solver,
                              std::mt19937* rng, Model* model) {
    const int cap = opt.min_sample_multiplicator_ * solver.min_sample_size();
    std::vector<int> inl;
    const int num = Inliers(solver, *model, thr, &inl);
    if (num < solver.min_sample_size()) return;
    lomsac_detail::shuffle_and_resize(std::min(cap, num), rng, &inl);
    solver.LeastSquares(inl, model);
  }
  // Algorithms 2 and 3 of Lebeda et al.; ransac.h:337-406
  static void LocalOptimization(const LORansacOptions& opt, const Solver& solver, Model* best,
                                double* best_score) {
    const int n = solver.num_data();
    const int min_non_min = solver.non_minimal_sample_size();
    if (min_non_min > n) return;
    const int k = solver.min_sample_size();
    const double thr = opt.squared_inlier_threshold_, mult = opt.threshold_multiplier_;
    std::mt19937 rng;
    rng.seed(opt.random_seed_);
    Model m_init = *best;
    LeastSquaresFit(opt, thr * mult, solver, &rng, &m_init);
    double score = Score(solver, m_init, thr);
    KeepBetter(score, m_init, best_score, best);
    std::vector<int> base;
    Inliers(solver, m_init, thr, &base);
    const int non_min_size =
        std::max(min_non_min, std::min(k * opt.non_min_sample_multiplier_,
                                       static_cast<int>(base.size()) / 2));
    std::vector<int> sample;
    for (int r = 0; r < opt.num_lo_steps_; ++r) {
      sample = base;
      lomsac_detail::shuffle_and_resize(non_min_size, &rng, &sample);
      Model m;
      if (!solver.NonMinimalSolver(sample, &m)) continue;
      score = Score(solver, m, thr);
      KeepBetter(score, m, best_score, best);
      LeastSquaresFit(opt, thr, solver, &rng, &m);
      double t = mult * thr;
      const double dt = (mult - 1.0) * thr / static_cast<int>(opt.num_lsq_iterations_ - 1);
      for (int i = 0; i < opt.num_lsq_iterations_; ++i) {
        LeastSquaresFit(opt, t, solver, &rng, &m);
        score = Score(solver, m, thr);
        KeepBetter(score, m, best_score, best);
        t -= dt;
      }
    }
  }
};

}  // namespace ppsfm
