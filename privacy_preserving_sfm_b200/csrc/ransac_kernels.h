// ransac_kernels.h — launch wrappers of ransac_kernels.cu (all asynchronous on `s`).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace ppsfm {

// bounds (device, 3 doubles): max |X_k|, max(|l_0|, |l_1|), max |l_2| over the set (score filter)
// corr6f (device, ceil(n/2) x 12 floats): the set rounded to float, two correspondences per record
void launch_pack_corr(const double* lines, const double* points, size_t n, double* corr6,
                      float* corr6f, double* bounds, cudaStream_t s);
// lanes_per_warp: how many lanes of each warp take a hypothesis (divergence vs. warp count);
// kSolveOctet: the eight-lanes-per-hypothesis kernel (shortest latency, twice the issue slots)
constexpr int kSolveOctet = -8;
void launch_p6l_solve(const double* corr6, const uint8_t* aligned, const uint32_t* samples,
                      int num_trials, double* models_out, int* num_models_out, cudaStream_t s,
                      int lanes_per_warp = 32, int threads_per_cta = 64);
void launch_model_offsets(const int* num_models, int num_trials, int* offsets, cudaStream_t s);
// Shape of the scoring grid (launch_score): models per CTA and resident CTAs per SM, for the
// callers that choose the segment count.
constexpr int kScoreModelsPerCta = 512;
constexpr int kScoreCtasPerSm = 2;
// 96 registers (the kernel needs 95 - 106 depending on scheduling; no spills at 96): two scoring
// CTAs (2 x 256 x 96) leave a quarter of the register file to whatever else is on the SM.
constexpr int kScoreMaxRegs = 96;
// Exact pruning state of launch_score (all device pointers; see the comment at launch_score).
struct ScorePrune {
  unsigned* best_lb = nullptr;  // running maximum of the final counts; nullptr: no pruning
  int n_first = 0;              // > 0: two phases, the first over correspondences [0, n_first)
                                // (a multiple of 128), segments of the second phase below
  int num_segs2 = 0, seg_len2 = 0;
  int* list = nullptr;          // kcap survivors of the first phase
  int* list_count = nullptr;
};
// One call sharded over the ranks of a communicator: this rank scores the model
// blocks b (kScoreModelsPerCta models each) with b % world == rank and writes zeros for the
// others, so that a sum all-reduce of cnt_out (all kcap slots) gives every rank every count.
struct ScoreShard {
  int world = 1, rank = 0;
};
// Inlier counts of every compact model.  part_cnt: max(num_segs, num_segs2) x kcap scratch;
// cnt_out: kcap (first K valid; exact for every model that can matter, see ScorePrune).
void launch_score(const double* corr6, const float* corr6f, const double* bounds, int n,
                  const double* models, const int* offsets, int num_trials, int num_segs,
                  int seg_len, double max_residual, int kcap, unsigned* part_cnt,
                  unsigned* cnt_out, cudaStream_t s, const ScorePrune& prune = ScorePrune(),
                  const ScoreShard& shard = ScoreShard());
// best_lb = max(best_lb, max of the first K counts): after the all-reduce of a sharded wave
void launch_raise_best_lb(const unsigned* cnt, const int* offsets, int num_trials,
                          unsigned* best_lb, cudaStream_t s);
// dst[e * 12 + j] = src[off[e] + j], e < ne (off on the device)
void launch_gather_models(const double* src, const long long* off, int ne, double* dst,
                          cudaStream_t s);
// Largest double r with fl(r*r) <= max_residual.
double inlier_abs_threshold(double max_residual);
// rbuf: num_e x n residuals; mask (optional): num_e x n; ecnt/esum (optional): num_e.
void launch_exact(const double* corr6, int n, const double* emodels, int num_e,
                  double max_residual, double* rbuf, uint8_t* mask, unsigned long long* ecnt,
                  double* esum, cudaStream_t s);

}  // namespace ppsfm
