/* ppsfm_b200.h — C-ABI of libppsfm_b200.so, the B200 (sm_100a) implementation of the
 * privacy-preserving-SfM hot path (line-lifted absolute pose under RANSAC, line-reprojection
 * bundle adjustment).
 *
 * The reference (colmap/privacy_preserving_sfm) has no FFI layer; the seam is its C++ API.
 * Each entry point below names the reference interface it replaces (file:line, relative to the
 * upstream repository).  The header-only C++ adaptor in privacy_preserving_sfm_b200/cpp/
 * re-creates the reference signatures on top of this ABI.
 *
 * Conventions
 *   - all pointers are HOST memory, caller-allocated, unless the name says `_resident`;
 *   - calls are blocking; one context per host thread (a context owns one CUDA device, its
 *     streams, scratch buffers and the PRNG that mirrors the reference's thread_local mt19937);
 *   - return value: PPSFM_OK (0); PPSFM_NO_SOLUTION (1) where the reference returns `false`;
 *     negative = contract violation / CUDA failure (where the reference CHECK-aborts) —
 *     ppsfm_last_error() describes it.  There is NO CPU fallback: without a CUDA device
 *     ppsfm_ctx_create fails with PPSFM_ERR_CUDA.
 *   - lines  : n x 3 doubles row-major (a,b,c), normalised camera coordinates, ||(a,b)|| = 1
 *              (FeatureLine::Line(), src/feature/types.h:98-138)
 *     aligned: n bytes (FeatureLine::IsAligned()); may be NULL (= all false)
 *     points : n x 3 doubles row-major (std::vector<Eigen::Vector3d> memory)
 *     model  : 12 doubles = Eigen::Matrix3x4d column-major (R col0, R col1, R col2, t)
 *     qvec   : (w, x, y, z)  (src/base/pose.cc:41-44)
 */
#ifndef PPSFM_B200_H_
#define PPSFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPSFM_OK 0
#define PPSFM_NO_SOLUTION 1
#define PPSFM_ERR_INVALID (-1)
#define PPSFM_ERR_CUDA (-2)
#define PPSFM_ERR_NCCL (-3)

typedef struct ppsfm_ctx ppsfm_ctx;
typedef struct ppsfm_corr ppsfm_corr; /* correspondence set resident in HBM */

/* RANSACOptions, src/optim/ransac.h:47-76 (same fields, same defaults via ppsfm_ransac_options_default) */
typedef struct {
  double max_error;
  double min_inlier_ratio;
  double confidence;
  double dyn_num_trials_multiplier;
  uint64_t min_num_trials;
  uint64_t max_num_trials;
} ppsfm_ransac_options;

/* RANSAC<P6LEstimator>::Report, src/optim/ransac.h:82-99 (+ provenance of the winner) */
typedef struct {
  int32_t success;
  uint64_t num_trials;
  uint64_t num_inliers;  /* support.num_inliers */
  double residual_sum;   /* support.residual_sum (summed in index order, as the reference) */
  double model[12];
  int64_t best_trial;     /* trial index that produced the best model, -1 if none */
  int32_t best_model_idx; /* index inside that trial's solver output */
  uint64_t num_models_scored;
} ppsfm_ransac_report;

/* Device-side timings of the last RANSAC call (CUDA events on the context's stream). */
typedef struct {
  double solve_ms;   /* batched P6L / re3q3 kernel(s) */
  double score_ms;   /* batched scoring kernel(s) (the dominant kernel) */
  double exact_ms;   /* exact-order support + mask kernels */
  double total_ms;   /* first launch to last kernel end */
  uint64_t score_pairs;    /* (model, correspondence) pairs evaluated by the scoring kernel */
  uint64_t score_launches; /* scoring-kernel launches */
  uint64_t kernel_launches; /* all kernels launched by the call */
} ppsfm_ransac_timing;

/* ---- context -------------------------------------------------------------------------- */
int ppsfm_ctx_create(int device, ppsfm_ctx** out);
void ppsfm_ctx_destroy(ppsfm_ctx* ctx);
const char* ppsfm_last_error(const ppsfm_ctx* ctx);
const char* ppsfm_version(void);

/* SetPRNGSeed, src/util/random.cc:38-50 (the context's generator starts at seed 0 like
 * kDefaultPRNGSeed, src/util/random.h:46). */
void ppsfm_set_prng_seed(ppsfm_ctx* ctx, uint32_t seed);
/* next raw mt19937 output without advancing (state fingerprint for parity tests) */
uint32_t ppsfm_prng_peek(const ppsfm_ctx* ctx);

void ppsfm_ransac_options_default(ppsfm_ransac_options* opt);

/* RANSAC<P6LEstimator>::ComputeNumTrials, src/optim/ransac.h:158-176 */
uint64_t ppsfm_compute_num_trials(uint64_t num_inliers, uint64_t num_samples, double confidence,
                                  double num_trials_multiplier);

/* RandomSampler::Sample x num_trials on a freshly Initialize()d sampler,
 * src/optim/random_sampler.cc:40-62 + src/util/random.h:120-128; advances the context PRNG. */
int ppsfm_sample_table(ppsfm_ctx* ctx, size_t n, size_t num_trials, uint32_t* table_out);

/* ---- A5 + A8: batched scoring ------------------------------------------------------------
 * P6LEstimator::Residuals -> ComputeSquaredLineReprojectionError (src/estimators/utils.cc:40-89)
 * for K models at once, fused with InlierSupportMeasurer::Evaluate
 * (src/optim/support_measurement.cc:36-60).  residuals_out (K x n, may be NULL),
 * num_inliers_out[K], residual_sum_out[K] (index-order sums, bit-identical to the reference). */
int ppsfm_line_residuals(ppsfm_ctx* ctx, const double* lines, const double* points, size_t n,
                         const double* models, size_t num_models, double max_residual,
                         double* residuals_out, uint64_t* num_inliers_out,
                         double* residual_sum_out);

/* ---- A3 + A4: batched minimal solver -------------------------------------------------------
 * P6LEstimator::Estimate (src/estimators/absolute_pose.cc:79-162) incl. re3q3
 * (lib/re3q3/re3q3/re3q3.h:16-200) for H samples of 6 correspondences, one thread per
 * hypothesis.  models_out: H x 8 x 12, num_models_out: H. */
int ppsfm_p6l_solve_batch(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                          const double* points, size_t n, const uint32_t* sample_idx,
                          size_t num_samples, double* models_out, int32_t* num_models_out);

/* ---- A6 + A7: RANSAC<P6LEstimator, InlierSupportMeasurer, RandomSampler>::Estimate ---------
 * src/optim/ransac.h:144-278.  Hypotheses are generated and scored on the GPU in waves; the
 * sequential best-so-far / adaptive-abort logic is replayed exactly on the host.
 * inlier_mask: n bytes or NULL. Returns PPSFM_OK also when report->success == 0. */
int ppsfm_ransac_p6l(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                     const double* points, size_t n, const ppsfm_ransac_options* options,
                     ppsfm_ransac_report* report, uint8_t* inlier_mask);

/* ---- A9: EstimateAbsolutePoseFromLines, src/estimators/pose.cc:52-94 ------------------------
 * Returns PPSFM_OK (true) or PPSFM_NO_SOLUTION (false). report may be NULL. */
int ppsfm_estimate_absolute_pose_from_lines(ppsfm_ctx* ctx, const double* lines,
                                            const uint8_t* aligned, const double* points,
                                            size_t n, const ppsfm_ransac_options* options,
                                            double* qvec, double* tvec, uint64_t* num_inliers,
                                            uint8_t* inlier_mask, ppsfm_ransac_report* report);

/* ---- resident variants (inputs already in HBM; used for the kernel-side metric) ------------ */
int ppsfm_corr_upload(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                      const double* points, size_t n, ppsfm_corr** out);
void ppsfm_corr_free(ppsfm_ctx* ctx, ppsfm_corr* corr);
int ppsfm_ransac_p6l_resident(ppsfm_ctx* ctx, const ppsfm_corr* corr,
                              const ppsfm_ransac_options* options, ppsfm_ransac_report* report,
                              uint8_t* inlier_mask);
void ppsfm_get_ransac_timing(const ppsfm_ctx* ctx, ppsfm_ransac_timing* out);

/* ---- measurement helpers (bench.py only; not part of the reference surface) ------------------
 * FP64 issue rate in 1e12 thread-instructions/s: fused (DFMA) and unfused (DMUL/DADD mix). */
int ppsfm_bench_fp64_peak(ppsfm_ctx* ctx, double* dfma_tips, double* dmuladd_tips);
/* Evict L2 by writing `bytes` of scratch HBM (blocking). */
int ppsfm_bench_l2_flush(ppsfm_ctx* ctx, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PPSFM_B200_H_ */
