/* ppsfm_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a dependency-free CPU restatement of the
 * reference's (colmap/privacy_preserving_sfm) hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (privacy_preserving_sfm_b200/) never links, imports or calls anything declared here.
 *
 * Parity status:
 *   - line residual / support / sampler / RANSAC control flow: restated 1:1 from reference
 *     sources that are fully in-tree (no third-party arithmetic) — pinned by construction and
 *     by known-answer tests.
 *   - P6L + re3q3: every expression the reference writes out itself (tt / Rcoeffs rows,
 *     rotation_to_e3q3, the 33 resultant coefficients a11..a313, c(0)..c(8), A(x), the Cramer
 *     quotients, cayley_param) is evaluated operation by operation in the reference's order
 *     (re3q3_resultant.inc is generated from re3q3.h:84-150 and pinned to those expressions by
 *     tests/golden/re3q3_resultant_vectors.json).  What the reference delegates to Eigen
 *     (3x3 determinant, PartialPivLU::solve, the 3x3 * 3x9 products, EigenSolver<8x8>) is NOT in
 *     /root/reference and not installed: restated from Eigen 3.3's algorithms; pinned by the
 *     reference's known-answer properties (lib/re3q3/test_re3q3.cpp) and numpy.roots; bit-level
 *     parity with an Eigen build is UNPINNED for exactly those calls (they live in
 *     eigen_restated.h).
 *   - Pinned against the reference's OWN SOURCES: oracle/build_ref.sh compiles
 *     absolute_pose.cc, re3q3.h, utils.cc, ransac.h, random_sampler.cc, random.cc and
 *     support_measurement.cc from /root/reference against Eigen / glog stand-ins
 *     (oracle/ref/shim/, whose Eigen calls are eigen_restated.h) into oracle/_ref/libref_p6l.so;
 *     tests/test_ref_p6l.py requires this library and the reference build to agree bit for bit
 *     (P6L, re3q3, residuals, supports, sample tables, whole RANSAC calls incl. BASELINE
 *     config 2 at full size).
 *   - BA / pose refinement: Ceres is not in /root/reference; the SOLVER is parity-UNPINNED (see
 *     ba_oracle.h); the cost functors and camera models are pinned the same way
 *     (oracle/_ref/libref_cost.so, tests/test_ref_cost.py).
 *
 * Layout conventions (shared with include/ppsfm_b200.h):
 *   lines  : N x 3 doubles, row-major (a, b, c) per correspondence  (FeatureLine::Line())
 *   aligned: N bytes, FeatureLine::IsAligned()
 *   points : N x 3 doubles, row-major (X, Y, Z)
 *   model  : 12 doubles, Eigen::Matrix3x4d column-major: R col0, R col1, R col2, t
 */
#ifndef PPSFM_ORACLE_H_
#define PPSFM_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double max_error;                 /* src/optim/ransac.h:50 */
  double min_inlier_ratio;          /* :54 */
  double confidence;                /* :58 */
  double dyn_num_trials_multiplier; /* :62 */
  uint64_t min_num_trials;          /* :65 */
  uint64_t max_num_trials;          /* :66 */
} orc_ransac_options;

typedef struct {
  int32_t success;        /* Report::success            src/optim/ransac.h:86 */
  uint64_t num_trials;    /* Report::num_trials         :89 */
  uint64_t num_inliers;   /* Report::support.num_inliers */
  double residual_sum;    /* Report::support.residual_sum */
  double model[12];       /* Report::model (col-major 3x4) */
  int64_t best_trial;     /* extra: trial index that produced the best model (-1 if none) */
  int32_t best_model_idx; /* extra: index of the best model inside that trial's solver output */
  uint64_t num_models_scored; /* extra: total models scored (for throughput accounting) */
} orc_ransac_report;

/* util/random.{h,cc}: SetPRNGSeed / thread-local mt19937 (one generator per oracle process). */
void orc_set_prng_seed(uint32_t seed);
/* Next raw 32-bit output WITHOUT advancing the stream (fingerprint of the PRNG state). */
uint32_t orc_prng_peek(void);

/* src/estimators/utils.cc:40-89 ComputeSquaredLineReprojectionError */
void orc_line_residuals(const double* lines, const double* points, size_t n,
                        const double* model, double* residuals_out);

/* src/optim/support_measurement.cc:36-60 InlierSupportMeasurer::Evaluate */
void orc_inlier_support(const double* residuals, size_t n, double max_residual,
                        uint64_t* num_inliers, double* residual_sum);
/* src/optim/support_measurement.cc:62-83 MEstimatorSupportMeasurer::Evaluate */
void orc_mestimator_support(const double* residuals, size_t n, double max_residual,
                            uint64_t* num_inliers, double* score);

/* src/optim/ransac.h:158-176 RANSAC<P6LEstimator>::ComputeNumTrials (kMinNumSamples = 6) */
uint64_t orc_compute_num_trials(uint64_t num_inliers, uint64_t num_samples, double confidence,
                                double num_trials_multiplier);

/* src/optim/random_sampler.cc:40-62 + src/util/random.h:120-128: draws `num_trials` consecutive
 * samples of 6 from a freshly Initialize()d sampler over n items using the process PRNG. */
void orc_sample_table(size_t n, size_t num_trials, uint32_t* table_out /* num_trials x 6 */);

/* lib/re3q3/re3q3/re3q3.h:16-200.  coeffs: 3x10 row-major (row = equation, monomial order
 * x^2 xy xz y^2 yz z^2 x y z 1).  solutions: 3x8 column-major (solution k at [3k..3k+2]). */
int orc_re3q3(const double* coeffs, double* solutions);
/* re3q3.h:84-150: P (3x7 row-major, after `P = -A.lu().solve(P)`) -> a11..a313 (33) and c(0..8);
 * re3q3.h:177-188: y, z for a root x.  Hooks for tests/golden/re3q3_resultant_vectors.json, which
 * holds the REFERENCE's own expressions evaluated in IEEE double by tests/golden/make_re3q3_golden.py. */
void orc_re3q3_resultant(const double* P, double* a_out, double* c_out);
void orc_re3q3_backsubstitute(const double* a, double x, double* yz_out);
/* real roots of c[0] x^8 + ... + c[8] as the re3q3 companion/eigenvalue step returns them
 * (|imag| <= 1e-8, Schur-diagonal order).  Returns the count. */
int orc_poly8_real_roots(const double* c, double* roots_out);
/* all 8 eigenvalues (re, im interleaved) of the companion matrix of c (Schur-diagonal order). */
int orc_poly8_all_roots(const double* c, double* re_im_out);

/* src/estimators/absolute_pose.cc:79-162 P6LEstimator::Estimate on 6 correspondences. */
int orc_p6l_estimate(const double* lines6, const uint8_t* aligned6, const double* points6,
                     double* models_out /* 8 x 12 */);

/* src/base/pose.cc:41-44 RotationMatrixToQuaternion (Eigen::Quaterniond(R)); R col-major. */
void orc_rotation_matrix_to_quaternion(const double* R, double* qvec_wxyz);

/* src/optim/ransac.h:178-278 RANSAC<P6LEstimator, InlierSupportMeasurer, RandomSampler>::Estimate
 * (constructor cap :144-156 included). inlier_mask may be NULL. */
void orc_ransac_p6l(const double* lines, const uint8_t* aligned, const double* points, size_t n,
                    const orc_ransac_options* options, orc_ransac_report* report,
                    uint8_t* inlier_mask);

/* src/estimators/pose.cc:52-94 EstimateAbsolutePoseFromLines. Returns 1 (true) / 0 (false). */
int orc_estimate_absolute_pose_from_lines(const double* lines, const uint8_t* aligned,
                                          const double* points, size_t n,
                                          const orc_ransac_options* options, double* qvec,
                                          double* tvec, uint64_t* num_inliers,
                                          uint8_t* inlier_mask, orc_ransac_report* report);

/* Throughput helper for bench.py's cpu_baseline: runs trials [0, num_trials) of the serial loop
 * (sample -> P6L -> score every model on all n) without the adaptive abort; returns models scored. */
uint64_t orc_ransac_p6l_fixed_trials(const double* lines, const uint8_t* aligned,
                                     const double* points, size_t n, double max_error,
                                     uint64_t num_trials, orc_ransac_report* report);

#ifdef __cplusplus
}
#endif
#endif /* PPSFM_ORACLE_H_ */
