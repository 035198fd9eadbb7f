// ransac_kernels.cu — sm_100a kernels of the absolute-pose RANSAC path.
//   * p6l_solve_kernel        one thread per hypothesis (P6L + re3q3)         [A3, A4]
//   * model_offsets_kernel    exclusive scan of per-trial model counts -> compact model ids
//   * score_kernel            (256-model block) x (correspondence segment) tiles; inlier counts
//                             through a float filter -> FP64 filter -> reference arithmetic
//                             cascade; correspondence tiles are staged in shared memory by the
//                             TMA bulk-copy engine (cp.async.bulk + mbarrier)          [A5, A8]
//   * reduce_parts_kernel     per-model combination of the per-segment partial supports
//   * exact_residual_kernel / seq_support_kernel: index-order (reference-order) supports
//     and inlier mask for the few candidate models that can become "best"
//
// Compiled with --fmad=false: every FP64 operation rounds separately, in the order the reference
// evaluates it (src/estimators/utils.cc:64-88), so residuals, masks and supports are bit-exact.
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <cmath>
#include <cstdint>

#include "p6l_octet.cuh"
#include "ransac_kernels.h"
#include "common.h"

namespace ppsfm {

// ------------------------------------------------------------------------------------------
// Correspondence packing: (lines n x 3, points n x 3) -> corr6 n x 6 (48-byte records).
// ------------------------------------------------------------------------------------------
__global__ void pack_corr_kernel(const double* __restrict__ lines,
                                 const double* __restrict__ points, size_t n,
                                 double* __restrict__ corr6, float* __restrict__ corr6f,
                                 double* __restrict__ bounds) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  unsigned long long bx = 0, bl = 0, b2 = 0;  // bit patterns of |X|, of |l_0| / |l_1|, of |l_2|
  if (i < n * 3) {
    const size_t r = i / 3, c = i % 3;
    const double l = lines[i], x = points[i];
    corr6[r * 6 + c] = l;
    corr6[r * 6 + 3 + c] = x;
    // float copy for the first stage of the score filter: record r / 2, slot r % 2
    float* f = corr6f + (r >> 1) * 12 + (r & 1);
    f[2 * c] = __double2float_rn(l);
    f[6 + 2 * c] = __double2float_rn(x);
    bx = (unsigned long long)__double_as_longlong(fabs(x));
    if (c < 2) bl = (unsigned long long)__double_as_longlong(fabs(l));
    else b2 = (unsigned long long)__double_as_longlong(fabs(l));
  }
  // bounds[0] = max |X_k|, [1] = max(|l_0|, |l_1|), [2] = max |l_2| over the set (non-negative doubles order
  // like their bit patterns; a NaN input yields a NaN bound, which disables the fast path)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ox = __shfl_xor_sync(0xffffffffu, bx, o);
    const unsigned long long ol = __shfl_xor_sync(0xffffffffu, bl, o);
    const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, b2, o);
    bx = ox > bx ? ox : bx;
    bl = ol > bl ? ol : bl;
    b2 = o2 > b2 ? o2 : b2;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(bounds), bx);
    atomicMax(reinterpret_cast<unsigned long long*>(bounds) + 1, bl);
    atomicMax(reinterpret_cast<unsigned long long*>(bounds) + 2, b2);
  }
}

void launch_pack_corr(const double* lines, const double* points, size_t n, double* corr6,
                      float* corr6f, double* bounds, cudaStream_t s) {
  const int threads = 256;
  const size_t total = n * 3;
  if (n == 0) return;
  cudaMemsetAsync(bounds, 0, 3 * sizeof(double), s);
  if (n & 1) cudaMemsetAsync(corr6f + (n >> 1) * 12, 0, 12 * sizeof(float), s);  // empty slot
  pack_corr_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(lines, points,
                                                                                    n, corr6, corr6f,
                                                                                    bounds);
}

// ------------------------------------------------------------------------------------------
// P6L solve: one thread per hypothesis.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
p6l_solve_kernel(const double* __restrict__ corr6, const uint8_t* __restrict__ aligned,
                 const uint32_t* __restrict__ samples, int num_trials,
                 double* __restrict__ models_out, int* __restrict__ num_models_out,
                 int lanes_per_warp) {
  // Only the first `lanes_per_warp` lanes of a warp take a hypothesis: the solver's data-dependent
  // loops (QR iterations, root polishing) diverge between lanes, and the kernel is latency-bound
  // with the SMs nearly empty, so spreading the hypotheses over more warps shortens every warp's
  // instruction stream at no cost.
  const int lane = threadIdx.x & 31;
  if (lane >= lanes_per_warp) return;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int t = warp * lanes_per_warp + lane;
  if (t >= num_trials) return;
  double lines[6][3], points[6][3];
  bool all_aligned = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const uint32_t idx = samples[6 * (size_t)t + i];
    const double* c = corr6 + 6 * (size_t)idx;
    lines[i][0] = c[0]; lines[i][1] = c[1]; lines[i][2] = c[2];
    points[i][0] = c[3]; points[i][1] = c[4]; points[i][2] = c[5];
    all_aligned = all_aligned && (aligned != nullptr && aligned[idx] != 0);
  }
  double models[8][12];
  const int n = dev::p6l_estimate(lines, all_aligned, points, models);
  num_models_out[t] = n;
  double* out = models_out + (size_t)t * 96;
  for (int m = 0; m < n; ++m)
    for (int j = 0; j < 12; ++j) out[m * 12 + j] = models[m][j];
}

// ------------------------------------------------------------------------------------------
// P6L solve, eight lanes per hypothesis (p6l_octet.cuh): the latency-critical variant for the
// head of a call.  Bit-identical models.
// ------------------------------------------------------------------------------------------
constexpr int kOctetThreads = 64;
__global__ void __launch_bounds__(kOctetThreads)
p6l_solve_octet_kernel(const double* __restrict__ corr6, const uint8_t* __restrict__ aligned,
                       const uint32_t* __restrict__ samples, int num_trials,
                       double* __restrict__ models_out, int* __restrict__ num_models_out,
                       int octets_per_warp) {
  __shared__ double Tsm[kOctetThreads / 8][8 * dev::Octet::kLd];
  const int oct = threadIdx.x >> 3;
  // octets of a warp follow different control paths; with fewer than four octets per warp the
  // other lanes of the warp stay idle (octets_per_warp: measured trade-off, see launch)
  const int oct_in_warp = oct & 3;
  if (oct_in_warp >= octets_per_warp) return;
  const int warp = (blockIdx.x * kOctetThreads + threadIdx.x) >> 5;
  const int t = warp * octets_per_warp + oct_in_warp;
  if (t >= num_trials) return;  // whole octets leave
  dev::Octet o;
  o.T = Tsm[oct];
  o.sub = threadIdx.x & 7;
  o.mask = 0xffu << ((threadIdx.x & 31) & ~7);
  double lines[6][3], points[6][3];
  bool all_aligned = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const uint32_t idx = samples[6 * (size_t)t + i];
    const double* c = corr6 + 6 * (size_t)idx;
    lines[i][0] = c[0]; lines[i][1] = c[1]; lines[i][2] = c[2];
    points[i][0] = c[3]; points[i][1] = c[4]; points[i][2] = c[5];
    all_aligned = all_aligned && (aligned != nullptr && aligned[idx] != 0);
  }
  const int n = dev::p6l_estimate_octet(o, lines, all_aligned, points,
                                        models_out + (size_t)t * 96);
  if (o.sub == 0) num_models_out[t] = n;
}

void launch_p6l_solve(const double* corr6, const uint8_t* aligned, const uint32_t* samples,
                      int num_trials, double* models_out, int* num_models_out, cudaStream_t s,
                      int lanes_per_warp, int threads_per_cta) {
  if (num_trials <= 0) return;
  if (lanes_per_warp == kSolveOctet) {
    static const int opw = std::max(1, std::min(4, tune_int("PPSFM_OCTETS_PER_WARP", 4)));
    const int per_cta = (kOctetThreads / 32) * opw;
    p6l_solve_octet_kernel<<<(num_trials + per_cta - 1) / per_cta, kOctetThreads, 0, s>>>(
        corr6, aligned, samples, num_trials, models_out, num_models_out, opw);
    return;
  }
  const int threads = std::max(32, std::min(256, threads_per_cta)) / 32 * 32;
  const int lanes = std::max(1, std::min(32, lanes_per_warp));
  const int warps = (num_trials + lanes - 1) / lanes;
  p6l_solve_kernel<<<(warps * 32 + threads - 1) / threads, threads, 0, s>>>(
      corr6, aligned, samples, num_trials, models_out, num_models_out, lanes);
}

// ------------------------------------------------------------------------------------------
// Exclusive scan of model counts (single block; H <= a few 10^5).
// offsets[t] = first compact id of trial t; offsets[H] = K.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
model_offsets_kernel(const int* __restrict__ num_models, int num_trials,
                     int* __restrict__ offsets) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < num_trials; base += 1024) {
    const int t = base + tid;
    const int v = (t < num_trials) ? num_models[t] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int incl = x + (warp > 0 ? warp_sums[warp - 1] : 0) + carry;
    if (t < num_trials) offsets[t] = incl - v;
    __syncthreads();
    if (tid == 1023) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) offsets[num_trials] = carry_s;
}

void launch_model_offsets(const int* num_models, int num_trials, int* offsets, cudaStream_t s) {
  model_offsets_kernel<<<1, 1024, 0, s>>>(num_models, num_trials, offsets);
}

// ------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier helpers (sm_90+/sm_100a PTX).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// Scoring kernel.  Thread <-> M models (k, k + 256, ...), warp <-> 32 consecutive compact models
// per slot, block <-> 8 warps sharing one correspondence segment; grid = model blocks x segments,
// per-segment counts are combined by reduce_parts_kernel.
//
// Inlier COUNT only: residual sums are needed solely to break count-ties and are then computed in
// reference (index) order by the exact kernels below.  The count of a (model, correspondence)
// pair is decided by a cascade of three evaluations, each exact about what it decides:
//   1. float stage   d_32 = |px l_0 + py l_1 + pz l_2| - r_max pz on float copies, packed FFMA2;
//                    decides iff |d_32| > band_32                     (all but ~1e-5 of the pairs)
//   2. FP64 stage    the same d with FMA chains in double; decides iff |d| > band_64
//   3. reference     src/estimators/utils.cc:64-88 operation by operation (score_one)
// Stage 3 is the definition of the result; stages 1 and 2 only ever answer where their error
// bounds prove that stage 3 would answer the same, so the counts are bit-identical to the
// reference's.  For pz > 0:  |res| <= r_max  <=>  |px l_0 + py l_1 + l_2 pz| <= r_max pz, which
// needs no division; the bands bound the distance between d and the reference's own
// (|res| - r_max) pz, rounding and cancellation included (derivations at filter_band_hi and
// filter32_band_bits).
// ------------------------------------------------------------------------------------------
constexpr int kScoreThreads = 256;
constexpr int kScoreWarps = kScoreThreads / 32;
constexpr int kStages = 4;

// Stage 3.  The two comparisons are done on the bit patterns with integer instructions:
//   px_2 > DBL_EPSILON        <=>  (int64)bits(px_2) > (int64)bits(eps)     (NaN: see below)
//   res * res <= max_residual  <=>  bits(|res|) <= bits(r_max)  with r_max the largest double whose
//                                  rounded square is <= max_residual (rounding is monotone)
// A NaN px_2 passes the first test only to produce a NaN residual, which fails the second: the
// pair is not counted, exactly as with the floating-point comparisons of the reference.
// With max_residual >= DBL_MAX (RANSACOptions::max_error = inf) the reference also counts the
// pairs that fail the cheirality test (their residual is DBL_MAX, utils.cc:85-86); r_max is
// DBL_MAX exactly then (inlier_abs_threshold), which flags the case.
constexpr unsigned long long kAllInliers = 0x7fefffffffffffffull;  // bits(DBL_MAX)

__device__ __forceinline__ void score_one(const double* __restrict__ c, const double (&P)[12],
                                          const long long eps_bits,
                                          const unsigned long long rmax_bits, unsigned& cnt) {
  // src/estimators/utils.cc:64-88 — same operations, same order, no contraction.
  const double2* c2 = reinterpret_cast<const double2*>(c);  // 48-byte records, 16-B aligned
  const double2 v0 = c2[0], v1 = c2[1], v2 = c2[2];
  const double l_0 = v0.x, l_1 = v0.y, l_2 = v1.x;
  const double X_0 = v1.y, X_1 = v2.x, X_2 = v2.y;
  const double px_2 = P[2] * X_0 + P[5] * X_1 + P[8] * X_2 + P[11];
  if (rmax_bits == kAllInliers && !(px_2 > DBL_EPSILON)) {
    cnt += 1;  // residual DBL_MAX <= max_residual
    return;
  }
  if (__double_as_longlong(px_2) > eps_bits) {
    const double px_0 = P[0] * X_0 + P[3] * X_1 + P[6] * X_2 + P[9];
    const double px_1 = P[1] * X_0 + P[4] * X_1 + P[7] * X_2 + P[10];
    const double inv_px_2 = 1.0 / px_2;
    const double res = px_0 * l_0 * inv_px_2 + px_1 * l_1 * inv_px_2 + l_2;
    const unsigned long long ares =
        (unsigned long long)__double_as_longlong(res) & 0x7fffffffffffffffull;
    // UNSIGNED compare: the compiler may take |res| through the FP64 pipe, which turns a NaN into
    // the canonical negative one — it must still fail.  A negative r_max (NaN / negative
    // max_residual: nothing is an inlier) has the sign bit set and is tested for separately.
    if (ares <= rmax_bits && (long long)rmax_bits >= 0) cnt += 1;  // support_measurement.cc:42-45
  }
}

// ------------------------------------------------------------------------------------------
// Stage 2: FP64 filter.  Projections and numerator with FMA chains,
//     d = fma(-r_max, pz, |fma(l_2, pz, fma(py, l_1, px l_0))|),
// 13 FP64-pipe instructions and no division (the reference arithmetic costs ~33, a third of them
// in the IEEE division).  Per model, with B_* = sum_k |P_*k| max|X| + |P_*3| and the maxima of the
// correspondence set (pack_corr_kernel):
//     S = (B_x + B_y) Lmax + (L2max + r_max) B_z  >=  |num| + r_max |pz|,
//     band_64 = 2^-40 (S + (L2max + r_max) B_z) + 2^-50 r_max.
// The fused d and the reference's own (|res| - r_max) px_2 each carry at most ~16 roundings of
// relative size 2^-53 on terms bounded by S, so they differ by less than 2^-48 S: the 2^-40 leaves
// a factor > 2^7.  A pair is decided iff |d| > band_64 (compared on the high words):
//   * "not an inlier" (d > band) is right whichever way the reference's cheirality test
//     px_2 > DBL_EPSILON falls, because a failed test also means "not counted";
//   * "inlier" (d < -band) implies r_max pz > band, hence pz > 2^-50 + 2^-39 B_z, and the
//     reference's own px_2 (within 2^-50 B_z of pz) is above DBL_EPSILON = 2^-52.
// No overflow test: the constants are only usable when every intermediate is below 2^940, and
// r_max = DBL_MAX ("count everything", see kAllInliers) or a negative / NaN r_max disable it.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int filter_band_hi(const double (&P)[12], const double* bounds,
                                              double r_max, bool live) {
  if (!live) return -1;  // P = 0: d = +0 is "decided", its sign bit is 0 — never counted, never slow
  const double xmax = bounds[0], lmax = bounds[1], l2max = bounds[2];
  const double bx = (fabs(P[0]) + fabs(P[3]) + fabs(P[6])) * xmax + fabs(P[9]);
  const double by = (fabs(P[1]) + fabs(P[4]) + fabs(P[7])) * xmax + fabs(P[10]);
  const double bz = (fabs(P[2]) + fabs(P[5]) + fabs(P[8])) * xmax + fabs(P[11]);
  const double scale = (bx + by) * lmax + (l2max + r_max) * bz;   // >= |num| + r_max |pz|
  const double band = (scale + (l2max + r_max) * bz) * 0x1p-40 + r_max * 0x1p-50;
  const bool usable = band >= 0x1p-900 && band < 0x1p900 && r_max >= 0.0 && r_max < DBL_MAX;
  return usable ? __double2hiint(band) : 0x7fffffff;
}

// ------------------------------------------------------------------------------------------
// Stage 1: float filter, and the kernel.
//
// The sign of d = |px l_0 + py l_1 + pz l_2| - r_max pz is all the count needs, and for all but a
// few pairs in a million it is already certain in single precision.  The first stage evaluates d
// on float copies of the model and of the correspondences, two correspondences per lane with the
// packed FFMA2 of sm_100 (13 FMAs per pair -> 6.5 FFMA2; the FP64 pipe runs a DFMA warp
// instruction every other cycle, FFMA2 does two FMAs per lane at that rate), and decides the pair
// iff |d_32| > band_32:
//   inputs rounded to float (relative 2^-24 each), three FMA chains of depth 3, the numerator
//   chain and the final FMA give  |d_32 - d| <= 10 * 2^-24 * S,  S = (B_x + B_y) Lmax +
//   (L2max + r_max) B_z >= |num| + r_max |pz|;  band_32 = 2^-20 S + band_64 + 2^-60 leaves a factor
//   1.6 on that, adds the FP64 band (distance between the exact d and the reference's own
//   evaluation, see filter_band_hi) and covers float underflow: the stage is used only when all
//   inputs are below 2^40 in magnitude, so that an absolute error of 2^-150 (input or
//   intermediate flushed / rounded in the subnormal range) reaches d with a factor < 2^80.
// Everything else — a group in which some pair is undecided, models or sets outside the float
// range — goes through score_slow: stages 2 and 3 on the double correspondences read from global
// memory (L2-resident), with the model reloaded from global memory so that the hot loop carries
// only the float model.
//
// Correspondence records (48 B, two correspondences) are staged in shared memory by the TMA bulk
// copy engine (cp.async.bulk + mbarrier), kStages tiles deep.  There is no CTA-wide barrier in the
// tile loop: every warp counts itself off on a shared-memory counter when it is done with a stage
// and the last one re-arms the stage's copy, so the warps of a CTA drift apart by up to kStages
// tiles instead of meeting every tile in the same phase of the instruction stream.
// Measured on the bench workload (1.9e9 pairs): reference arithmetic with TMA tiles 3.79 ms,
// + FP64 filter 2.03 ms, + float stage 1.11 ms (profiles/r01_s3_score_variants_*.txt).
// ------------------------------------------------------------------------------------------
constexpr int kTileR = 128;  // float records (= 256 correspondences, 6 KB) per stage

__device__ __forceinline__ int filter32_band_bits(const double (&P)[12], const double* bounds,
                                                  double r_max, bool live) {
  if (!live) return -1;  // P = 0: d = +0, "decided" (band -1), sign bit 0
  const double xmax = bounds[0], lmax = bounds[1], l2max = bounds[2];
  const double bx = (fabs(P[0]) + fabs(P[3]) + fabs(P[6])) * xmax + fabs(P[9]);
  const double by = (fabs(P[1]) + fabs(P[4]) + fabs(P[7])) * xmax + fabs(P[10]);
  const double bz = (fabs(P[2]) + fabs(P[5]) + fabs(P[8])) * xmax + fabs(P[11]);
  double pmax = 0.0;
#pragma unroll
  for (int j = 0; j < 12; ++j) pmax = fmax(pmax, fabs(P[j]));
  const double scale = (bx + by) * lmax + (l2max + r_max) * bz;
  const double band64 = (scale + (l2max + r_max) * bz) * 0x1p-40 + r_max * 0x1p-50;
  const double band = scale * 0x1p-20 + band64 + 0x1p-60;
  // (comparisons are false for NaN; fmax drops a NaN entry of P but then B_* is NaN)
  const bool usable = xmax < 0x1p40 && lmax < 0x1p40 && l2max < 0x1p40 && pmax < 0x1p40 &&
                      r_max >= 0.0 && r_max < 0x1p40 && scale < 0x1p80 && band < 0x1p80;
  return usable ? __float_as_int(__double2float_ru(band)) : 0x7fffffff;
}

template <int OFF>
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr), "n"(OFF));
  return v;
}

template <int G, int M, int U = 0>
__device__ __forceinline__ void filter32_group(uint32_t addr, const float2 (&Pf)[M][12], float2 nr,
                                               const int (&band)[M], bool (&ok)[M],
                                               unsigned (&gcnt)[M]) {
  if constexpr (U < G) {
    const float4 v0 = lds_f32x4<U * 48>(addr), v1 = lds_f32x4<U * 48 + 16>(addr),
                 v2 = lds_f32x4<U * 48 + 32>(addr);
    const float2 l_0 = make_float2(v0.x, v0.y), l_1 = make_float2(v0.z, v0.w);
    const float2 l_2 = make_float2(v1.x, v1.y), X_0 = make_float2(v1.z, v1.w);
    const float2 X_1 = make_float2(v2.x, v2.y), X_2 = make_float2(v2.z, v2.w);
#pragma unroll
    for (int m = 0; m < M; ++m) {  // the record is read once for the thread's M models
      const float2(&P)[12] = Pf[m];
      const float2 pz = __ffma2_rn(P[8], X_2, __ffma2_rn(P[5], X_1, __ffma2_rn(P[2], X_0, P[11])));
      const float2 px = __ffma2_rn(P[6], X_2, __ffma2_rn(P[3], X_1, __ffma2_rn(P[0], X_0, P[9])));
      const float2 py = __ffma2_rn(P[7], X_2, __ffma2_rn(P[4], X_1, __ffma2_rn(P[1], X_0, P[10])));
      float2 num = __ffma2_rn(l_2, pz, __ffma2_rn(py, l_1, __fmul2_rn(px, l_0)));
      num.x = fabsf(num.x);
      num.y = fabsf(num.y);
      const float2 d = __ffma2_rn(nr, pz, num);
      const int da = __float_as_int(d.x), db = __float_as_int(d.y);
      ok[m] = ok[m] && (da & 0x7fffffff) > band[m] && (db & 0x7fffffff) > band[m];
      gcnt[m] += ((unsigned)da >> 31) + ((unsigned)db >> 31);  // d < 0: inlier
    }
    filter32_group<G, M, U + 1>(addr, Pf, nr, band, ok, gcnt);
  }
}

// `count` correspondences from global memory against the model at Psrc: FP64 filter, then the
// reference arithmetic.  Not inlined: nothing of the caller's state is passed by reference.
__device__ __noinline__ unsigned score_slow(const double* __restrict__ c, int count,
                                            const double* __restrict__ Psrc,
                                            const double* __restrict__ bounds, double r_max) {
  double P[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) P[j] = Psrc[j];
  const int band_hi = filter_band_hi(P, bounds, r_max, true);
  const double neg_rmax = -r_max;
  const long long eps_bits = __double_as_longlong(DBL_EPSILON);
  const unsigned long long rmax_bits = (unsigned long long)__double_as_longlong(r_max);
  unsigned cnt = 0;
#pragma unroll 1
  for (int u = 0; u < count; ++u, c += 6) {
    const double l_0 = c[0], l_1 = c[1], l_2 = c[2], X_0 = c[3], X_1 = c[4], X_2 = c[5];
    const double pz = fma(P[8], X_2, fma(P[5], X_1, fma(P[2], X_0, P[11])));
    const double px = fma(P[6], X_2, fma(P[3], X_1, fma(P[0], X_0, P[9])));
    const double py = fma(P[7], X_2, fma(P[4], X_1, fma(P[1], X_0, P[10])));
    const double num = fma(l_2, pz, fma(py, l_1, px * l_0));
    const int d_hi = __double2hiint(fma(neg_rmax, pz, fabs(num)));
    if ((d_hi & 0x7fffffff) > band_hi) cnt += (unsigned)d_hi >> 31;
    else score_one(c, P, eps_bits, rmax_bits, cnt);
  }
  return cnt;
}

template <int G, int M, int MAXREG>
__global__ void __maxnreg__(MAXREG)
score_kernel(const double* __restrict__ corr6, const float* __restrict__ corr6f, int n,
             const double* __restrict__ models, const int* __restrict__ offsets, int num_trials,
             int seg_len, double r_max, int kcap, unsigned* __restrict__ part_cnt,
             const double* __restrict__ bounds, const int* __restrict__ list,
             const int* __restrict__ list_count, int shard_world, int shard_rank) {
  // list != nullptr: second phase of a pruned wave — slot i of the grid scores model list[i]
  // shard_world > 1 (first phase only): ONE call sharded over the GPUs of a communicator; this
  // rank scores the model blocks b with b % shard_world == shard_rank (K is only known on the
  // device, so the interleaving is what balances the ranks)
  __shared__ __align__(128) float tile[kStages][kTileR * 12];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ unsigned done[kStages];  // warps that have finished with the stage

  const int K = list ? *list_count : offsets[num_trials];
  if ((int)(blockIdx.x * shard_world + shard_rank) * (kScoreThreads * M) >= K) return;
  const int seg = blockIdx.y;
  const int i0 = seg * seg_len;  // even: seg_len is a multiple of 128
  const int i1 = min(n, i0 + seg_len);
  const int len = max(0, i1 - i0);
  const int nrec = len >> 1;  // full records; an odd last correspondence goes through score_slow
  const int num_tiles = (nrec + kTileR - 1) / kTileR;
  const float* recs = corr6f + (size_t)(i0 >> 1) * 12;
  const float nrf = -__double2float_rn(r_max);
  const float2 nr = make_float2(nrf, nrf);
  const uint32_t tile_addr = smem_u32(&tile[0][0]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      done[s] = 0;
    }
    mbar_fence_init();
  }
  unsigned phase_bits = 0;  // bit s: parity of the next completion of full_bar[s]

  // A CTA takes the model blocks blockIdx.x, blockIdx.x + gridDim.x, ...: one block per CTA in
  // the first phase (grid sized to the capacity), a short grid walking the survivor list in the
  // second phase of a pruned wave.  The barriers live across blocks; phase_bits carries on.
  for (int mbase = (blockIdx.x * shard_world + shard_rank) * (kScoreThreads * M); mbase < K;
       mbase += gridDim.x * shard_world * (kScoreThreads * M)) {
  // thread <-> models mbase + m * 256 + tid, m < M
  const double* src[M];
  float2 Pf[M][12];
  int band[M];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int slot = mbase + m * kScoreThreads + threadIdx.x;
    const int k = slot < K ? (list ? list[slot] : slot) : -1;
    double P[12];
    src[m] = nullptr;
    if (k >= 0) {
      int lo = 0, hi = num_trials;  // offsets[lo] <= k < offsets[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= k) lo = mid; else hi = mid;
      }
      src[m] = models + (size_t)lo * 96 + (size_t)(k - offsets[lo]) * 12;
#pragma unroll
      for (int j = 0; j < 12; ++j) P[j] = src[m][j];
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) P[j] = 0.0;
    }
    band[m] = filter32_band_bits(P, bounds, r_max, k >= 0);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const float f = __double2float_rn(P[j]);
      Pf[m][j] = make_float2(f, f);
    }
  }
  __syncthreads();  // barriers initialised / every warp is done with the previous block's tiles
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages && s < num_tiles; ++s) {
      const uint32_t bytes = (uint32_t)min(kTileR, nrec - s * kTileR) * 48u;
      mbar_arrive_expect_tx(&full_bar[s], bytes);
      bulk_copy_g2s(&tile[s][0], recs + (size_t)s * kTileR * 12, bytes, &full_bar[s]);
    }
  }

  unsigned cnt[M];
#pragma unroll
  for (int m = 0; m < M; ++m) cnt[m] = 0;
  for (int t = 0; t < num_tiles; ++t) {
    const int s = t % kStages;
    mbar_wait(&full_bar[s], (phase_bits >> s) & 1u);
    phase_bits ^= 1u << s;
    const int cnt_t = min(kTileR, nrec - t * kTileR);
    const uint32_t ta = tile_addr + (uint32_t)s * (kTileR * 48);
    const double* gc = corr6 + (size_t)(i0 + 2 * t * kTileR) * 6;  // the tile's doubles (slow path)
    int j = 0;
#pragma unroll 1
    for (; j + G <= cnt_t; j += G) {
      bool ok[M];
      unsigned gcnt[M];
#pragma unroll
      for (int m = 0; m < M; ++m) {
        ok[m] = true;
        gcnt[m] = 0;
      }
      filter32_group<G, M>(ta + (uint32_t)j * 48u, Pf, nr, band, ok, gcnt);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        if (ok[m]) cnt[m] += gcnt[m];
        else cnt[m] += score_slow(gc + (size_t)j * 12, 2 * G, src[m], bounds, r_max);
      }
    }
    if (j < cnt_t) {  // ragged end of the set
#pragma unroll
      for (int m = 0; m < M; ++m)
        if (src[m] != nullptr)
          cnt[m] += score_slow(gc + (size_t)j * 12, 2 * (cnt_t - j), src[m], bounds, r_max);
    }
    // this warp is done with stage s; the last warp of the CTA to get here refills it
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      __threadfence_block();
      if (atomicAdd(&done[s], 1u) == kScoreWarps - 1) {
        done[s] = 0;
        const int tn = t + kStages;
        if (tn < num_tiles) {
          const uint32_t bytes = (uint32_t)min(kTileR, nrec - tn * kTileR) * 48u;
          __threadfence_block();
          fence_proxy_async();
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          bulk_copy_g2s(&tile[s][0], recs + (size_t)tn * kTileR * 12, bytes, &full_bar[s]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) {
    if (src[m] != nullptr) {
      if (len & 1) cnt[m] += score_slow(corr6 + (size_t)(i1 - 1) * 6, 1, src[m], bounds, r_max);
      part_cnt[(size_t)seg * kcap + mbase + m * kScoreThreads + threadIdx.x] = cnt[m];
    }
  }
  }  // model blocks
}

// Per-model combination of the per-segment counts.  best_lb (optional): running maximum of the
// FINAL counts over everything scored so far in the call (see launch_score).
__global__ void reduce_parts_kernel(const unsigned* __restrict__ part_cnt, int num_segs, int kcap,
                                    const int* __restrict__ offsets, int num_trials,
                                    unsigned* __restrict__ cnt_out,
                                    unsigned* __restrict__ best_lb, int shard_world,
                                    int shard_rank, int models_per_block) {
  const int K = offsets[num_trials];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kcap) return;
  // sharded call: the counts of the other ranks' models arrive through the all-reduce (sum), so
  // this rank contributes zeros for them — and for the unused slots above K
  const bool mine = k < K && (k / models_per_block) % shard_world == shard_rank;
  if (!mine) {
    if (shard_world > 1) cnt_out[k] = 0;
    return;
  }
  unsigned c = 0;
  for (int g = 0; g < num_segs; ++g) c += part_cnt[(size_t)g * kcap + k];
  cnt_out[k] = c;
  if (best_lb) atomicMax(best_lb, c);
}

// Sharded call, after the all-reduce of the counts: the pruning bound of the following waves is
// the best count over ALL ranks' models.
__global__ void raise_best_lb_kernel(const unsigned* __restrict__ cnt, const int* __restrict__ offsets,
                                     int num_trials, unsigned* __restrict__ best_lb) {
  const int K = offsets[num_trials];
  unsigned m = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x)
    m = max(m, cnt[k]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(best_lb, m);
}

// Second phase of a pruned wave: slot i holds the remaining count of model list[i].
__global__ void reduce_parts_list_kernel(const unsigned* __restrict__ part_cnt, int num_segs,
                                         int kcap, const int* __restrict__ list,
                                         const int* __restrict__ list_count,
                                         unsigned* __restrict__ cnt_out,
                                         unsigned* __restrict__ best_lb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *list_count) return;
  unsigned c = 0;
  for (int g = 0; g < num_segs; ++g) c += part_cnt[(size_t)g * kcap + i];
  const int k = list[i];
  c += cnt_out[k];
  cnt_out[k] = c;
  atomicMax(best_lb, c);
}

// Models that can still reach the best count known so far: first-phase count + everything left.
__global__ void survivors_kernel(const unsigned* __restrict__ cnt_first, int remaining,
                                 const int* __restrict__ offsets, int num_trials,
                                 const unsigned* __restrict__ best_lb, int* __restrict__ list,
                                 int* __restrict__ list_count, int shard_world, int shard_rank,
                                 int models_per_block) {
  const int K = offsets[num_trials];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool keep = k < K && (k / models_per_block) % shard_world == shard_rank &&
                    cnt_first[k] + (unsigned)remaining >= *best_lb;  // '>=': ties count
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (bal == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(list_count, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (keep) list[base + __popc(bal & ((1u << lane) - 1u))] = k;
}

// Largest double r with fl(r * r) <= max_residual (host).  IEEE multiplication is monotone, so
// the predicate is monotone in r and a bisection over the bit patterns of the non-negative doubles
// finds the boundary in 63 steps — also where r * r underflows (a step-by-step search from
// sqrt(max_residual) would walk through all denormals for max_residual = 0).
double inlier_abs_threshold(double max_residual) {
  if (!(max_residual >= 0.0)) return -1.0;  // nothing is an inlier (bits compare fails)
  auto ok = [&](unsigned long long bits) {
    double r;
    std::memcpy(&r, &bits, sizeof(r));
    return r * r <= max_residual;
  };
  unsigned long long lo = 0, hi = 0x7fefffffffffffffull;  // +0 .. DBL_MAX; ok(lo) always holds
  if (ok(hi)) return DBL_MAX;
  while (hi - lo > 1) {
    const unsigned long long mid = lo + (hi - lo) / 2;
    if (ok(mid)) lo = mid; else hi = mid;
  }
  double r;
  std::memcpy(&r, &lo, sizeof(r));
  return r;
}

// Exact pruning (prune.best_lb != nullptr).  The replay on the host only ever looks at a model's
// count to ask whether it beats or ties the best count so far (InlierSupportMeasurer::Compare);
// a model that cannot reach the best count of the waves scored BEFORE its own can therefore be
// dropped without changing anything the reference computes.  best_lb is that count, maintained
// on the device (atomicMax over the final counts of every launch, in stream order), so it is
// valid even though the host has not consumed those waves yet.  With prune.n_first > 0 the wave
// is scored in two phases: all models on correspondences [0, n_first), then only the models with
// first-phase count + (n - n_first) >= best_lb on the rest; dropped models keep their first-phase
// count, which is below best_lb like their true count.
void launch_score(const double* corr6, const float* corr6f, const double* bounds, int n,
                  const double* models, const int* offsets, int num_trials, int num_segs,
                  int seg_len, double max_residual, int kcap, unsigned* part_cnt,
                  unsigned* cnt_out, cudaStream_t s, const ScorePrune& prune,
                  const ScoreShard& shard) {
  if (num_trials <= 0) return;
  const int W = shard.world > 1 ? shard.world : 1, R = shard.world > 1 ? shard.rank : 0;
  const double r_max = inlier_abs_threshold(max_residual);
  // 4 records (8 correspondences) per unrolled group, 2 models per thread (every record read
  // from shared memory serves both), 2 CTAs per SM (100 registers): the best of the (group,
  // models per thread, occupancy) variants measured on the bench workload — (4,2,2) 1.085 ms,
  // (4,1,3) 1.126, (2,2,2) 1.131, (2,2,3) 1.254, (1,3,2) 1.341, (1,2,3) 1.363, (1,2,4) 2.180.
  constexpr int kG = 4, kM = kScoreModelsPerCta / kScoreThreads, kMinB = kScoreMaxRegs;
  const int xblocks = (kcap + kScoreThreads * kM - 1) / (kScoreThreads * kM);
  const bool two_phase = prune.best_lb && prune.n_first > 0 && prune.n_first < n;
  const int n1 = two_phase ? prune.n_first : n;  // multiple of 128 when two_phase (caller)
  score_kernel<kG, kM, kMinB><<<dim3((xblocks + W - 1) / W, num_segs), kScoreThreads, 0, s>>>(
      corr6, corr6f, n1, models, offsets, num_trials, seg_len, r_max, kcap, part_cnt, bounds,
      nullptr, nullptr, W, R);
  reduce_parts_kernel<<<(kcap + 255) / 256, 256, 0, s>>>(
      part_cnt, num_segs, kcap, offsets, num_trials, cnt_out, two_phase ? nullptr : prune.best_lb,
      W, R, kScoreThreads * kM);
  if (!two_phase) return;
  cudaMemsetAsync(prune.list_count, 0, sizeof(int), s);
  survivors_kernel<<<(kcap + 255) / 256, 256, 0, s>>>(cnt_out, n - n1, offsets, num_trials,
                                                      prune.best_lb, prune.list, prune.list_count,
                                                      W, R, kScoreThreads * kM);
  // (few models survive: a short grid walks the list instead of one mostly empty CTA per block)
  score_kernel<kG, kM, kMinB><<<dim3(std::min(xblocks, 16), prune.num_segs2), kScoreThreads, 0, s>>>(
      corr6 + (size_t)n1 * 6, corr6f + (size_t)(n1 / 2) * 12, n - n1, models, offsets, num_trials,
      prune.seg_len2, r_max, kcap, part_cnt, bounds, prune.list, prune.list_count, 1, 0);
  reduce_parts_list_kernel<<<(kcap + 255) / 256, 256, 0, s>>>(
      part_cnt, prune.num_segs2, kcap, prune.list, prune.list_count, cnt_out, prune.best_lb);
}

void launch_raise_best_lb(const unsigned* cnt, const int* offsets, int num_trials,
                          unsigned* best_lb, cudaStream_t s) {
  raise_best_lb_kernel<<<64, 256, 0, s>>>(cnt, offsets, num_trials, best_lb);
}

// ------------------------------------------------------------------------------------------
// Exact (reference-order) support for a handful of models.
// ------------------------------------------------------------------------------------------
__global__ void exact_residual_kernel(const double* __restrict__ corr6, int n,
                                      const double* __restrict__ emodels, int num_e,
                                      double max_residual, double* __restrict__ rbuf,
                                      uint8_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* c = corr6 + (size_t)i * 6;
  const double l_0 = c[0], l_1 = c[1], l_2 = c[2];
  const double X_0 = c[3], X_1 = c[4], X_2 = c[5];
  // blockIdx.y strides over the models (few correspondences x many tied candidates — the
  // mapper's registration calls — would otherwise run on a couple of CTAs)
  for (int e = blockIdx.y; e < num_e; e += gridDim.y) {
    const double* P = emodels + (size_t)e * 12;
    const double px_2 = P[2] * X_0 + P[5] * X_1 + P[8] * X_2 + P[11];
    double r2;
    if (px_2 > DBL_EPSILON) {
      const double px_0 = P[0] * X_0 + P[3] * X_1 + P[6] * X_2 + P[9];
      const double px_1 = P[1] * X_0 + P[4] * X_1 + P[7] * X_2 + P[10];
      const double inv_px_2 = 1.0 / px_2;
      const double res = px_0 * l_0 * inv_px_2 + px_1 * l_1 * inv_px_2 + l_2;
      r2 = res * res;
    } else {
      r2 = DBL_MAX;
    }
    rbuf[(size_t)e * n + i] = r2;
    if (mask != nullptr) mask[(size_t)e * n + i] = (r2 <= max_residual) ? 1 : 0;
  }
}

// One block per model: index-order sum of the inlier residuals (support_measurement.cc:42-47).
// 256 threads compact the inlier residuals of a 2048-wide chunk into shared memory in index
// order (ballot + popc prefix), then one thread adds them sequentially — the additions must
// round exactly like the reference's serial loop, so only the compaction is parallel.
constexpr int kSeqThreads = 256;
constexpr int kSeqChunk = 2048;
__global__ void __launch_bounds__(kSeqThreads)
seq_support_kernel(const double* __restrict__ rbuf, int n, int num_e, double max_residual,
                   unsigned long long* __restrict__ ecnt, double* __restrict__ esum) {
  __shared__ double buf[kSeqChunk];
  __shared__ int warp_cnt[kSeqThreads / 32];
  const int e = blockIdx.x;
  if (e >= num_e) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* r = rbuf + (size_t)e * n;
  unsigned long long cnt = 0;
  double sum = 0.0;
  for (int base = 0; base < n; base += kSeqChunk) {
    double v[8];
    unsigned bal[8];
    int wtotal = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i = base + warp * 256 + it * 32 + lane;
      v[it] = (i < n) ? r[i] : DBL_MAX;
      const bool inl = (i < n) && (v[it] <= max_residual);
      bal[it] = __ballot_sync(0xffffffffu, inl);
      wtotal += __popc(bal[it]);
    }
    if (lane == 0) warp_cnt[warp] = wtotal;
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kSeqThreads / 32; ++w) {
      const int c = warp_cnt[w];
      if (w < warp) off += c;
      total += c;
    }
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if ((bal[it] >> lane) & 1u) buf[off + __popc(bal[it] & lt)] = v[it];
      off += __popc(bal[it]);
    }
    __syncthreads();
    if (tid == 0) {
      int i = 0;
      for (; i + 8 <= total; i += 8) {
        const double a0 = buf[i], a1 = buf[i + 1], a2 = buf[i + 2], a3 = buf[i + 3];
        const double a4 = buf[i + 4], a5 = buf[i + 5], a6 = buf[i + 6], a7 = buf[i + 7];
        sum += a0; sum += a1; sum += a2; sum += a3;
        sum += a4; sum += a5; sum += a6; sum += a7;
      }
      for (; i < total; ++i) sum += buf[i];
      cnt += (unsigned long long)total;
    }
    __syncthreads();
  }
  if (tid == 0) {
    ecnt[e] = cnt;
    esum[e] = sum;
  }
}

// dst[e] = the 12 doubles at src + off[e] (candidate models of a wave -> a dense batch)
__global__ void gather_models_kernel(const double* __restrict__ src,
                                     const long long* __restrict__ off, int ne,
                                     double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 12 * ne) return;
  const int e = i / 12, j = i - 12 * e;
  dst[i] = src[off[e] + j];
}
void launch_gather_models(const double* src, const long long* off, int ne, double* dst,
                          cudaStream_t s) {
  if (ne > 0) gather_models_kernel<<<(12 * ne + 255) / 256, 256, 0, s>>>(src, off, ne, dst);
}

void launch_exact(const double* corr6, int n, const double* emodels, int num_e,
                  double max_residual, double* rbuf, uint8_t* mask, unsigned long long* ecnt,
                  double* esum, cudaStream_t s) {
  if (num_e <= 0 || n <= 0) return;
  const int bx = (n + 255) / 256;
  int by = 1;  // ~ 4 CTAs per SM in total
  while (by < num_e && bx * by < 592) by *= 2;
  exact_residual_kernel<<<dim3(bx, by), 256, 0, s>>>(corr6, n, emodels, num_e, max_residual,
                                                     rbuf, mask);
  if (ecnt != nullptr) {
    seq_support_kernel<<<num_e, kSeqThreads, 0, s>>>(rbuf, n, num_e, max_residual, ecnt, esum);
  }
}

}  // namespace ppsfm
