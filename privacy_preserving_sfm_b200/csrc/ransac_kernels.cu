// ransac_kernels.cu — sm_100a kernels of the absolute-pose RANSAC path.
//   * p6l_solve_kernel        one thread per hypothesis (P6L + re3q3)         [A3, A4]
//   * model_offsets_kernel    exclusive scan of per-trial model counts -> compact model ids
//   * score_kernel            (32-model group) x (correspondence segment) tiles; correspondence
//                             tiles are staged in shared memory by the TMA bulk-copy engine
//                             (cp.async.bulk + mbarrier) and broadcast to the 8 warps   [A5, A8]
//   * reduce_parts_kernel     per-model combination of the per-segment partial supports
//   * exact_residual_kernel / seq_support_kernel: index-order (reference-order) supports
//     and inlier mask for the few candidate models that can become "best"
//
// Compiled with --fmad=false: every FP64 operation rounds separately, in the order the reference
// evaluates it (src/estimators/utils.cc:64-88), so residuals, masks and supports are bit-exact.
#include <cfloat>
#include <cstring>
#include <cmath>
#include <cstdint>

#include "p6l_device.cuh"
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
__global__ void __launch_bounds__(64)
p6l_solve_kernel(const double* __restrict__ corr6, const uint8_t* __restrict__ aligned,
                 const uint32_t* __restrict__ samples, int num_trials,
                 double* __restrict__ models_out, int* __restrict__ num_models_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
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

void launch_p6l_solve(const double* corr6, const uint8_t* aligned, const uint32_t* samples,
                      int num_trials, double* models_out, int* num_models_out, cudaStream_t s) {
  if (num_trials <= 0) return;
  const int threads = 64;
  p6l_solve_kernel<<<(num_trials + threads - 1) / threads, threads, 0, s>>>(
      corr6, aligned, samples, num_trials, models_out, num_models_out);
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
// Scoring kernel.  Thread <-> model (12 doubles in registers), warp <-> 32 consecutive compact
// models, block <-> 8 warps sharing one correspondence segment.  Every lane walks the segment in
// index order, so each per-segment partial sum is an index-order sum.
// ------------------------------------------------------------------------------------------
constexpr int kScoreThreads = 256;
constexpr int kTile = 128;    // correspondences per shared-memory tile (6 KB)
constexpr int kStages = 4;

// Inlier COUNT only: residual sums are needed solely to break count-ties and are then computed
// in reference (index) order by the exact kernels below.  The two comparisons are done on the
// bit patterns with integer instructions so that they do not occupy the FP64 pipe:
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
// Filtered evaluation.  The reference arithmetic above costs ~33 FP64-pipe instructions per
// (model, correspondence) pair, a third of them in the IEEE division.  The inlier COUNT only needs
// the sign of |res| - r_max, and for pz > 0
//     |res| <= r_max   <=>   |px l_0 + py l_1 + l_2 pz| <= r_max pz,
// so almost every pair can be decided without any division: projections with FMA chains,
//     d = |fma(l_2, pz, fma(py, l_1, px l_0))| - r_max pz,
// and a rigorous bound on the difference between d / pz and the reference's |res| - r_max:
//     band = k0 + k1 pz,   k0 = 2^-40 (B_x + B_y) Lmax,   k1 = 2^-40 (L2max + r_max + B_z (...)),
// built from per-model sums B_* = sum_k |P_*k| max|X| + |P_*3| and the maxima of the correspondence
// set (2^-40 leaves a factor > 2^9 over all rounding and cancellation errors of both evaluations).
// A pair is decided here only if |pz| >= 2^-30 B_z (which settles the cheirality test
// pz > DBL_EPSILON either way) and, in front of the camera, |d| > band; otherwise — about one pair in 10^8 — the reference arithmetic
// above decides.  Counts are therefore still bit-identical to the reference; the fast path costs
// 15 FP64-pipe instructions and no MUFU.
// ------------------------------------------------------------------------------------------
struct FastConsts {
  int zmin_hi;          // high word of the threshold on |pz| (INT_MAX: never use the fast path)
  double k0, k1;        // band = k0 + k1 pz
  double rmax;
};

struct Corr {  // one correspondence, read once from shared memory for all models of the thread
  double l_0, l_1, l_2, X_0, X_1, X_2;
};
__device__ __forceinline__ Corr load_corr(const double* __restrict__ c) {
  const double2* c2 = reinterpret_cast<const double2*>(c);  // 48-byte records, 16-B aligned
  const double2 v0 = c2[0], v1 = c2[1], v2 = c2[2];
  return Corr{v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
}

// returns false if the pair could not be decided (the caller then runs the reference arithmetic)
__device__ __forceinline__ bool score_fast(const Corr& c, const double (&P)[12],
                                           const FastConsts& fc, unsigned& cnt) {
  const double pz = fma(P[8], c.X_2, fma(P[5], c.X_1, fma(P[2], c.X_0, P[11])));
  const double px = fma(P[6], c.X_2, fma(P[3], c.X_1, fma(P[0], c.X_0, P[9])));
  const double py = fma(P[7], c.X_2, fma(P[4], c.X_1, fma(P[1], c.X_0, P[10])));
  const double num = fma(c.l_2, pz, fma(py, c.l_1, px * c.l_0));
  const double d = fabs(num) - fc.rmax * pz;   // pz < 0: d > 0, "not an inlier", as it must be
  const double band = fma(fc.k1, fabs(pz), fc.k0);
  // The three tests run on the HIGH words with 32-bit integer compares (strict '>' on the high
  // words implies '>' on the doubles; the bounds have a factor 2^9 to spare):
  //   |pz| > zmin (cheirality settled either way),  |d| > band,  |d| finite
  const int pz_hi = __double2hiint(pz) & 0x7fffffff;
  const int d_hi = __double2hiint(d);
  const int ad_hi = d_hi & 0x7fffffff;
  const bool decided = pz_hi > fc.zmin_hi && ad_hi > __double2hiint(band) && ad_hi < 0x7ff00000;
  cnt += decided ? ((unsigned)d_hi >> 31) : 0u;  // d < 0: inlier
  return decided;
}

// per-model constants of the filter
__device__ __forceinline__ FastConsts fast_consts(const double (&P)[12], const double* bounds,
                                                  double r_max, bool live) {
  FastConsts fc;
  const double xmax = bounds[0], lmax = bounds[1], l2max = bounds[2];
  const double bx = (fabs(P[0]) + fabs(P[3]) + fabs(P[6])) * xmax + fabs(P[9]);
  const double by = (fabs(P[1]) + fabs(P[4]) + fabs(P[7])) * xmax + fabs(P[10]);
  const double bz = (fabs(P[2]) + fabs(P[5]) + fabs(P[8])) * xmax + fabs(P[11]);
  // |d_exact - d| <= 8u [(B_x + B_y) Lmax + (L2max + r_max) B_z] (u = 2^-53) for the fused
  // evaluation, and the reference's |res| - r_max times pz differs from d_exact by at most the
  // same plus 8u (|num| + r_max pz) <= 16u (...): everything is below
  //   2^-40 [(B_x + B_y) Lmax + (L2max + r_max) B_z]  +  2^-40 (L2max + r_max) pz
  const double zmin = bz * 0x1p-30;
  fc.k0 = ((bx + by) * lmax + (l2max + r_max) * bz) * 0x1p-40;
  fc.k1 = (l2max + r_max) * 0x1p-40;
  fc.rmax = r_max;
  // usable only for normal, finite constants and a non-negative r_max (a negative one means
  // "nothing is an inlier"); otherwise every pair takes the reference path
  const bool usable = zmin >= 0x1p-900 && zmin < 0x1p900 && fc.k0 < 0x1p900 && fc.k1 < 0x1p900 &&
                      r_max >= 0.0 && r_max < DBL_MAX && live;
  fc.zmin_hi = usable ? __double2hiint(zmin) : 0x7fffffff;
  return fc;
}

// Thread <-> kModelsPerThread models (12 doubles each in registers); every correspondence read
// from shared memory is used for all of them (a broadcast LDS.128 costs four 128-byte wavefronts
// whatever the number of distinct addresses: 12 wavefronts per correspondence and warp).
constexpr int kModelsPerThread = 1;  // 2 halves the shared-memory wavefronts but needs 114
                                     // registers (16 warps / SM): measured slower (3.07 vs 2.74 ms)
constexpr int kModelsPerCta = kScoreThreads * kModelsPerThread;

__global__ void __launch_bounds__(kScoreThreads)
score_kernel(const double* __restrict__ corr6, int n, const double* __restrict__ models,
             const int* __restrict__ offsets, int num_trials, int seg_len, double r_max,
             int kcap, unsigned* __restrict__ part_cnt, const double* __restrict__ bounds) {
  __shared__ __align__(128) double tile[kStages][kTile * 6];
  __shared__ __align__(8) uint64_t full_bar[kStages];

  const int K = offsets[num_trials];
  const int mbase = blockIdx.x * kModelsPerCta;
  if (mbase >= K) return;
  const int seg = blockIdx.y;
  const int i0 = seg * seg_len;
  const int i1 = min(n, i0 + seg_len);

  // Locate (trial, m) of compact model k: largest t with offsets[t] <= k.
  double P[kModelsPerThread][12];
  int kk[kModelsPerThread];
#pragma unroll
  for (int v = 0; v < kModelsPerThread; ++v) {
    const int k = mbase + v * kScoreThreads + threadIdx.x;
    kk[v] = k;
    if (k < K) {
      int lo = 0, hi = num_trials;  // offsets[lo] <= k < offsets[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= k) lo = mid; else hi = mid;
      }
      const double* src = models + (size_t)lo * 96 + (size_t)(k - offsets[lo]) * 12;
#pragma unroll
      for (int j = 0; j < 12; ++j) P[v][j] = src[j];
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) P[v][j] = 0.0;  // px_2 = 0 -> never counted
    }
  }

  const int len = max(0, i1 - i0);
  const int num_tiles = (len + kTile - 1) / kTile;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages && s < num_tiles; ++s) {
      const int cnt_s = min(kTile, len - s * kTile);
      const uint32_t bytes = (uint32_t)cnt_s * 48u;
      mbar_arrive_expect_tx(&full_bar[s], bytes);
      bulk_copy_g2s(&tile[s][0], corr6 + (size_t)(i0 + s * kTile) * 6, bytes, &full_bar[s]);
    }
  }

  unsigned cnt[kModelsPerThread];
  FastConsts fc[kModelsPerThread];
#pragma unroll
  for (int v = 0; v < kModelsPerThread; ++v) {
    cnt[v] = 0;
    fc[v] = fast_consts(P[v], bounds, r_max, kk[v] < K);
  }
  const long long eps_bits = __double_as_longlong(DBL_EPSILON);
  const unsigned long long rmax_bits = (unsigned long long)__double_as_longlong(r_max);
  for (int t = 0; t < num_tiles; ++t) {
    const int s = t % kStages;
    const uint32_t parity = (uint32_t)((t / kStages) & 1);
    mbar_wait(&full_bar[s], parity);
    const int cnt_t = min(kTile, len - t * kTile);
    const double* tp = &tile[s][0];
    int j = 0;
#pragma unroll 1
    constexpr int kGroup = 4 / kModelsPerThread;  // pairs per unrolled group
    for (; j + kGroup <= cnt_t; j += kGroup) {
      unsigned pend = 0;  // undecided (pair, model) combinations of this group (about 1 in 10^8)
#pragma unroll
      for (int u = 0; u < kGroup; ++u) {
        const Corr c = load_corr(tp + (j + u) * 6);
#pragma unroll
        for (int v = 0; v < kModelsPerThread; ++v)
          pend |= score_fast(c, P[v], fc[v], cnt[v]) ? 0u : (1u << (u * kModelsPerThread + v));
      }
      if (pend != 0) {
#pragma unroll
        for (int b = 0; b < kGroup * kModelsPerThread; ++b)  // (static indices: P stays in registers)
          if ((pend >> b) & 1u)
            score_one(tp + (j + b / kModelsPerThread) * 6, P[b % kModelsPerThread], eps_bits,
                      rmax_bits, cnt[b % kModelsPerThread]);
      }
    }
#pragma unroll 1
    for (; j < cnt_t; ++j) {
      const Corr c = load_corr(tp + j * 6);
#pragma unroll
      for (int v = 0; v < kModelsPerThread; ++v)
        if (!score_fast(c, P[v], fc[v], cnt[v]))
          score_one(tp + j * 6, P[v], eps_bits, rmax_bits, cnt[v]);
    }
    __syncthreads();  // everyone is done reading stage s
    if (threadIdx.x == 0 && t + kStages < num_tiles) {
      const int tn = t + kStages;
      const int cnt_n = min(kTile, len - tn * kTile);
      const uint32_t bytes = (uint32_t)cnt_n * 48u;
      fence_proxy_async();
      mbar_arrive_expect_tx(&full_bar[s], bytes);
      bulk_copy_g2s(&tile[s][0], corr6 + (size_t)(i0 + tn * kTile) * 6, bytes, &full_bar[s]);
    }
  }
#pragma unroll
  for (int v = 0; v < kModelsPerThread; ++v)
    if (kk[v] < K) part_cnt[(size_t)seg * kcap + kk[v]] = cnt[v];
}

__global__ void reduce_parts_kernel(const unsigned* __restrict__ part_cnt, int num_segs, int kcap,
                                    const int* __restrict__ offsets, int num_trials,
                                    unsigned* __restrict__ cnt_out) {
  const int K = offsets[num_trials];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  unsigned c = 0;
  for (int g = 0; g < num_segs; ++g) c += part_cnt[(size_t)g * kcap + k];
  cnt_out[k] = c;
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

// ------------------------------------------------------------------------------------------
// Scoring kernel, second form.  Differences from score_kernel above (kept for A/B timing):
//   * the band is a per-model CONSTANT, k0 + k1 B_z + r_max 2^-50 >= k0 + k1 |pz|, and the test on
//     |pz| is gone: a pair is decided iff |d| > band.  "Not an inlier" (d > band) is then right
//     whichever way the reference's cheirality test px_2 > DBL_EPSILON falls, because a failed
//     test also means "not counted"; "inlier" (d < -band) implies r_max pz > band, hence
//     pz > 2^-50 + 2^-39 B_z, and the reference's own px_2 (within 2^-50 B_z of pz) is above
//     DBL_EPSILON = 2^-52.  No overflow test either: the per-model constants are only usable when
//     every intermediate is bounded by 2^940.
//   * 13 FP64-pipe instructions per pair (9 projection FMAs, 3 for the numerator, 1 for d), one
//     LOP + one chained ISETP + one LEA.HI of integer work; the group's predicate is tested once
//     and a failed group is redone pair by pair (filter, then reference arithmetic).
//   * no CTA-wide barrier in the tile loop: every warp counts itself off on a shared-memory
//     counter when it is done with a stage and the last one re-arms the stage's TMA copy, so the
//     warps of a CTA drift apart by up to kStages tiles instead of meeting every 128
//     correspondences in the same phase of the instruction stream.
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

template <int OFF>
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(a), "=d"(b) : "r"(addr), "n"(OFF));
}

// d = |px l_0 + py l_1 + pz l_2| - r_max pz for the correspondence at shared address addr + OFF
template <int OFF>
__device__ __forceinline__ int filter_d_hi(uint32_t addr, const double (&P)[12], double neg_rmax) {
  double l_0, l_1, l_2, X_0, X_1, X_2;
  lds_f64x2<OFF>(addr, l_0, l_1);
  lds_f64x2<OFF + 16>(addr, l_2, X_0);
  lds_f64x2<OFF + 32>(addr, X_1, X_2);
  const double pz = fma(P[8], X_2, fma(P[5], X_1, fma(P[2], X_0, P[11])));
  const double px = fma(P[6], X_2, fma(P[3], X_1, fma(P[0], X_0, P[9])));
  const double py = fma(P[7], X_2, fma(P[4], X_1, fma(P[1], X_0, P[10])));
  const double num = fma(l_2, pz, fma(py, l_1, px * l_0));
  return __double2hiint(fma(neg_rmax, pz, fabs(num)));
}

template <int G, int U = 0>
__device__ __forceinline__ void filter_group(uint32_t addr, const double (&P)[12], double neg_rmax,
                                             int band_hi, bool& ok, unsigned& gcnt) {
  if constexpr (U < G) {
    const int d_hi = filter_d_hi<U * 48>(addr, P, neg_rmax);
    ok = ok && (d_hi & 0x7fffffff) > band_hi;
    gcnt += (unsigned)d_hi >> 31;  // d < 0: inlier
    filter_group<G, U + 1>(addr, P, neg_rmax, band_hi, ok, gcnt);
  }
}

// one pair, filter first and the reference arithmetic if undecided (rare path, not unrolled)
__device__ __forceinline__ void score_checked(uint32_t addr, const double* __restrict__ c,
                                              const double (&P)[12], double neg_rmax, int band_hi,
                                              long long eps_bits, unsigned long long rmax_bits,
                                              unsigned& cnt) {
  const int d_hi = filter_d_hi<0>(addr, P, neg_rmax);
  if ((d_hi & 0x7fffffff) > band_hi) cnt += (unsigned)d_hi >> 31;
  else score_one(c, P, eps_bits, rmax_bits, cnt);
}

constexpr int kScoreWarps = kScoreThreads / 32;

template <int G, int MINB>
__global__ void __launch_bounds__(kScoreThreads, MINB)
score_kernel_v2(const double* __restrict__ corr6, int n, const double* __restrict__ models,
                const int* __restrict__ offsets, int num_trials, int seg_len, double r_max,
                int kcap, unsigned* __restrict__ part_cnt, const double* __restrict__ bounds) {
  __shared__ __align__(128) double tile[kStages][kTile * 6];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ unsigned done[kStages];   // warps that have finished with the stage

  const int K = offsets[num_trials];
  const int mbase = blockIdx.x * kScoreThreads;
  if (mbase >= K) return;
  const int seg = blockIdx.y;
  const int i0 = seg * seg_len;
  const int i1 = min(n, i0 + seg_len);

  double P[12];
  const int k = mbase + threadIdx.x;
  if (k < K) {
    int lo = 0, hi = num_trials;  // offsets[lo] <= k < offsets[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (offsets[mid] <= k) lo = mid; else hi = mid;
    }
    const double* src = models + (size_t)lo * 96 + (size_t)(k - offsets[lo]) * 12;
#pragma unroll
    for (int j = 0; j < 12; ++j) P[j] = src[j];
  } else {
#pragma unroll
    for (int j = 0; j < 12; ++j) P[j] = 0.0;
  }

  const int len = max(0, i1 - i0);
  const int num_tiles = (len + kTile - 1) / kTile;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      done[s] = 0;
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages && s < num_tiles; ++s) {
      const uint32_t bytes = (uint32_t)min(kTile, len - s * kTile) * 48u;
      mbar_arrive_expect_tx(&full_bar[s], bytes);
      bulk_copy_g2s(&tile[s][0], corr6 + (size_t)(i0 + s * kTile) * 6, bytes, &full_bar[s]);
    }
  }

  const int band_hi = filter_band_hi(P, bounds, r_max, k < K);
  const double neg_rmax = -r_max;
  const long long eps_bits = __double_as_longlong(DBL_EPSILON);
  const unsigned long long rmax_bits = (unsigned long long)__double_as_longlong(r_max);
  const uint32_t tile_addr = smem_u32(&tile[0][0]);
  unsigned cnt = 0;
  for (int t = 0; t < num_tiles; ++t) {
    const int s = t % kStages;
    mbar_wait(&full_bar[s], (uint32_t)((t / kStages) & 1));
    const int cnt_t = min(kTile, len - t * kTile);
    const double* tp = &tile[s][0];
    const uint32_t ta = tile_addr + (uint32_t)s * (kTile * 48);
    int j = 0;
#pragma unroll 1
    for (; j + G <= cnt_t; j += G) {
      bool ok = true;
      unsigned gcnt = 0;
      filter_group<G>(ta + (uint32_t)j * 48u, P, neg_rmax, band_hi, ok, gcnt);
      if (ok) {
        cnt += gcnt;
      } else {  // some pair of the group is undecided (about one pair in 10^8)
#pragma unroll 1
        for (int u = 0; u < G; ++u)
          score_checked(ta + (uint32_t)(j + u) * 48u, tp + (j + u) * 6, P, neg_rmax, band_hi,
                        eps_bits, rmax_bits, cnt);
      }
    }
#pragma unroll 1
    for (; j < cnt_t; ++j)
      score_checked(ta + (uint32_t)j * 48u, tp + j * 6, P, neg_rmax, band_hi, eps_bits, rmax_bits,
                    cnt);
    // this warp is done with stage s; the last warp of the CTA to get here refills it
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      __threadfence_block();
      if (atomicAdd(&done[s], 1u) == kScoreWarps - 1) {
        done[s] = 0;
        const int tn = t + kStages;
        if (tn < num_tiles) {
          const uint32_t bytes = (uint32_t)min(kTile, len - tn * kTile) * 48u;
          __threadfence_block();
          fence_proxy_async();
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          bulk_copy_g2s(&tile[s][0], corr6 + (size_t)(i0 + tn * kTile) * 6, bytes, &full_bar[s]);
        }
      }
    }
  }
  if (k < K) part_cnt[(size_t)seg * kcap + k] = cnt;
}

// ------------------------------------------------------------------------------------------
// Scoring kernel, third form: a FLOAT first stage in front of the FP64 filter.
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
// range — goes through score_slow: the FP64 filter and then the reference arithmetic, on the
// double correspondences read from global memory (L2-resident), with the model reloaded from
// global memory so that the hot loop carries only the float model.
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

template <int G, int U = 0>
__device__ __forceinline__ void filter32_group(uint32_t addr, const float2 (&Pf)[12], float2 nr,
                                               int band, bool& ok, unsigned& gcnt) {
  if constexpr (U < G) {
    const float4 v0 = lds_f32x4<U * 48>(addr), v1 = lds_f32x4<U * 48 + 16>(addr),
                 v2 = lds_f32x4<U * 48 + 32>(addr);
    const float2 l_0 = make_float2(v0.x, v0.y), l_1 = make_float2(v0.z, v0.w);
    const float2 l_2 = make_float2(v1.x, v1.y), X_0 = make_float2(v1.z, v1.w);
    const float2 X_1 = make_float2(v2.x, v2.y), X_2 = make_float2(v2.z, v2.w);
    const float2 pz = __ffma2_rn(Pf[8], X_2, __ffma2_rn(Pf[5], X_1, __ffma2_rn(Pf[2], X_0, Pf[11])));
    const float2 px = __ffma2_rn(Pf[6], X_2, __ffma2_rn(Pf[3], X_1, __ffma2_rn(Pf[0], X_0, Pf[9])));
    const float2 py = __ffma2_rn(Pf[7], X_2, __ffma2_rn(Pf[4], X_1, __ffma2_rn(Pf[1], X_0, Pf[10])));
    float2 num = __ffma2_rn(l_2, pz, __ffma2_rn(py, l_1, __fmul2_rn(px, l_0)));
    num.x = fabsf(num.x);
    num.y = fabsf(num.y);
    const float2 d = __ffma2_rn(nr, pz, num);
    const int da = __float_as_int(d.x), db = __float_as_int(d.y);
    ok = ok && (da & 0x7fffffff) > band && (db & 0x7fffffff) > band;
    gcnt += ((unsigned)da >> 31) + ((unsigned)db >> 31);  // d < 0: inlier
    filter32_group<G, U + 1>(addr, Pf, nr, band, ok, gcnt);
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

template <int G, int MINB>
__global__ void __launch_bounds__(kScoreThreads, MINB)
score_kernel_v3(const double* __restrict__ corr6, const float* __restrict__ corr6f, int n,
                const double* __restrict__ models, const int* __restrict__ offsets,
                int num_trials, int seg_len, double r_max, int kcap,
                unsigned* __restrict__ part_cnt, const double* __restrict__ bounds) {
  __shared__ __align__(128) float tile[kStages][kTileR * 12];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ unsigned done[kStages];  // warps that have finished with the stage

  const int K = offsets[num_trials];
  const int mbase = blockIdx.x * kScoreThreads;
  if (mbase >= K) return;
  const int seg = blockIdx.y;
  const int i0 = seg * seg_len;  // even: seg_len is a multiple of 128
  const int i1 = min(n, i0 + seg_len);
  const int len = max(0, i1 - i0);
  const int nrec = len >> 1;  // full records; an odd last correspondence goes through score_slow
  const int num_tiles = (nrec + kTileR - 1) / kTileR;
  const float* recs = corr6f + (size_t)(i0 >> 1) * 12;

  const int k = mbase + threadIdx.x;
  const double* src = nullptr;
  float2 Pf[12];
  int band = -1;
  {
    double P[12];
    if (k < K) {
      int lo = 0, hi = num_trials;  // offsets[lo] <= k < offsets[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= k) lo = mid; else hi = mid;
      }
      src = models + (size_t)lo * 96 + (size_t)(k - offsets[lo]) * 12;
#pragma unroll
      for (int j = 0; j < 12; ++j) P[j] = src[j];
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) P[j] = 0.0;
    }
    band = filter32_band_bits(P, bounds, r_max, k < K);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const float f = __double2float_rn(P[j]);
      Pf[j] = make_float2(f, f);
    }
  }
  const float nrf = -__double2float_rn(r_max);
  const float2 nr = make_float2(nrf, nrf);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      done[s] = 0;
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages && s < num_tiles; ++s) {
      const uint32_t bytes = (uint32_t)min(kTileR, nrec - s * kTileR) * 48u;
      mbar_arrive_expect_tx(&full_bar[s], bytes);
      bulk_copy_g2s(&tile[s][0], recs + (size_t)s * kTileR * 12, bytes, &full_bar[s]);
    }
  }

  const uint32_t tile_addr = smem_u32(&tile[0][0]);
  unsigned cnt = 0;
  for (int t = 0; t < num_tiles; ++t) {
    const int s = t % kStages;
    mbar_wait(&full_bar[s], (uint32_t)((t / kStages) & 1));
    const int cnt_t = min(kTileR, nrec - t * kTileR);
    const uint32_t ta = tile_addr + (uint32_t)s * (kTileR * 48);
    const double* gc = corr6 + (size_t)(i0 + 2 * t * kTileR) * 6;  // the tile's doubles (slow path)
    int j = 0;
#pragma unroll 1
    for (; j + G <= cnt_t; j += G) {
      bool ok = true;
      unsigned gcnt = 0;
      filter32_group<G>(ta + (uint32_t)j * 48u, Pf, nr, band, ok, gcnt);
      if (ok) cnt += gcnt;
      else cnt += score_slow(gc + (size_t)j * 12, 2 * G, src, bounds, r_max);
    }
    if (j < cnt_t && src != nullptr)  // ragged end of the set
      cnt += score_slow(gc + (size_t)j * 12, 2 * (cnt_t - j), src, bounds, r_max);
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
  if (k < K) {
    if (len & 1) cnt += score_slow(corr6 + (size_t)(i1 - 1) * 6, 1, src, bounds, r_max);
    part_cnt[(size_t)seg * kcap + k] = cnt;
  }
}

void launch_score(const double* corr6, const float* corr6f, const double* bounds, int n,
                  const double* models, const int* offsets, int num_trials, int num_segs,
                  int seg_len, double max_residual, int kcap, unsigned* part_cnt,
                  unsigned* cnt_out, cudaStream_t s) {
  if (num_trials <= 0) return;
  const double r_max = inlier_abs_threshold(max_residual);
  dim3 grid((kcap + kModelsPerCta - 1) / kModelsPerCta, num_segs);
#define PPSFM_SCORE_V2(G, MINB)                                                                \
  score_kernel_v2<G, MINB><<<grid, kScoreThreads, 0, s>>>(corr6, n, models, offsets, num_trials, \
                                                          seg_len, r_max, kcap, part_cnt, bounds)
#define PPSFM_SCORE_V3(G, MINB)                                                              \
  score_kernel_v3<G, MINB><<<grid, kScoreThreads, 0, s>>>(corr6, corr6f, n, models, offsets, \
                                                          num_trials, seg_len, r_max, kcap,  \
                                                          part_cnt, bounds)
  switch (tune_int("PPSFM_SCORE_VARIANT", 13)) {
    case 0:
      score_kernel<<<grid, kScoreThreads, 0, s>>>(corr6, n, models, offsets, num_trials, seg_len,
                                                  r_max, kcap, part_cnt, bounds);
      break;
    case 2: PPSFM_SCORE_V2(8, 2); break;
    case 4: PPSFM_SCORE_V2(8, 3); break;
    case 5: PPSFM_SCORE_V2(4, 4); break;
    case 10: PPSFM_SCORE_V3(2, 2); break;
    case 11: PPSFM_SCORE_V3(4, 2); break;
    case 12: PPSFM_SCORE_V3(2, 3); break;
    case 13: PPSFM_SCORE_V3(4, 3); break;
    case 14: PPSFM_SCORE_V3(2, 4); break;
    case 15: PPSFM_SCORE_V3(4, 4); break;
    case 16: PPSFM_SCORE_V3(8, 2); break;
    case 17: PPSFM_SCORE_V3(1, 4); break;
    default: PPSFM_SCORE_V2(4, 2); break;
  }
#undef PPSFM_SCORE_V2
#undef PPSFM_SCORE_V3
  reduce_parts_kernel<<<(kcap + 255) / 256, 256, 0, s>>>(part_cnt, num_segs, kcap, offsets,
                                                         num_trials, cnt_out);
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
  for (int e = 0; e < num_e; ++e) {
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

void launch_exact(const double* corr6, int n, const double* emodels, int num_e,
                  double max_residual, double* rbuf, uint8_t* mask, unsigned long long* ecnt,
                  double* esum, cudaStream_t s) {
  if (num_e <= 0 || n <= 0) return;
  exact_residual_kernel<<<(n + 255) / 256, 256, 0, s>>>(corr6, n, emodels, num_e, max_residual,
                                                        rbuf, mask);
  if (ecnt != nullptr) {
    seq_support_kernel<<<num_e, kSeqThreads, 0, s>>>(rbuf, n, num_e, max_residual, ecnt, esum);
  }
}

}  // namespace ppsfm
