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

namespace ppsfm {

// ------------------------------------------------------------------------------------------
// Correspondence packing: (lines n x 3, points n x 3) -> corr6 n x 6 (48-byte records).
// ------------------------------------------------------------------------------------------
__global__ void pack_corr_kernel(const double* __restrict__ lines,
                                 const double* __restrict__ points, size_t n,
                                 double* __restrict__ corr6, double* __restrict__ bounds) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  unsigned long long bx = 0, bl = 0, b2 = 0;  // bit patterns of |X|, of |l_0| / |l_1|, of |l_2|
  if (i < n * 3) {
    const size_t r = i / 3, c = i % 3;
    const double l = lines[i], x = points[i];
    corr6[r * 6 + c] = l;
    corr6[r * 6 + 3 + c] = x;
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
                      double* bounds, cudaStream_t s) {
  const int threads = 256;
  const size_t total = n * 3;
  cudaMemsetAsync(bounds, 0, 3 * sizeof(double), s);
  pack_corr_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(lines, points,
                                                                                    n, corr6, bounds);
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
    if (ares <= rmax_bits) cnt += 1;  // src/optim/support_measurement.cc:42-45
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

void launch_score(const double* corr6, const double* bounds, int n, const double* models,
                  const int* offsets, int num_trials, int num_segs, int seg_len,
                  double max_residual, int kcap, unsigned* part_cnt, unsigned* cnt_out,
                  cudaStream_t s) {
  if (num_trials <= 0) return;
  const double r_max = inlier_abs_threshold(max_residual);
  dim3 grid((kcap + kModelsPerCta - 1) / kModelsPerCta, num_segs);
  score_kernel<<<grid, kScoreThreads, 0, s>>>(corr6, n, models, offsets, num_trials, seg_len,
                                              r_max, kcap, part_cnt, bounds);
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
