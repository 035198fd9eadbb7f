// ba_schur.cu — reduced camera system S = blockdiag(U + D_c) - W (V + D_p)^-1 W^T without atomics.
//
// (Inside the reference this is ceres' SchurEliminator, reached from ceres::Solve at
// src/optim/bundle_adjustment.cc:306.)
//
// With V_p + D_p = L L^T per point and Z_e = (J_c,e^T J_p,e) L^-T (6x3) per observation, the block
// of S for the camera pair (i, j) is  - sum over points seen by both of  Z_e Z_f^T.  The sparsity
// structure never changes during a solve, so it is turned ONCE into gather lists: every
// (observation e, observation f) pair of one point whose camera blocks satisfy i >= j, sorted by
// (i, j) and then by e (CUB radix sort — structure set-up, not the per-iteration path).  Per LM
// iteration:
//   ba_point_damp_kernel   thread / point        L^-1, (V + D)^-1, h = L^-1 g_p
//   ba_zbuild_kernel       thread / observation  192-byte records [Z_e | Z_e h] (coalesced through
//                                                shared memory)
//   ba_schur_gather_kernel warp / chunk of <= 128 list entries of one camera pair; ONE FP64
//                          tensor-core MMA (mma.sync.m8n8k4.f64) per entry accumulates
//                          [Z_e | z_e] (6x4, padded to 8x4) x [Z_f | 1{e=f}]^T into the warp's 8x8
//                          accumulator fragment: the two 192-byte records are one coalesced load
//                          each, the kernel needs ~40 registers, and the sums land directly in the
//                          lanes that store them (no read-modify-write, no memset of the 72 MB
//                          matrix, bit-reproducible)
//   ba_schur_multi_kernel  only for camera pairs that span several chunks (few cameras, many
//                          points): ordered sum of the chunk partials
// The former one-warp-per-point kernel issued 36 FP64 atomics per camera pair (396 M per
// iteration at 500 cameras / 2 M observations) and was bound by the SMs' RED issue rate.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cfloat>
#include <cstdint>
#include <initializer_list>

#include "ba_kernels.h"
#include "common.h"

namespace ppsfm {

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 128;  // list entries per warp task
constexpr int kRec = 24;     // doubles per observation record: rows [Z_a0 Z_a1 Z_a2 z_a], a = 0..5

__device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

// ------------------------------------------------------------------------------------------
// structure set-up
// ------------------------------------------------------------------------------------------
// entries contributed by one point: ordered observation pairs (e, f) with block(e) > block(f),
// or block(e) == block(f) (both orders and e == f, so that diagonal blocks come out symmetric)
__global__ void sch_count_kernel(BaDev d, int64_t* __restrict__ counts) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.P) return;
  int64_t n = 0;
  if (d.pt_var[p]) {
    const int64_t k0 = d.pt_start[p], k1 = d.pt_start[p + 1];
    for (int64_t e = k0; e < k1; ++e) {
      const int bi = d.cam_block[d.obs_cam[e]];
      if (bi < 0) continue;
      for (int64_t f = k0; f < k1; ++f) {
        const int bj = d.cam_block[d.obs_cam[f]];
        if (bj >= 0 && bi >= bj) ++n;
      }
    }
  }
  counts[p] = n;
}

__global__ void sch_emit_kernel(BaDev d, const int64_t* __restrict__ offs,
                                uint64_t* __restrict__ keys, int* __restrict__ vals) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.P || !d.pt_var[p]) return;
  int64_t o = offs[p];
  const int64_t k0 = d.pt_start[p], k1 = d.pt_start[p + 1];
  for (int64_t e = k0; e < k1; ++e) {
    const int bi = d.cam_block[d.obs_cam[e]];
    if (bi < 0) continue;
    for (int64_t f = k0; f < k1; ++f) {
      const int bj = d.cam_block[d.obs_cam[f]];
      if (bj < 0 || bi < bj) continue;
      keys[o] = ((uint64_t)(uint32_t)(tri(bi) + bj) << 32) | (uint32_t)e;
      vals[o] = (int)f;
      ++o;
    }
  }
}

// pair_start[q] = first sorted entry whose pair id is >= q (binary search), q in [0, npairs]
__global__ void sch_pair_start_kernel(const uint64_t* __restrict__ keys, int64_t nent, int npairs,
                                      int64_t* __restrict__ pair_start,
                                      int* __restrict__ pair_nchunks) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > npairs) return;
  auto lower = [&](uint32_t pk) {
    int64_t lo = 0, hi = nent;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((uint32_t)(keys[mid] >> 32) < pk) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const int64_t a = lower((uint32_t)q);
  pair_start[q] = a;
  if (q < npairs) {
    const int64_t b = lower((uint32_t)q + 1u);
    const int64_t cnt = b - a;
    pair_nchunks[q] = cnt == 0 ? 1 : (int)((cnt + kChunk - 1) / kChunk);  // empty pairs write zeros
  }
}

__global__ void sch_fill_chunks_kernel(int npairs, const int* __restrict__ pair_chunk,
                                       int* __restrict__ chunk_pair, int* __restrict__ multi,
                                       int* __restrict__ n_multi) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npairs) return;
  const int c0 = pair_chunk[q], c1 = pair_chunk[q + 1];
  for (int c = c0; c < c1; ++c) chunk_pair[c] = q;
  if (c1 - c0 > 1 && multi != nullptr) multi[atomicAdd(n_multi, 1)] = q;
}

__global__ void sch_extract_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ vals,
                                   int64_t nent, int2* __restrict__ ent) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nent) ent[i] = make_int2((int)(uint32_t)keys[i], vals[i]);
}

// ------------------------------------------------------------------------------------------
// per LM iteration
// ------------------------------------------------------------------------------------------
// Damped point block V + clamp(diag V) / radius = L L^T; M = L^-1 (lower), Vinv = M^T M,
// h = M g_p.  Lv[p] = {m00, m10, m11, m20, m21, m22, h0, h1, h2}.
__global__ void __launch_bounds__(kThreads)
ba_point_damp_kernel(BaDev d, double radius, double min_diag, double max_diag) {
  const int p = blockIdx.x * kThreads + threadIdx.x;
  const int P = d.P;
  if (p >= P) return;
  double m00 = 0, m10 = 0, m11 = 0, m20 = 0, m21 = 0, m22 = 0;
  if (d.pt_var[p]) {
    double a = d.V[p], b = d.V[P + p], c = d.V[2 * P + p];
    double e = d.V[3 * P + p], f = d.V[4 * P + p], i = d.V[5 * P + p];
    a += fmin(fmax(a, min_diag), max_diag) / radius;
    e += fmin(fmax(e, min_diag), max_diag) / radius;
    i += fmin(fmax(i, min_diag), max_diag) / radius;
    if (a > 0.0) {
      const double r00 = rsqrt(a);
      const double l10 = b * r00, l20 = c * r00;
      const double d1 = e - l10 * l10;
      if (d1 > 0.0) {
        const double r11 = rsqrt(d1);
        const double l21 = (f - l20 * l10) * r11;
        const double d2 = i - l20 * l20 - l21 * l21;
        if (d2 > 0.0) {
          const double r22 = rsqrt(d2);
          m00 = r00; m11 = r11; m22 = r22;
          m10 = -m11 * l10 * m00;
          m21 = -m22 * l21 * m11;
          m20 = -(m21 * l10 + m22 * l20) * m00;
        }
      }
    }
  }
  // Vinv = M^T M (symmetric: 00 01 02 11 12 22)
  d.Vinv[p] = m00 * m00 + m10 * m10 + m20 * m20;
  d.Vinv[P + p] = m10 * m11 + m20 * m21;
  d.Vinv[2 * P + p] = m20 * m22;
  d.Vinv[3 * P + p] = m11 * m11 + m21 * m21;
  d.Vinv[4 * P + p] = m21 * m22;
  d.Vinv[5 * P + p] = m22 * m22;
  const double g0 = d.gp[p], g1 = d.gp[P + p], g2 = d.gp[2 * P + p];
  double* lv = d.Lv + 9 * (size_t)p;
  lv[0] = m00; lv[1] = m10; lv[2] = m11; lv[3] = m20; lv[4] = m21; lv[5] = m22;
  lv[6] = m00 * g0;
  lv[7] = m10 * g0 + m11 * g1;
  lv[8] = m20 * g0 + m21 * g1 + m22 * g2;
}

// One thread per observation: record rows [Z_a | z_a] (6 x 4, the A operand of the gather MMA).  The 192-byte records of a warp are contiguous
// (6 KB); they are transposed through shared memory so that the warp stores full 512-byte rows.
constexpr int kZThreads = 128;
__global__ void __launch_bounds__(kZThreads) ba_zbuild_kernel(BaDev d) {
  __shared__ double stage[kZThreads / 32][32 * (kRec + 1)];
  const int64_t K = d.K;
  const int64_t k = (int64_t)blockIdx.x * kZThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double rec[kRec];
#pragma unroll
  for (int i = 0; i < kRec; ++i) rec[i] = 0.0;
  if (k < K) {
    const int p = d.obs_pt[k];
    const double* lv = d.Lv + 9 * (size_t)p;
    const double m00 = lv[0], m10 = lv[1], m11 = lv[2], m20 = lv[3], m21 = lv[4], m22 = lv[5];
    const double h0 = lv[6], h1 = lv[7], h2 = lv[8];
    double jp[2][3], jc[2][6];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) jp[r][c] = d.J[ba_jidx(14 + 3 * r + c, k)];
#pragma unroll
      for (int c = 0; c < 6; ++c) jc[r][c] = d.J[ba_jidx(2 + 6 * r + c, k)];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const double w0 = jc[0][a] * jp[0][0] + jc[1][a] * jp[1][0];
      const double w1 = jc[0][a] * jp[0][1] + jc[1][a] * jp[1][1];
      const double w2 = jc[0][a] * jp[0][2] + jc[1][a] * jp[1][2];
      const double z0 = w0 * m00;                          // Z = W M^T
      const double z1 = w0 * m10 + w1 * m11;
      const double z2 = w0 * m20 + w1 * m21 + w2 * m22;
      rec[4 * a] = z0; rec[4 * a + 1] = z1; rec[4 * a + 2] = z2;
      rec[4 * a + 3] = z0 * h0 + z1 * h1 + z2 * h2;
    }
  }
  double* st = stage[warp];
#pragma unroll
  for (int i = 0; i < kRec; ++i) st[lane * (kRec + 1) + i] = rec[i];
  __syncwarp();
  const int64_t kw = (int64_t)blockIdx.x * kZThreads + warp * 32;  // first observation of the warp
  double* out = d.Zrec + (size_t)kw * kRec;
  const int64_t nvalid = (K - kw < 32 ? (K - kw) : 32) * kRec;
#pragma unroll
  for (int i = 0; i < kRec; ++i) {
    const int idx = i * 32 + lane;  // linear index into the warp's 32 x 24 doubles
    if (idx < nvalid) out[idx] = st[(idx / kRec) * (kRec + 1) + (idx % kRec)];
  }
}

// value `idx` (0..35 = S block element a*6+c, 36..41 = rhs element) of camera pair (bi, bj):
// S = [diag: U + D] - sum,  rhs = -g_c + sum.  Fixed tangent dimensions get identity rows.
__device__ __forceinline__ void schur_store(const BaDev& d, int bi, int bj, int idx, double sum,
                                            double radius, double min_diag, double max_diag,
                                            int include_cam) {
  if (idx >= 42) return;
  const unsigned mask_i = d.cam_mask[d.block_img[bi]];
  if (idx >= 36) {
    if (bi != bj) return;
    const int a = idx - 36;
    const bool on = (mask_i >> a) & 1u;
    double v = sum;
    if (include_cam) v -= d.gc[6 * (size_t)bi + a];
    d.S[(size_t)d.n * d.ld + 6 * bi + a] = on ? v : 0.0;
    return;
  }
  const int a = idx / 6, c = idx - 6 * a;
  double v = -sum;
  if (bi == bj) {
    const bool on_a = (mask_i >> a) & 1u, on_c = (mask_i >> c) & 1u;
    if (include_cam) {
      const double u = d.U[36 * (size_t)bi + idx];
      v += u;
      if (a == c) v += fmin(fmax(u, min_diag), max_diag) / radius;
    }
    if (!on_a || !on_c) v = (a == c && include_cam) ? 1.0 : 0.0;
  }
  d.S[(size_t)(6 * bi + a) * d.ld + 6 * bj + c] = v;
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int kGThreads = 256;
constexpr int kUnroll = 4;  // entries whose record loads are in flight together, per warp

// Five CTAs (40 warps) per SM with four entries per batch (40 registers): the kernel is bound by
// the latency of its scattered 192-byte record reads, and WARPS in flight are what hides it — at
// config 4: 16 warps x 8 entries 0.90 ms, 24 x 8 0.70 ms, 40 x 4 0.62 ms, 64 x 2 0.61 ms (spills),
// 32 x 8 (spills) 0.70 ms; 16 entries per batch at 16 warps gains 4 %, a second accumulator
// chain nothing, staging the records through shared memory with cp.async (3 stages x 8 entries
// per warp in flight, 16 warps) nothing (profiles/r02_gather_variants.txt).
__global__ void __launch_bounds__(kGThreads, 5)
ba_schur_gather_kernel(BaDev d, double radius, double min_diag, double max_diag,
                       int include_cam) {
  const int w = (int)(((int64_t)blockIdx.x * kGThreads + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= d.sch_nchunks) return;  // whole warps leave: the collectives below stay converged
  const int pair = d.sch_chunk_pair[w];
  const int c0 = d.sch_pair_chunk[pair], c1 = d.sch_pair_chunk[pair + 1];
  const int64_t begin = d.sch_pair_start[pair] + (int64_t)(w - c0) * kChunk;
  int64_t end = d.sch_pair_start[pair + 1];
  if (end > begin + kChunk) end = begin + kChunk;
  int bi = (int)((sqrtf(8.0f * (float)pair + 1.0f) - 1.0f) * 0.5f);
  while (tri(bi + 1) <= pair) ++bi;
  while (tri(bi) > pair) --bi;
  const int bj = pair - tri(bi);
  // MMA fragment coordinates: A[m = g][k = q], B[k = q][n = g], C[m = g][n = 2q, 2q+1]
  const int g = lane >> 2, q = lane & 3;
  const bool a_lane = lane < 24;            // rows 0..5 of [Z | z]: record element `lane`
  const bool b_lane = lane < 24 && q != 3;  // Z_f[n = g][k = q]
  const bool one_lane = lane == 27;         // B[k = 3][n = 6] = 1 for e == f: column 6 sums z_e
  double acc0 = 0.0, acc1 = 0.0;
  const int n = (int)(end - begin);
  const int lc = lane < 24 ? lane : 23;  // record element this lane loads (clamped: loads are
                                         // unconditional so that they can all be in flight)
  const double* __restrict__ Z = d.Zrec;
  int2 ef_next = make_int2(0, 0);
  if (lane < n) ef_next = d.sch_ent[begin + lane];
  for (int base = 0; base < n; base += 32) {
    const int2 ef = ef_next;
    if (base + 32 + lane < n) ef_next = d.sch_ent[begin + base + 32 + lane];
    const int m = min(32, n - base);
    for (int j0 = 0; j0 < m; j0 += kUnroll) {
      int ke[kUnroll], kf[kUnroll];
      double av[kUnroll], bv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {  // (dead slots read entry (0, 0): valid addresses)
        ke[u] = __shfl_sync(0xffffffffu, ef.x, (j0 + u) & 31);
        kf[u] = __shfl_sync(0xffffffffu, ef.y, (j0 + u) & 31);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        av[u] = Z[(size_t)ke[u] * kRec + lc];
        bv[u] = Z[(size_t)kf[u] * kRec + lc];
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const bool live = j0 + u < m;
        const double a = (live && a_lane) ? av[u] : 0.0;
        const double b = (live && b_lane) ? bv[u] : ((live && one_lane && ke[u] == kf[u]) ? 1.0 : 0.0);
        dmma_m8n8k4(acc0, acc1, a, b);
      }
    }
  }
  if (g >= 6) return;
  // lane (g, q) holds the sums of row g, columns 2q and 2q + 1 (column 6 = right-hand side)
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int col = 2 * q + v;
    const int idx = col < 6 ? 6 * g + col : (col == 6 ? 36 + g : -1);
    if (idx < 0) continue;
    const double sum = v ? acc1 : acc0;
    if (c1 - c0 == 1) schur_store(d, bi, bj, idx, sum, radius, min_diag, max_diag, include_cam);
    else d.sch_partial[(size_t)w * 48 + idx] = sum;
  }
}

// camera pairs spanning several chunks: ordered sum of the chunk partials (thread / value)
__global__ void __launch_bounds__(kThreads)
ba_schur_multi_kernel(BaDev d, double radius, double min_diag, double max_diag, int include_cam) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  const int m = t / 48, idx = t - 48 * m;
  if (m >= d.sch_nmulti) return;
  const int pair = d.sch_multi[m];
  int bi = (int)((sqrtf(8.0f * (float)pair + 1.0f) - 1.0f) * 0.5f);
  while (tri(bi + 1) <= pair) ++bi;
  while (tri(bi) > pair) --bi;
  const int bj = pair - tri(bi);
  double sum = 0.0;
  for (int c = d.sch_pair_chunk[pair]; c < d.sch_pair_chunk[pair + 1]; ++c)
    sum += d.sch_partial[(size_t)c * 48 + idx];
  schur_store(d, bi, bj, idx, sum, radius, min_diag, max_diag, include_cam);
}

}  // namespace

// ============================================================================================
// Builds the gather lists of `d` (d.cam_block, d.obs_cam, d.pt_start, d.pt_var must be resident).
// alloc(bytes) hands out device memory that lives as long as the problem.
cudaError_t build_schur_lists(BaDev& d, void* (*alloc)(void*, size_t), void* alloc_ctx,
                              cudaStream_t s) {
  d.sch_npairs = d.NB * (d.NB + 1) / 2;
  d.sch_nchunks = 0;
  d.sch_nmulti = 0;
  d.Lv = (double*)alloc(alloc_ctx, sizeof(double) * 9 * (size_t)(d.P > 0 ? d.P : 1));
  if (!d.Lv) return cudaErrorMemoryAllocation;
  if (d.NB == 0) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  auto tmp_alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMallocAsync(p, bytes < 16 ? 16 : bytes, s);
  };
  const int P = d.P, npairs = d.sch_npairs;
  int64_t *counts = nullptr, *offs = nullptr;
  tmp_alloc((void**)&counts, sizeof(int64_t) * ((size_t)P + 1));
  tmp_alloc((void**)&offs, sizeof(int64_t) * ((size_t)P + 1));
  if (e != cudaSuccess) return e;
  cudaMemsetAsync(counts, 0, sizeof(int64_t) * ((size_t)P + 1), s);
  if (P > 0) sch_count_kernel<<<(P + 255) / 256, 256, 0, s>>>(d, counts);
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, counts, offs, P + 1, s);
  tmp_alloc(&cub_tmp, cub_bytes);
  if (e != cudaSuccess) return e;
  cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, counts, offs, P + 1, s);
  int64_t nent = 0;
  e = cudaMemcpyAsync(&nent, offs + P, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  cudaFreeAsync(cub_tmp, s);
  cub_tmp = nullptr;

  uint64_t *keys = nullptr, *keys2 = nullptr;
  int *vals = nullptr, *vals2 = nullptr;
  tmp_alloc((void**)&keys, sizeof(uint64_t) * (size_t)nent);
  tmp_alloc((void**)&keys2, sizeof(uint64_t) * (size_t)nent);
  tmp_alloc((void**)&vals, sizeof(int) * (size_t)nent);
  tmp_alloc((void**)&vals2, sizeof(int) * (size_t)nent);
  if (e != cudaSuccess) return e;
  if (P > 0 && nent > 0) sch_emit_kernel<<<(P + 255) / 256, 256, 0, s>>>(d, offs, keys, vals);
  const uint64_t* skeys = keys;
  const int* svals = vals;
  if (nent > 0) {
    int pair_bits = 1;
    while ((1ll << pair_bits) < (long long)npairs) ++pair_bits;
    cub::DoubleBuffer<uint64_t> kb(keys, keys2);
    cub::DoubleBuffer<int> vb(vals, vals2);
    cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, kb, vb, (int64_t)nent, 0, 32 + pair_bits, s);
    tmp_alloc(&cub_tmp, cub_bytes);
    if (e != cudaSuccess) return e;
    cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kb, vb, (int64_t)nent, 0, 32 + pair_bits, s);
    skeys = kb.Current();
    svals = vb.Current();
  }
  // persistent lists
  d.sch_pair_start = (int64_t*)alloc(alloc_ctx, sizeof(int64_t) * ((size_t)npairs + 1));
  d.sch_pair_chunk = (int*)alloc(alloc_ctx, sizeof(int) * ((size_t)npairs + 1));
  d.sch_ent = (int2*)alloc(alloc_ctx, sizeof(int2) * (size_t)(nent > 0 ? nent : 1));
  int* pair_nchunks = nullptr;
  tmp_alloc((void**)&pair_nchunks, sizeof(int) * ((size_t)npairs + 1));
  if (!d.sch_pair_start || !d.sch_pair_chunk || !d.sch_ent || e != cudaSuccess)
    return e != cudaSuccess ? e : cudaErrorMemoryAllocation;
  cudaMemsetAsync(pair_nchunks, 0, sizeof(int) * ((size_t)npairs + 1), s);
  sch_pair_start_kernel<<<(npairs + 1 + 255) / 256, 256, 0, s>>>(skeys, nent, npairs,
                                                                  d.sch_pair_start, pair_nchunks);
  void* cub_tmp2 = nullptr;
  cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, pair_nchunks, d.sch_pair_chunk, npairs + 1, s);
  tmp_alloc(&cub_tmp2, cub_bytes);
  if (e != cudaSuccess) return e;
  cub::DeviceScan::ExclusiveSum(cub_tmp2, cub_bytes, pair_nchunks, d.sch_pair_chunk, npairs + 1, s);
  int nchunks = 0;
  e = cudaMemcpyAsync(&nchunks, d.sch_pair_chunk + npairs, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  d.sch_nchunks = nchunks;
  d.sch_chunk_pair = (int*)alloc(alloc_ctx, sizeof(int) * (size_t)nchunks);
  int* n_multi_dev = nullptr;
  tmp_alloc((void**)&n_multi_dev, sizeof(int));
  const int max_multi = nchunks > npairs ? (nchunks - npairs) : 0;  // pairs with >= 2 chunks
  d.sch_multi = max_multi > 0 ? (int*)alloc(alloc_ctx, sizeof(int) * (size_t)max_multi) : nullptr;
  if (!d.sch_chunk_pair || e != cudaSuccess) return e != cudaSuccess ? e : cudaErrorMemoryAllocation;
  cudaMemsetAsync(n_multi_dev, 0, sizeof(int), s);
  sch_fill_chunks_kernel<<<(npairs + 255) / 256, 256, 0, s>>>(npairs, d.sch_pair_chunk,
                                                               d.sch_chunk_pair, d.sch_multi,
                                                               n_multi_dev);
  if (nent > 0)
    sch_extract_kernel<<<(unsigned)((nent + 255) / 256), 256, 0, s>>>(skeys, svals, nent, d.sch_ent);
  int n_multi = 0;
  e = cudaMemcpyAsync(&n_multi, n_multi_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  d.sch_nmulti = n_multi;
  if (n_multi > 0) {  // (pairs are independent: the order of sch_multi does not matter)
    d.sch_partial = (double*)alloc(alloc_ctx, sizeof(double) * 48 * (size_t)nchunks);
    if (!d.sch_partial) return cudaErrorMemoryAllocation;
  }
  d.Zrec = (double*)alloc(alloc_ctx, sizeof(double) * kRec * (size_t)(d.K > 0 ? d.K : 1));
  if (!d.Zrec) return cudaErrorMemoryAllocation;
  for (void* p : {(void*)counts, (void*)offs, (void*)keys, (void*)keys2, (void*)vals,
                  (void*)vals2, cub_tmp, cub_tmp2, (void*)pair_nchunks, (void*)n_multi_dev})
    if (p) cudaFreeAsync(p, s);
  return cudaGetLastError();
}

int launch_build_reduced_system(const BaDev& d, double radius, double min_diag, double max_diag,
                                bool include_camera_terms, cudaStream_t s) {
  int n = 0;
  if (d.P > 0) {
    ba_point_damp_kernel<<<(d.P + kThreads - 1) / kThreads, kThreads, 0, s>>>(d, radius, min_diag,
                                                                              max_diag);
    ++n;
  }
  if (d.NB == 0) return n;
  if (d.K > 0) {
    ba_zbuild_kernel<<<(unsigned)((d.K + kZThreads - 1) / kZThreads), kZThreads, 0, s>>>(d);
    ++n;
  }
  const int inc = include_camera_terms ? 1 : 0;
  const int64_t threads = (int64_t)d.sch_nchunks * 32;
  ba_schur_gather_kernel<<<(unsigned)((threads + kGThreads - 1) / kGThreads), kGThreads, 0, s>>>(
      d, radius, min_diag, max_diag, inc);
  ++n;
  if (d.sch_nmulti > 0) {
    const int64_t t2 = (int64_t)d.sch_nmulti * 48;
    ba_schur_multi_kernel<<<(unsigned)((t2 + kThreads - 1) / kThreads), kThreads, 0, s>>>(
        d, radius, min_diag, max_diag, inc);
    ++n;
  }
  return n;
}

}  // namespace ppsfm
