// dense_chol.cu — FP64 dense Cholesky solve of the reduced camera system S dc = rhs.
//
// In the reference this step is hidden inside ceres::Solve (SPARSE_SCHUR / DENSE_SCHUR,
// src/optim/bundle_adjustment.cc:275-286).  The reduced camera matrix of a BA problem with
// hundreds of cameras that all share points is dense, so it is factored densely.
//
// The factorisation of a 3000 x 3000 system is LATENCY bound (47 dependent 64-wide steps), not
// flop bound (9 GFLOP), so it runs as ONE persistent dataflow kernel instead of ~140 dependent
// launches:
//   * the lower triangle is cut into 64x64 tiles; tile (i, j) is one task: accumulate
//       C = A_ij - sum_{k<j} X_ik X_jk^T        (FP64 tensor-core MMA, mma.sync.m8n8k4.f64 —
//                                                tcgen05 has no FP64 kind; the one genuinely
//                                                dense contraction of the path)
//     then   j == i : L_jj = chol(C)            (16x16 register-resident warp factorisations)
//            j <  i : X_ij = C L_jj^-T          (blocked triangular solve on the tensor cores)
//   * tasks are handed out through an atomic ticket in column-major order, so a task only ever
//     waits for tasks with a smaller ticket (they are running or finished): no deadlock,
//     whatever the number of resident CTAs;
//   * a finished tile is published with a release store of its flag; consumers poll with acquire
//     loads and stream the operand tiles through a cp.async double buffer (left-looking:
//     accumulators stay in registers, no read-modify-write of the trailing matrix).
// The right-hand side rides along as an extra matrix row ("bordered" factorisation), which
// yields y = L^-1 rhs for free; the backward substitution L^T x = y is a second dataflow kernel
// (one CTA per 64-block, chained by flags).
//
// Matrix layout: row-major, leading dimension ld (multiple of 64), rows [0, n) = S (lower
// triangle referenced), row n = rhs^T, rows (n, ld) zero padding.
#include "common.h"
#include "dense_chol.h"

namespace ppsfm {

namespace {

constexpr int NB = 64;
constexpr int kCS = 68;  // row stride (doubles) of a 64x64 tile in shared memory
constexpr int kKC = 32;  // k-chunk width of the operand pipeline
constexpr int kKS = 36;  // row stride (doubles) of a 64x32 chunk in shared memory
constexpr int kStageDoubles = 64 * kKS;
constexpr int kFactorSmem = 2 * 2 * kStageDoubles * (int)sizeof(double);  // 73 728 B
static_assert(2 * 64 * kCS * (int)sizeof(double) <= kFactorSmem, "tile pair must fit the stages");

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Polling uses RELAXED loads: an acquire load is followed by an L1 invalidation (CCTL.IVALL),
// and CTAs that spin on flags would invalidate the L1 / stall the LSU of the SM they share with
// the CTA on the critical path.  One acquire fence after the flag has been seen orders the data.
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ double2 ld_cg2(const double* p) {
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_cg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// warp shuffle of a double without the convergence bookkeeping nvcc adds around __shfl_sync in
// warp-specialised code
__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(lo) : "r"(src));
  asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(hi) : "r"(src));
  return __hiloint2double(hi, lo);
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Optional per-tile event trace (PPSFM_CHOL_TRACE=<file>): 16 timestamp slots per tile.
#define CHOL_TRACE(slot)                                                          \
  do {                                                                            \
    if (trace != nullptr && threadIdx.x == 0) trace[16 * (size_t)(i * T + j) + (slot)] = global_ns(); \
  } while (0)

// Work-buffer layout (doubles unless noted): one Lpack and one Linv per 64-block, y/x scratch,
// then the int flags.
struct Work {
  double* lpack;  // [nblk][64*64]  L_jj with its 16x16 diagonal sub-blocks replaced by their inverses
  double* linv;   // [nblk][64*64]  dense L_jj^-1 (back-substitution)
  int* flags;     // [0] factor ticket, [1] back-solve ticket, [2 .. 2+T*T) tile flags,
                  // then [nblk] x flags
};
__host__ __device__ inline Work work_layout(double* base, int n) {
  const size_t nblk = (size_t)(n + NB - 1) / NB;
  Work w;
  w.lpack = base;
  w.linv = base + nblk * NB * NB;
  w.flags = reinterpret_cast<int*>(base + 2 * nblk * NB * NB);
  return w;
}
inline size_t work_flag_ints(int n) {
  const size_t T = (size_t)(n + 1 + NB - 1) / NB, nblk = (size_t)(n + NB - 1) / NB;
  return 2 + T * T + nblk;
}

// ------------------------------------------------------------------------------------------
// 8x8 Cholesky of the diagonal sub-block at (k1, k1) of the tile in shared memory, by ONE warp.
// The symmetric block is spread over the warp, two elements per lane: lane (r, c) = (lane >> 3,
// lane & 7) holds A[r][c] and A[r+4][c].  A pivot step then needs only four shuffles (pivot,
// row element a_jc, column elements a_rj and a_(r+4)j) and two FMAs per lane, so the warp's
// instruction stream stays far below the length of the dependency chain
// shuffle -> reciprocal -> multiply -> FMA (~110 cycles per pivot; measured on B200: DFMA 8,
// SHFL.64 26, reciprocal 58 cycles).  Square roots are taken once, after the eight steps.
// Writes L (lower) back, and rdiag[k1 + j] = 1 / l_jj.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool factor8(double* Cs, int k1, double* rdiag, int lane) {
  const int r = lane >> 3, c = lane & 7;
  double e0, e1;
  {
    const int r0 = r > c ? r : c, c0 = r > c ? c : r;
    const int r1 = r + 4 > c ? r + 4 : c, c1 = r + 4 > c ? c : r + 4;
    e0 = Cs[(k1 + r0) * kCS + k1 + c0];
    e1 = Cs[(k1 + r1) * kCS + k1 + c1];
  }
  __syncwarp();
  bool bad = false;
  double l0 = 0.0, l1 = 0.0, pv = 1.0;  // column c of the factor (unscaled) and its pivot
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int jl = (j & 3) << 3;
    const double src = (j < 4) ? e0 : e1;  // row j lives in slot j >> 2 of lanes (j & 3, *)
    double piv = shfl_d(src, jl | j);
    const double u = shfl_d(src, jl | c);          // a_jc
    const double t0 = shfl_d(e0, (r << 3) | j);    // a_rj
    const double t1 = shfl_d(e1, (r << 3) | j);    // a_(r+4)j
    bad = bad || !(piv > 0.0);
    piv = (piv > 0.0) ? piv : 1.0;
    if (c == j) {
      l0 = e0;
      l1 = e1;
      pv = piv;
    }
    const double sc = u * (1.0 / piv);
    e0 -= t0 * sc;
    e1 -= t1 * sc;
  }
  const double rs = rsqrt(pv);
  if (r >= c) Cs[(k1 + r) * kCS + k1 + c] = l0 * rs;
  Cs[(k1 + r + 4) * kCS + k1 + c] = l1 * rs;  // r + 4 >= c for the rows that matter; the
  if (r == 0) rdiag[k1 + c] = rs;             // entries above the diagonal are never read
  return bad;
}

// ------------------------------------------------------------------------------------------
// Diagonal task: Cs (64 x kCS, lower triangle valid) -> L_jj.  kb = number of real rows in this
// block (< 64 only for the last block, whose tile also carries the right-hand-side row).
// Publishes L (into A), Lpack (for the triangular solves of the tiles below) and then — off the
// critical path — the dense inverse for the back-substitution.
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void diag_task(double* __restrict__ A, int ld, int n, int j, double* smem,
                          const Work& w, int T, int* __restrict__ status,
                          unsigned long long* __restrict__ trace) {
  const int i = j;
  double* Cs = smem;                                   // [64][kCS]
  double* Tm = smem + 64 * kCS;                        // [64][kCS] L^-1: 16x16 diagonal blocks
                                                       // first, the rest after the publish
  __shared__ double rdiag[NB];
  __shared__ double rhs_row[NB];
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31;
  // warp index broadcast from lane 0: the compiler then knows the warp-specialised branches
  // below are warp-uniform and emits plain SHFL instead of WARPSYNC.COLLECTIVE sequences
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int k0 = j * NB;
  const int kb = min(NB, n - k0);
  if (tid == 0) s_bad = 0;
  if (kb < NB) {
    // last, partial block: keep the right-hand-side row aside, pad with the identity
    if (tid < NB) rhs_row[tid] = (tid < kb) ? Cs[kb * kCS + tid] : 0.0;
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int r = idx >> 6, c = idx & 63;
      if (r >= kb || c >= kb) Cs[r * kCS + c] = (r == c) ? 1.0 : 0.0;
    }
  }
  // (only the lower triangle of Cs is ever read; Tm is fully written before it is read)
  __syncthreads();
  CHOL_TRACE(4);
  const int g = lane >> 2, q = lane & 3;  // mma fragment coordinates
#pragma unroll 1
  for (int bk = 0; bk < 8; ++bk) {
    const int k1 = 8 * bk;
    if (warp == 0) {
      const bool bad = factor8(Cs, k1, rdiag, lane);
      if (bad && lane == 0) s_bad = 1;
    }
    __syncthreads();
    const int below = NB - (k1 + 8);
    if (below > 0) {
      // panel: row r of X = A_r L^-T by forward substitution, one thread per row
      if (tid < below) {
        double* rowp = Cs + (k1 + 8 + tid) * kCS + k1;
        double x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = rowp[c];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          x[c] *= rdiag[k1 + c];
#pragma unroll
          for (int p = c + 1; p < 8; ++p) x[p] -= x[c] * Cs[(k1 + p) * kCS + k1 + c];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) rowp[c] = x[c];
      }
      __syncthreads();
      // trailing update inside the tile on the FP64 tensor cores: 8x8 output tiles (lower
      // part), C -= X_ti X_tj^T with k = 8 (two m8n8k4 steps)
      const int nb = below >> 3;
      const int ntile = nb * (nb + 1) / 2;
      const double* X = Cs + (k1 + 8) * kCS + k1;
      for (int t = warp; t < ntile; t += 8) {
        int ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        const int tj = t - ti * (ti + 1) / 2;
        double* cp = Cs + (k1 + 8 + 8 * ti + g) * kCS + k1 + 8 + 8 * tj + 2 * q;
        double2 cv = *reinterpret_cast<double2*>(cp);
        const double a0 = -X[(8 * ti + g) * kCS + q], a1 = -X[(8 * ti + g) * kCS + 4 + q];
        const double b0 = X[(8 * tj + g) * kCS + q], b1 = X[(8 * tj + g) * kCS + 4 + q];
        dmma_m8n8k4(cv.x, cv.y, a0, b0);
        dmma_m8n8k4(cv.x, cv.y, a1, b1);
        *reinterpret_cast<double2*>(cp) = cv;
      }
      __syncthreads();
    }
    if (bk & 1) CHOL_TRACE(5 + bk - 1);
  }
  // Inverses of the 16x16 diagonal sub-blocks (what the triangular solves of the tiles below
  // use): 8x8 inverses by forward substitution, one thread per column, then
  // inv([A 0; B C]) = [inv A, 0; -inv(C) B inv(A), inv C].
  if (tid < 64) {
    const int blk = tid >> 3, c = tid & 7, o = 8 * blk;
    double m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i >= c) {
        double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < i && k >= c) acc -= Cs[(o + i) * kCS + o + k] * m[k];
        m[i] = acc * rdiag[o + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) Tm[(o + i) * kCS + o + c] = m[i];
    // upper-right 8x8 of the 16-block this 8-block belongs to is zero
    if ((blk & 1) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) Tm[(o + i) * kCS + o + 8 + c] = 0.0;
    }
  }
  __syncthreads();
  {
    // lower-left 8x8 of each 16-block: T = B inv(A) (into scratch right of the tile's 16-block
    // row, Tm columns [48, 56) are free until the dense inverse is assembled), then -inv(C) T
    const int blk = tid >> 6, r = (tid >> 3) & 7, c = tid & 7, o = 16 * blk;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += Cs[(o + 8 + r) * kCS + o + k] * Tm[(o + k) * kCS + o + c];
    __syncthreads();
    double* scratch = Tm + (size_t)(o + r) * kCS + ((blk == 3) ? 0 : 56);  // outside block `blk`
    scratch[c] = acc;
    __syncthreads();
    double res = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      res -= Tm[(o + 8 + r) * kCS + o + 8 + k] * Tm[(size_t)(o + k) * kCS + ((blk == 3) ? 0 : 56) + c];
    __syncthreads();
    Tm[(o + 8 + r) * kCS + o + c] = res;
  }
  __syncthreads();
  CHOL_TRACE(13);
  // publish Lpack (what the triangular solves of the tiles below need), then raise the flag
  double* lpack = w.lpack + (size_t)j * NB * NB;
  for (int idx = tid; idx < NB * NB / 2; idx += 256) {
    const int r = idx >> 5, c = (idx & 31) * 2;
    const double* src = (((r >> 4) == (c >> 4)) ? Tm : Cs) + r * kCS + c;
    *reinterpret_cast<double2*>(lpack + r * NB + c) = *reinterpret_cast<const double2*>(src);
  }
  __syncthreads();
  CHOL_TRACE(14);
  if (tid == 0) {
    if (s_bad) atomicExch(status, 1);
    __threadfence();
    st_release(w.flags + 2 + j * T + j, 1);
  }
  CHOL_TRACE(2);
  // ---- off the critical path: L into A (read by the back-substitution kernel), then the dense
  //      L^-1 by block forward substitution
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    if (r < kb && c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = Cs[r * kCS + c];
  }
  //      Linv[bi][bj] = -I16[bi] * sum_{bk = bj}^{bi-1} L[bi][bk] Linv[bk][bj]
#pragma unroll 1
  for (int dist = 1; dist < 4; ++dist) {
    const int nblk = 4 - dist;  // blocks (bi = bj + dist, bj)
    double tmp[3];
    int cnt = 0;
    for (int idx = tid; idx < nblk * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, c = idx & 15;
      double acc = 0.0;
      for (int p2 = 16 * bj; p2 < 16 * bi; ++p2)
        acc += Cs[(16 * bi + r) * kCS + p2] * Tm[p2 * kCS + 16 * bj + c];
      tmp[cnt] = acc;
    }
    // tmp -> scratch above the diagonal of Tm (unused otherwise): Tm[bj-rows][bi-cols]
    cnt = 0;
    for (int idx = tid; idx < nblk * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, c = idx & 15;
      Tm[(16 * bj + r) * kCS + 16 * bi + c] = tmp[cnt];
    }
    __syncthreads();
    cnt = 0;
    for (int idx = tid; idx < nblk * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, c = idx & 15;
      double acc = 0.0;
#pragma unroll
      for (int p2 = 0; p2 < 16; ++p2)
        acc += Tm[(16 * bi + r) * kCS + 16 * bi + p2] * Tm[(16 * bj + p2) * kCS + 16 * bi + c];
      tmp[cnt] = -acc;
    }
    __syncthreads();
    cnt = 0;
    for (int idx = tid; idx < nblk * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, c = idx & 15;
      Tm[(16 * bi + r) * kCS + 16 * bj + c] = tmp[cnt];
    }
    __syncthreads();
  }
  double* linv = w.linv + (size_t)j * NB * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    linv[idx] = (c <= r) ? Tm[r * kCS + c] : 0.0;
  }
  // the right-hand-side row of the last, partial block: y = rhs L^-T
  if (kb < NB && tid < kb) {
    double acc = 0.0;
    for (int p = 0; p <= tid; ++p) acc += rhs_row[p] * Tm[tid * kCS + p];
    A[(size_t)n * ld + k0 + tid] = acc;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Off-diagonal task: X = C L_jj^-T for the 64x64 tile in Cs, warp-local (warp w owns rows
// 8w..8w+7), in four 16-column steps:  X_b = (C_b - sum_{b'<b} X_b' L_bb'^T) inv(L_bb)^T.
// acc[nb] are the tile's C fragments (nb = 8-column block).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void trsm_task(double* __restrict__ A, int ld, int i, int j,
                                          double (&acc)[8][2], double* smem, const Work& w,
                                          int T, unsigned long long* __restrict__ trace) {
  double* Xs = smem;             // [64][kCS]: own rows, X blocks as they are produced
  double* Lp = smem + 64 * kCS;  // [64][kCS]: Lpack_j
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int row = warp * 8 + g;
  if (tid == 0) {
    const int* f = w.flags + 2 + j * T + j;
    while (ld_relaxed(f) == 0) __nanosleep(40);
    fence_acquire();
  }
  __syncthreads();
  {
    const double* src = w.lpack + (size_t)j * NB * NB;
    for (int p = tid; p < NB * 32; p += 256) {  // 16-byte pieces
      const int r = p >> 5, s = p & 31;
      cp_async16(Lp + r * kCS + 2 * s, src + r * NB + 2 * s);
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    double t0[2] = {acc[2 * b][0], acc[2 * b][1]};
    double t1[2] = {acc[2 * b + 1][0], acc[2 * b + 1][1]};
#pragma unroll
    for (int kk = 0; kk < 16 * b; kk += 4) {
      const double a = -Xs[row * kCS + kk + q];
      const double b0 = Lp[(16 * b + g) * kCS + kk + q];
      const double b1 = Lp[(16 * b + 8 + g) * kCS + kk + q];
      dmma_m8n8k4(t0[0], t0[1], a, b0);
      dmma_m8n8k4(t1[0], t1[1], a, b1);
    }
    // T (C-fragment layout) -> shared memory -> A-fragment layout; rows are warp-private
    *reinterpret_cast<double2*>(Xs + row * kCS + 16 * b + 2 * q) = make_double2(t0[0], t0[1]);
    *reinterpret_cast<double2*>(Xs + row * kCS + 16 * b + 8 + 2 * q) = make_double2(t1[0], t1[1]);
    __syncwarp();
    double x0[2] = {0.0, 0.0}, x1[2] = {0.0, 0.0};
#pragma unroll
    for (int kk = 0; kk < 16; kk += 4) {
      const double a = Xs[row * kCS + 16 * b + kk + q];
      const double b0 = Lp[(16 * b + g) * kCS + 16 * b + kk + q];
      const double b1 = Lp[(16 * b + 8 + g) * kCS + 16 * b + kk + q];
      dmma_m8n8k4(x0[0], x0[1], a, b0);
      dmma_m8n8k4(x1[0], x1[1], a, b1);
    }
    __syncwarp();
    *reinterpret_cast<double2*>(Xs + row * kCS + 16 * b + 2 * q) = make_double2(x0[0], x0[1]);
    *reinterpret_cast<double2*>(Xs + row * kCS + 16 * b + 8 + 2 * q) = make_double2(x1[0], x1[1]);
    __syncwarp();
    double* dst = A + (size_t)(i * NB + row) * ld + j * NB + 16 * b + 2 * q;
    *reinterpret_cast<double2*>(dst) = make_double2(x0[0], x0[1]);
    *reinterpret_cast<double2*>(dst + 8) = make_double2(x1[0], x1[1]);
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    st_release(w.flags + 2 + i * T + j, 1);
  }
  CHOL_TRACE(2);
}

// ------------------------------------------------------------------------------------------
// The factorisation kernel: persistent CTAs (256 threads) pulling tile tasks from a ticket.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3)
chol_factor_kernel(double* __restrict__ A, int ld, int n, double* __restrict__ work_base,
                   int* __restrict__ status, unsigned long long* __restrict__ trace) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_ticket, s_ready;
  const Work w = work_layout(work_base, n);
  const int T = ld / NB;
  const int ncols = (n + NB - 1) / NB;
  const int ntiles = ncols * T - ncols * (ncols - 1) / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int row = warp * 8 + g;
  int* tile_flags = w.flags + 2;

  for (;;) {
    __syncthreads();  // previous task is completely done with shared memory
    if (tid == 0) s_ticket = atomicAdd(w.flags, 1);
    __syncthreads();
    const int t = s_ticket;
    if (t >= ntiles) break;
    // column-major order, diagonal tile first: column j starts at j*T - j(j-1)/2
    int j = 0, off = 0;
    while (j + 1 < ncols && off + (T - j) <= t) {
      off += T - j;
      ++j;
    }
    const int i = j + (t - off);
    CHOL_TRACE(0);

    // accumulators start as A_ij; the k loop subtracts X_ik X_jk^T
    double acc[8][2];
    {
      const double* src = A + (size_t)(i * NB + row) * ld + j * NB + 2 * q;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const double2 v = ld_cg2(src + 8 * nb);
        acc[nb][0] = v.x;
        acc[nb][1] = v.y;
      }
    }
    const int nchunks = 2 * j;
    const bool diag = (i == j);
    int issued = 0, known_k = 0;  // chunks issued; operand tiles of columns < known_k are ready
    auto issue = [&](int c) {
      const int k = c >> 1, half = c & 1, st = c & 1;
      double* si = smem + (size_t)(st * 2) * kStageDoubles;
      double* sj = si + kStageDoubles;
      const double* gi = A + (size_t)(i * NB) * ld + k * NB + half * kKC;
      const double* gj = A + (size_t)(j * NB) * ld + k * NB + half * kKC;
      for (int p = tid; p < 64 * 16; p += 256) {
        const int r = p >> 4, s = p & 15;
        cp_async16(si + r * kKS + 2 * s, gi + (size_t)r * ld + 2 * s);
        if (!diag) cp_async16(sj + r * kKS + 2 * s, gj + (size_t)r * ld + 2 * s);
      }
      cp_async_commit();
    };
    for (int c = 0; c < nchunks; ++c) {
      // keep up to two chunks in flight; block on a flag only when there is nothing to compute
      while (issued < nchunks && issued < c + 2) {
        const int k = issued >> 1;
        if (k >= known_k) {
          const bool must = (issued == c);
          if (tid == 0) {
            const int* fi = tile_flags + i * T + k;
            const int* fj = tile_flags + j * T + k;
            int ok = (ld_relaxed(fi) != 0) && (ld_relaxed(fj) != 0);
            while (!ok && must) {
              __nanosleep(40);
              ok = (ld_relaxed(fi) != 0) && (ld_relaxed(fj) != 0);
            }
            if (ok) fence_acquire();
            s_ready = ok;
          }
          __syncthreads();
          const int ok = s_ready;
          __syncthreads();
          if (!ok) break;
          known_k = k + 1;
        }
        issue(issued);
        ++issued;
      }
      if (issued - c - 1 >= 1) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();
      const int st = c & 1;
      const double* Xi = smem + (size_t)(st * 2) * kStageDoubles;
      const double* Xj = diag ? Xi : Xi + kStageDoubles;
#pragma unroll
      for (int kk = 0; kk < kKC; kk += 4) {
        const double a = -Xi[row * kKS + kk + q];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          const double b = Xj[(nb * 8 + g) * kKS + kk + q];
          dmma_m8n8k4(acc[nb][0], acc[nb][1], a, b);
        }
      }
      __syncthreads();  // stage st may be overwritten
    }
    CHOL_TRACE(1);
    if (diag) {
      // C fragments -> shared tile
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        *reinterpret_cast<double2*>(smem + row * kCS + nb * 8 + 2 * q) =
            make_double2(acc[nb][0], acc[nb][1]);
      __syncthreads();
      diag_task(A, ld, n, j, smem, w, T, status, trace);
      CHOL_TRACE(3);
    } else {
      trsm_task(A, ld, i, j, acc, smem, w, T, trace);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward substitution L^T x = y as a dataflow kernel: one CTA per 64-block (descending by
// ticket).  CTA j streams the tiles L[b][j], b > j, through a cp.async double buffer, applies
// y_j -= L[b][j]^T x_b as soon as x_b is published, then x_j = L_jj^-T y_j.
// ------------------------------------------------------------------------------------------
constexpr int kBackSmem = 3 * NB * NB * (int)sizeof(double);

__global__ void __launch_bounds__(256)
chol_backsolve_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ work_base,
                      double* __restrict__ x) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double yj[NB], xb[NB], red[4][NB];
  __shared__ int s_ticket;
  const Work w = work_layout(work_base, n);
  const int T = ld / NB;
  const int nblk = (n + NB - 1) / NB;
  int* xflags = w.flags + 2 + T * T;
  const int tid = threadIdx.x, c = tid & 63, part = tid >> 6;
  if (tid == 0) s_ticket = atomicAdd(w.flags + 1, 1);
  __syncthreads();
  const int j = nblk - 1 - s_ticket;
  if (j < 0) return;
  const int k0 = j * NB;
  const int kb = min(NB, n - k0);
  double* Linv = smem;                 // [64][64]
  double* stage0 = smem + NB * NB;     // two tile stages
  auto issue_tile = [&](int b, int st) {
    double* dst = stage0 + (size_t)st * NB * NB;
    const double* src = A + (size_t)(b * NB) * ld + k0;
    for (int p = tid; p < NB * 32; p += 256) {
      const int r = p >> 5, s = p & 31;
      cp_async16(dst + r * NB + 2 * s, src + (size_t)r * ld + 2 * s);
    }
    cp_async_commit();
  };
  {
    const double* src = w.linv + (size_t)j * NB * NB;
    for (int p = tid; p < NB * 32; p += 256) cp_async16(Linv + 2 * p, src + 2 * p);
    cp_async_commit();
  }
  if (tid < NB) yj[tid] = (tid < kb) ? ld_cg(A + (size_t)n * ld + k0 + tid) : 0.0;
  int issued_b = nblk - 1;  // next tile to issue
  for (int cnt = 0; cnt < 2 && issued_b > j; ++cnt, --issued_b) issue_tile(issued_b, (nblk - 1 - issued_b) & 1);
  for (int b = nblk - 1; b > j; --b) {
    const int st = (nblk - 1 - b) & 1;
    if (tid == 0) {
      while (ld_relaxed(xflags + b) == 0) __nanosleep(20);
      fence_acquire();
    }
    // tile b has been issued; at most one younger group is in flight
    if (issued_b < b - 1) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const int kbb = min(NB, n - b * NB);
    if (tid < NB) xb[tid] = (tid < kbb) ? ld_cg(x + b * NB + tid) : 0.0;
    __syncthreads();
    const double* Lt = stage0 + (size_t)st * NB * NB;
    double s = 0.0;
#pragma unroll 4
    for (int r = part; r < kbb; r += 4) s += Lt[r * NB + c] * xb[r];
    red[part][c] = s;
    __syncthreads();
    if (tid < NB) yj[tid] -= red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
    // stage st is free again
    if (issued_b > j) {
      issue_tile(issued_b, (nblk - 1 - issued_b) & 1);
      --issued_b;
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();
  {
    double s = 0.0;
#pragma unroll 4
    for (int r = part; r < NB; r += 4) s += Linv[r * NB + c] * yj[r];  // (L^-1)^T y
    red[part][c] = s;
  }
  __syncthreads();
  if (tid < kb) x[k0 + tid] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    st_release(xflags + j, 1);
  }
}

}  // namespace

int chol_ld(int n) { return ((n + 1 + NB - 1) / NB) * NB; }
size_t chol_work_doubles(int n) {
  const size_t nblk = (size_t)(n + NB - 1) / NB;
  return 2 * nblk * NB * NB + (work_flag_ints(n) + 1) / 2 + 2;
}

// Factor + solve.  A: ld x ld (see header).  x: n doubles (device).  status: device int, set to
// 1 if a non-positive pivot was met.  Asynchronous on `s`; returns the number of launches.
int chol_solve_bordered(double* A, int n, int ld, double* x, double* work, int* status,
                        cudaStream_t s) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(chol_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kFactorSmem);
    cudaFuncSetAttribute(chol_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kBackSmem);
  }
  const Work w = work_layout(work, n);
  cudaMemsetAsync(status, 0, sizeof(int), s);
  cudaMemsetAsync(w.flags, 0, sizeof(int) * work_flag_ints(n), s);
  const int T = ld / NB;
  const int ncols = (n + NB - 1) / NB;
  const int ntiles = ncols * T - ncols * (ncols - 1) / 2;
  int grid = 3 * num_sms;
  if (grid > ntiles) grid = ntiles;
  static const char* trace_path = std::getenv("PPSFM_CHOL_TRACE");
  unsigned long long* trace = nullptr;
  if (trace_path) {
    cudaMalloc(&trace, sizeof(unsigned long long) * 16 * (size_t)T * T);
    cudaMemsetAsync(trace, 0, sizeof(unsigned long long) * 16 * (size_t)T * T, s);
  }
  chol_factor_kernel<<<grid, 256, kFactorSmem, s>>>(A, ld, n, work, status, trace);
  chol_backsolve_kernel<<<ncols, 256, kBackSmem, s>>>(A, ld, n, work, x);
  if (trace_path) {  // development aid: dump "i j t0 t1 t2 t3" (ns) per tile
    std::vector<unsigned long long> h(16 * (size_t)T * T);
    cudaStreamSynchronize(s);
    cudaMemcpy(h.data(), trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost);
    cudaFree(trace);
    if (FILE* f = std::fopen(trace_path, "w")) {
      for (int i = 0; i < T; ++i)
        for (int j = 0; j <= i && j < ncols; ++j) {
          const unsigned long long* e = &h[16 * (size_t)(i * T + j)];
          std::fprintf(f, "%d %d", i, j);
          for (int k = 0; k < 16; ++k) std::fprintf(f, " %llu", e[k]);
          std::fprintf(f, "\n");
        }
      std::fclose(f);
    }
  }
  return 2;
}

}  // namespace ppsfm
