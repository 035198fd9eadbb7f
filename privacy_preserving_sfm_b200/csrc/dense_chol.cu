// dense_chol.cu — FP64 dense Cholesky solve of the reduced camera system S dc = rhs.
//
// In the reference this step is hidden inside ceres::Solve (SPARSE_SCHUR / DENSE_SCHUR,
// src/optim/bundle_adjustment.cc:275-286).  The reduced camera matrix of a BA problem with
// hundreds of cameras that all share points is dense, so it is factored densely.
//
// The factorisation of a 3000 x 3000 system is LATENCY bound (47 dependent 64-wide steps), not
// flop bound (9 GFLOP = 0.3 ms of FP64 tensor-core time), so it runs as ONE persistent dataflow
// kernel with two roles:
//   * HELPER CTAs pull 64x64 tile tasks (i, j) from an atomic ticket in column-major order and
//     accumulate  C = A_ij - sum_k X_ik X_jk^T  left-looking (accumulators stay in registers,
//     operand tiles stream through a 3-stage cp.async pipeline; FP64 tensor-core MMA
//     mma.sync.m8n8k4.f64 — tcgen05 has no FP64 kind).  Tiles with i >= j + 2 are then finished
//     by the blocked triangular solve X_ij = C L_jj^-T on the tensor cores; the diagonal and
//     sub-diagonal tiles are only PRE-accumulated (all terms that do not depend on the previous
//     column step) and handed to the walker.
//   * ONE WALKER CTA, alone on its SM, walks down the diagonal and keeps the whole dependency
//     chain  L_jj = chol(C_jj)  ->  X_(j+1)j = C_(j+1)j L_jj^-T  ->  C_(j+1)(j+1) -= X X^T  in
//     shared memory and registers: no flag round trip, no global-memory latency and no
//     co-resident tensor-core traffic on the critical path.
// A task only ever waits for tasks with a smaller ticket or for the walker, the walker is the
// first CTA that starts, and at most one CTA (the one that shares the walker's SM) leaves without
// working, so at least one helper exists and there is no deadlock whatever the number of resident
// CTAs and whatever else occupies the GPU.  Finished
// tiles are published with release stores; consumers poll with relaxed loads and one acquire fence.
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
constexpr int kStages = 3;
constexpr int kStageDoubles = 64 * kKS;
constexpr int kFactorSmem = kStages * 2 * kStageDoubles * (int)sizeof(double);  // 110 592 B
// the walker: three tiles, then 8 x (8x8) block inverses and 4 x (8x8) scratch blocks
constexpr int kWalkerInv = 3 * 64 * kCS;
static_assert((kWalkerInv + 12 * 64) * (int)sizeof(double) <= kFactorSmem, "walker's shared memory");

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// Polling uses RELAXED loads: an acquire load is followed by an L1 invalidation (CCTL.IVALL),
// and CTAs that spin on flags would invalidate the L1 / stall the LSU of the SM they share with
// the CTA on the critical path.  One acquire fence after the flag has been seen orders the data.
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
constexpr unsigned long long kXPending = ~0ull;  // x entries not yet solved (back-substitution)
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
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
  double out;  // (one asm block: ptxas shuffles the halves in place in the result's register pair)
  asm volatile(
      "{ .reg .b32 lo, hi;\n"
      "  mov.b64 {lo, hi}, %1;\n"
      "  shfl.sync.idx.b32 lo, lo, %2, 0x1f, 0xffffffff;\n"
      "  shfl.sync.idx.b32 hi, hi, %2, 0x1f, 0xffffffff;\n"
      "  mov.b64 %0, {lo, hi}; }"
      : "=d"(out)
      : "d"(v), "r"(src));
  return out;
}

// 1 / x to full double precision without the IEEE division sequence: the 20-bit hardware seed
// (MUFU.RCP64H) and two Newton steps (2^-20 -> 2^-40 -> 2^-80), four dependent DFMAs.  The pivot
// reciprocal sits on the factorisation's dependency chain (one per pivot, 3000 pivots in a row at
// 500 cameras); x is a positive, normal pivot here.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// tile t of a lower triangle of 8x8 blocks, row-major ((0,0), (1,0), (1,1), (2,0), ...): row << 4 | col
__constant__ unsigned char kTriRow[28] = {
    0x00, 0x10, 0x11, 0x20, 0x21, 0x22, 0x30, 0x31, 0x32, 0x33, 0x40, 0x41, 0x42, 0x43,
    0x44, 0x50, 0x51, 0x52, 0x53, 0x54, 0x55, 0x60, 0x61, 0x62, 0x63, 0x64, 0x65, 0x66};

// (optional per-tile event trace: PPSFM_CHOL_TRACE=<file>, 16 timestamp slots per tile)
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
// Work buffer: Lpack per 64-block (L_jj with its 16x16 diagonal sub-blocks replaced by their
// inverses: what the triangular solves and the back-substitution need), then the int flags:
// [0] factor ticket, [1] back-solve ticket, [2] role word, [3] walker SM id + 1,
// [4, 4+T*T) tile flags (X_ij final; for i == j: Lpack_j published), [.., +T*T) pre flags
// (diagonal / sub-diagonal tile pre-accumulated), then [nblk] x flags.
struct Work {
  double* lpack;
  int* flags;
  int *tile, *pre, *xf;
};
__host__ __device__ inline Work work_layout(double* base, int n) {
  const size_t nblk = (size_t)(n + NB - 1) / NB, T = (size_t)(n + 1 + NB - 1) / NB;
  Work w;
  w.lpack = base;
  w.flags = reinterpret_cast<int*>(base + nblk * NB * NB);
  w.tile = w.flags + 4;
  w.pre = w.tile + T * T;
  w.xf = w.pre + T * T;
  return w;
}
inline size_t work_flag_ints(int n) {
  const size_t T = (size_t)(n + 1 + NB - 1) / NB, nblk = (size_t)(n + NB - 1) / NB;
  return 4 + 2 * T * T + nblk;
}

__device__ __forceinline__ void wait_flag(const int* f) {
  while (ld_relaxed(f) == 0) __nanosleep(40);
  fence_acquire();
}

// ------------------------------------------------------------------------------------------
// 8x8 Cholesky of the diagonal sub-block at (k1, k1) of the tile in shared memory, by ONE warp,
// together with the inverse of its factor.
// The symmetric block is spread over the warp, two elements per lane: lane (r, c) = (lane >> 3,
// lane & 7) holds A[r][c] and A[r+4][c].  A pivot step then needs only four shuffles (pivot,
// row element a_jc, column elements a_rj and a_(r+4)j) and two FMAs per lane, so the warp's
// instruction stream stays far below the length of the dependency chain
// shuffle -> reciprocal -> multiply -> FMA (~110 cycles per pivot; measured on B200: DFMA 8,
// SHFL.64 26, reciprocal 58 cycles).  Square roots are taken once, after the eight steps.
// The same elimination is applied to an identity block held the same way (one more shuffle, a
// multiply and two predicated FMAs per pivot, none of them on the chain): A = L_u D L_u^T leaves
// L_u^-1 there, and inv(L) = D^-1/2 L_u^-1.  The panel below the block is then ONE small matrix
// product on the tensor cores instead of an eight-step substitution per row, and the 16x16
// inverses of the tile need no 8x8 substitutions either.
// Writes L (lower) to Cs and inv(L) (lower, zeros above the diagonal) to iv (8 x 8, row-major).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool factor8(double* Cs, int k1, double* iv, int lane) {
  const int r = lane >> 3, c = lane & 7;
  double e0, e1;
  {
    const int r0 = r > c ? r : c, c0 = r > c ? c : r;
    const int r1 = r + 4 > c ? r + 4 : c, c1 = r + 4 > c ? c : r + 4;
    e0 = Cs[(k1 + r0) * kCS + k1 + c0];
    e1 = Cs[(k1 + r1) * kCS + k1 + c1];
  }
  double i0 = (r == c) ? 1.0 : 0.0, i1 = (r + 4 == c) ? 1.0 : 0.0;  // the identity, same layout
  __syncwarp();
  // Fraction-free (Bareiss) elimination: with B_j the j-th leading minor (B_-1 = 1),
  //   b_rc <- (B_j b_rc - b_rj b_jc) / B_(j-1)      (b = true Schur complement x B_(j-1)),
  // so the reciprocal a step needs is that of the PREVIOUS pivot and leaves the dependency chain
  // (shuffle -> multiply -> FMA -> multiply, ~50 cycles, instead of shuffle -> reciprocal [58 +
  // Newton] -> multiply -> FMA).  The warp is alone on its scheduler while it factors, so the
  // step is bound by this chain and by its instruction count: columns <= j are simply left alone
  // (their lanes end up holding column j of the factor, scaled — no copies), non-positive pivots
  // are detected once at the end.  Magnitudes grow with the product of the pivots of ONE 8x8
  // block (<= 8 factors: far inside the double range for any matrix the factorisation survives).
  double rprev = 1.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int jl = (j & 3) << 3;
    const double src = (j < 4) ? e0 : e1;  // row j lives in slot j >> 2 of lanes (j & 3, *)
    const double isrc = (j < 4) ? i0 : i1;
    const double piv = shfl_d(src, jl | j);        // B_j
    const double u = shfl_d(src, jl | c);          // b_jc
    const double t0 = shfl_d(e0, (r << 3) | j);    // b_rj
    const double t1 = shfl_d(e1, (r << 3) | j);    // b_(r+4)j
    const double ij = shfl_d(isrc, jl | c);        // row j of the (scaled) inverse so far
    if (c > j) {
      e0 = (piv * e0 - t0 * u) * rprev;
      e1 = (piv * e1 - t1 * u) * rprev;
    }
    if (r > j) i0 = (piv * i0 - t0 * ij) * rprev;      // rows <= j of the inverse are final
    if (r + 4 > j) i1 = (piv * i1 - t1 * ij) * rprev;
    rprev = fast_rcp(piv);  // for the next step: off the chain
  }
  // lane (r, c) now holds b_rc at step c and J_rc = B_(r-1) x (unit-lower inverse)_rc.  With
  // s_k = 1 / sqrt(B_(k-1) B_k):  L_rc = b_rc s_c,  inv(L)_rc = J_rc s_r.
  const int dl = ((c & 3) << 3) | c;  // lane of element (c, c); its slot is c >> 2
  const double d0 = shfl_d(e0, dl), d1 = shfl_d(e1, dl);
  const double Bc = (c < 4) ? d0 : d1;  // B_c
  const int dlp = (((c + 7) & 3) << 3) | ((c + 7) & 7);  // element (c - 1, c - 1)
  const double q0 = shfl_d(e0, dlp), q1 = shfl_d(e1, dlp);
  const double Bp = (c == 0) ? 1.0 : ((c - 1 < 4) ? q0 : q1);  // B_(c-1)
  const bool bad = __any_sync(0xffffffffu, !(Bc > 0.0));
  const double l0 = e0, l1 = e1;
  const double rs = rsqrt(Bp) * rsqrt(Bc);  // s_c (two factors: the product could overflow)
  const double rs0 = shfl_d(rs, r), rs1 = shfl_d(rs, r + 4);  // lanes 0..7 hold columns 0..7
  if (r >= c) Cs[(k1 + r) * kCS + k1 + c] = l0 * rs;
  Cs[(k1 + r + 4) * kCS + k1 + c] = l1 * rs;  // r + 4 >= c for the rows that matter; the
                                              // entries above the diagonal are never read
  iv[8 * r + c] = i0 * rs0;
  iv[8 * (r + 4) + c] = i1 * rs1;
  return bad;
}

// ------------------------------------------------------------------------------------------
// In-place factorisation of the 64x64 tile in Cs (lower triangle valid): L overwrites the lower
// triangle, the inverses of the four 16x16 diagonal sub-blocks go to the diagonal blocks of Tm
// (written only at the very end).  Eight rounds: 8x8 factor + inverse (warp 0) -> panel
// X = A inv(L_8)^T and trailing update, both on the tensor cores.  All 256 threads call it.
// Iv: 12 x 64 doubles of scratch (the 8x8 inverses and the 16x16 assembly).
// Xd (optional): the tile still lacks the update C -= Xd Xd^T in its column blocks >= 2 (the
// caller applied column blocks 0 and 1); column block b + 2 gets it in round b, by the warps that
// wait for warp 0's factor step anyway, so that only 15 of the 36 blocks of that update are on
// the walker's dependency chain.  Xd (stride kCS) may be Tm.
// ------------------------------------------------------------------------------------------
// 64x64 tile global (leading dimension ld) -> shared (stride kCS), asynchronously
__device__ __forceinline__ void tile_prefetch(double* dst, const double* src, int ld) {
  for (int p = threadIdx.x; p < NB * 32; p += 256) {
    const int r = p >> 5, s = p & 31;
    cp_async16(dst + r * kCS + 2 * s, src + (size_t)r * ld + 2 * s);
  }
  cp_async_commit();
}

// Optional hook: while the block is being factored the walker wants the next tile it needs
// (pre-accumulated by the helpers) in shared memory as early as possible; thread 0 polls the
// tile's flag once per round and the CTA issues the cp.async prefetch as soon as it is up.
struct PrefetchHook {
  const int* flag = nullptr;  // nullptr: nothing to prefetch
  double* dst = nullptr;
  const double* src = nullptr;
  int ld = 0;
  bool issued = false;
  // a second flag that is only WATCHED (the tile that follows into the same buffer later):
  // *seen2 is set once it is up, so that the caller need not poll global memory on its own path
  const int* flag2 = nullptr;
  int* seen2 = nullptr;
};

// tr (optional): trace slots 8..14 of the diagonal tile (after the first 8x8 factor, after the
// panel / the update of round 0, after rounds 3 and 6, after the 8x8 inverses, at the end)
__device__ __noinline__ void potrf64(double* Cs, double* Tm, double* Iv, const double* Xd,
                                     int* s_bad, PrefetchHook* hook,
                                     unsigned long long* tr = nullptr) {
  __shared__ int s_hook;
  const int tid = threadIdx.x, lane = tid & 31;
  // warp index broadcast from lane 0: the compiler then knows the warp-specialised branches
  // below are warp-uniform and emits plain SHFL instead of WARPSYNC.COLLECTIVE sequences
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = lane >> 2, q = lane & 3;  // mma fragment coordinates
  if (warp == 0) {
    const bool bad = factor8(Cs, 0, Iv, lane);
    if (bad && lane == 0) *s_bad = 1;
  }
  const bool want = hook->flag != nullptr && !hook->issued;
  const bool watch2 = hook->flag2 != nullptr;
  int poll_kind = 0, poll_val = 0;  // (thread 255) the poll in flight: 1 = flag, 2 = flag2
  if (tid == 32) {  // (warp 1 idles during the first 8x8 factor)
    s_hook = want ? (ld_relaxed(hook->flag) != 0) : 0;
    if (s_hook) fence_acquire();
  }
  __syncthreads();
  if (tr && tid == 0) tr[8] = global_ns();
#pragma unroll 1
  for (int bk = 0; bk < 7; ++bk) {
    const int k1 = 8 * bk;
    const int below = NB - (k1 + 8);
    if (want && !hook->issued && s_hook) {
      tile_prefetch(hook->dst, hook->src, hook->ld);
      hook->issued = true;
    }
    // panel: X = A inv(L_8)^T, one 8x8 block of rows per warp (two m8n8k4 steps, in place)
    if (warp < (below >> 3)) {
      double* ab = Cs + (k1 + 8 + 8 * warp + g) * kCS + k1;
      const double* ib = Iv + 64 * bk + 8 * g;
      const double a0 = ab[q], a1 = ab[4 + q];
      const double b0 = ib[q], b1 = ib[4 + q];
      double x0 = 0.0, x1 = 0.0;
      dmma_m8n8k4(x0, x1, a0, b0);
      dmma_m8n8k4(x0, x1, a1, b1);
      __syncwarp();  // every lane has its operands: the block may be overwritten
      *reinterpret_cast<double2*>(ab + 2 * q) = make_double2(x0, x1);
    }
    __syncthreads();
    if (tr && tid == 0 && bk == 0) tr[9] = global_ns();
    // (s_hook is only written between the two barriers of a round and read after the second one,
    // so every thread takes the same prefetch decision)
    // Split-phase polling: the load issued in one round is looked at in the next, so its
    // ~800 cycles of global latency never sit between two barriers of a round (a warp that waits
    // for it there holds up all eight; in the late rounds the other warps have little else to do).
    if (tid == 255) {
      if (poll_kind == 1 && poll_val) {
        fence_acquire();
        s_hook = 1;
      } else if (poll_kind == 2 && poll_val) {
        fence_acquire();
        *hook->seen2 = 1;
      }
      poll_kind = 0;
      if (want && !hook->issued && !s_hook) {
        poll_val = ld_relaxed(hook->flag);
        poll_kind = 1;
      } else if (watch2 && !*hook->seen2) {
        poll_val = ld_relaxed(hook->flag2);
        poll_kind = 2;
      }
    }
    // trailing update inside the tile on the FP64 tensor cores: 8x8 output tiles (lower part),
    // C -= X_ti X_tj^T with k = 8 (two m8n8k4 steps).  Look-ahead: warp 0 updates the next
    // diagonal 8x8 block first and factors it right away, the other warps do the rest.
    const int nb = below >> 3;
    const int ntile = nb * (nb + 1) / 2;
    const double* X = Cs + (k1 + 8) * kCS + k1;
    // tile t of the lower triangle -> (ti, tj): a table instead of a square root (t < 28)
    auto tile_rc = [](int t, int& ti, int& tj) {
      const unsigned v = (unsigned)kTriRow[t];
      ti = (int)(v >> 4);
      tj = (int)(v & 15u);
    };
    auto update_tile = [&](int t) {
      int ti, tj;
      tile_rc(t, ti, tj);
      if (Xd && tj == 1 && ti >= 1) return;  // with the deferred update: heavy_tile below
      double* cp = Cs + (k1 + 8 + 8 * ti + g) * kCS + k1 + 8 + 8 * tj + 2 * q;
      double2 cv = *reinterpret_cast<double2*>(cp);
      const double a0 = -X[(8 * ti + g) * kCS + q], a1 = -X[(8 * ti + g) * kCS + 4 + q];
      const double b0 = X[(8 * tj + g) * kCS + q], b1 = X[(8 * tj + g) * kCS + 4 + q];
      dmma_m8n8k4(cv.x, cv.y, a0, b0);
      dmma_m8n8k4(cv.x, cv.y, a1, b1);
      *reinterpret_cast<double2*>(cp) = cv;
    };
    // tile (ti, 1) of the trailing part = block (bk + 1 + ti, bk + 2) of the 64x64 tile: this
    // round's update and the deferred C -= Xd Xd^T of column block bk + 2, one load and one store
    auto heavy_tile = [&](int ti) {
      double* cp = Cs + (k1 + 8 + 8 * ti + g) * kCS + k1 + 16 + 2 * q;
      double2 cv = *reinterpret_cast<double2*>(cp);
      // four accumulators: chains of 5, 5, 4 and 4 tensor-core steps instead of one of 18
      double2 c2 = make_double2(0.0, 0.0), c3 = c2, c4 = c2;
      const double a0 = -X[(8 * ti + g) * kCS + q], a1 = -X[(8 * ti + g) * kCS + 4 + q];
      const double b0 = X[(8 + g) * kCS + q], b1 = X[(8 + g) * kCS + 4 + q];
      dmma_m8n8k4(cv.x, cv.y, a0, b0);
      dmma_m8n8k4(c2.x, c2.y, a1, b1);
      const double* xa = Xd + (8 * (bk + 1 + ti) + g) * kCS + q;
      const double* xb = Xd + (8 * (bk + 2) + g) * kCS + q;
#pragma unroll
      for (int kk = 0; kk < NB; kk += 16) {
        dmma_m8n8k4(cv.x, cv.y, -xa[kk], xb[kk]);
        dmma_m8n8k4(c2.x, c2.y, -xa[kk + 4], xb[kk + 4]);
        dmma_m8n8k4(c3.x, c3.y, -xa[kk + 8], xb[kk + 8]);
        dmma_m8n8k4(c4.x, c4.y, -xa[kk + 12], xb[kk + 12]);
      }
      cv.x += (c2.x + c3.x) + c4.x;
      cv.y += (c2.y + c3.y) + c4.y;
      *reinterpret_cast<double2*>(cp) = cv;
    };
    if (warp == 0) {
      {  // tile (0, 0): the next diagonal 8x8 block
        double* cp = Cs + (k1 + 8 + g) * kCS + k1 + 8 + 2 * q;
        double2 cv = *reinterpret_cast<double2*>(cp);
        const double a0 = X[g * kCS + q], a1 = X[g * kCS + 4 + q];
        dmma_m8n8k4(cv.x, cv.y, -a0, a0);
        dmma_m8n8k4(cv.x, cv.y, -a1, a1);
        *reinterpret_cast<double2*>(cp) = cv;
      }
      __syncwarp();
      const bool bad = factor8(Cs, k1 + 8, Iv + 64 * (bk + 1), lane);
      if (bad && lane == 0) *s_bad = 1;
    } else {
      if (Xd && warp < nb) heavy_tile(warp);  // ti = 1 .. nb - 1, one per warp (the long one first)
      for (int t = warp; t < ntile; t += 7) update_tile(t);
    }
    __syncthreads();
    if (tr && tid == 0 && (bk == 0 || bk == 3 || bk == 6)) tr[bk == 0 ? 10 : (bk == 3 ? 11 : 12)] = global_ns();
  }
  if (want && !hook->issued && s_hook) {
    tile_prefetch(hook->dst, hook->src, hook->ld);
    hook->issued = true;
  }
  // Inverses of the 16x16 diagonal sub-blocks (what the triangular solves of the tiles below
  // use) from the 8x8 inverses of the factor steps:
  // inv([A 0; B C]) = [inv A, 0; -inv(C) B inv(A), inv C].  (Xd, if it is Tm, is dead by now.)
  if (tr && tid == 0) tr[13] = global_ns();
  {
    const int blk = tid >> 6, r = (tid >> 3) & 7, c = tid & 7, o = 16 * blk;
    const double* iA = Iv + 64 * (2 * blk);
    const double* iC = Iv + 64 * (2 * blk + 1);
    double* Ts = Iv + 64 * (8 + blk);
    double acc = 0.0;  // T = B inv(A)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += Cs[(o + 8 + r) * kCS + o + k] * iA[8 * k + c];
    Ts[8 * r + c] = acc;
    __syncthreads();
    double res = 0.0;  // - inv(C) T
#pragma unroll
    for (int k = 0; k < 8; ++k) res -= iC[8 * r + k] * Ts[8 * k + c];
    Tm[(o + r) * kCS + o + c] = iA[8 * r + c];
    Tm[(o + r) * kCS + o + 8 + c] = 0.0;
    Tm[(o + 8 + r) * kCS + o + c] = res;
    Tm[(o + 8 + r) * kCS + o + 8 + c] = iC[8 * r + c];
  }
  __syncthreads();
  if (tr && tid == 0) tr[14] = global_ns();
}

// ------------------------------------------------------------------------------------------
// X = C L^-T for a 64x64 tile, warp-local (warp w owns rows 8w..8w+7), in four 16-column steps:
//   X_b = (C_b - sum_{b'<b} X_b' L_bb'^T) inv(L_bb)^T.
// acc[nb] holds the tile's C fragments on entry and X's on exit; Lp = Lpack in shared memory;
// Xs receives X (row-major, stride kCS) as it is produced.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void trsm_core(double (&acc)[8][2], double* Xs, const double* Lp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int row = warp * 8 + g;
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
    acc[2 * b][0] = x0[0]; acc[2 * b][1] = x0[1];
    acc[2 * b + 1][0] = x1[0]; acc[2 * b + 1][1] = x1[1];
  }
}

// fragments <-> global / shared tiles (warp w rows 8w..8w+7; fragment nb = columns 8nb+2q, +1)
__device__ __forceinline__ void frag_store_global(const double (&acc)[8][2], double* tile, int ld) {
  const int lane = threadIdx.x & 31, row = (threadIdx.x >> 5) * 8 + (lane >> 2), q = lane & 3;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb)
    *reinterpret_cast<double2*>(tile + (size_t)row * ld + 8 * nb + 2 * q) =
        make_double2(acc[nb][0], acc[nb][1]);
}
__device__ __forceinline__ void frag_load_smem(double (&acc)[8][2], const double* tile) {
  const int lane = threadIdx.x & 31, row = (threadIdx.x >> 5) * 8 + (lane >> 2), q = lane & 3;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const double2 v = *reinterpret_cast<const double2*>(tile + row * kCS + 8 * nb + 2 * q);
    acc[nb][0] = v.x;
    acc[nb][1] = v.y;
  }
}

// ------------------------------------------------------------------------------------------
// The walker: diagonal and sub-diagonal tiles, one column after the other.
// Shared memory: Cs (current diagonal tile -> L -> Lpack), Tm (block inverses, then X of the
// sub-diagonal tile), Ps (prefetch buffer for the pre-accumulated tiles the helpers hand over;
// the next diagonal tile is updated in place there and the two buffers swap roles), Iv (scratch
// of the diagonal factorisation).
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void walker(double* __restrict__ A, int ld, int n, double* smem, const Work& w, int T,
                       int ncols, int* __restrict__ status,
                       unsigned long long* __restrict__ trace) {
  // Cs / Ps swap roles every column: one parity bit instead of two more live pointers
  int par = 0;
#define Cs (smem + (par ? 2 * 64 * kCS : 0))
#define Ps (smem + (par ? 0 : 2 * 64 * kCS))
  double* const Tm = smem + 64 * kCS;
  double* const Iv = smem + kWalkerInv;
  bool deferred = false;  // Cs still lacks C -= X X^T (X in Tm) in its column blocks >= 2
  __shared__ double rhs_row[NB];
  __shared__ int s_bad, s_poll, s_seen2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  if (tid == 0) s_bad = 0;
  int pending_diag = -1, pending_tile = -1;  // flags to raise (by warp 1) once their stores are ordered
  // C_00
  if (tid == 0) wait_flag(w.pre + 0);
  __syncthreads();
  tile_prefetch(Cs, A, ld);
  cp_async_wait<0>();
  __syncthreads();
  for (int j = 0; j < ncols; ++j) {
    const int k0 = j * NB;
    const int kb = min(NB, n - k0);
    const bool has_row = j + 1 < T;      // a tile row below (it may be the right-hand-side row only)
    const bool has_next = j + 1 < ncols; // a next diagonal block
    if (trace && tid == 0) trace[16 * (size_t)(j * T + j) + 1] = global_ns();
    // the pre-accumulated sub-diagonal tile is prefetched into Ps while the diagonal block is
    // factored, as soon as the helpers have it ready
    PrefetchHook hook;
    if (has_row) {
      hook.flag = w.pre + (j + 1) * T + j;
      hook.dst = Ps;
      hook.src = A + (size_t)(k0 + NB) * ld + k0;
      hook.ld = ld;
      if (has_next) {  // watched during the factorisation, fetched after the triangular solve
        hook.flag2 = w.pre + (j + 1) * T + (j + 1);
        hook.seen2 = &s_seen2;
      }
    }
    if (tid == 0) s_seen2 = 0;  // (ordered before its use by the barriers inside potrf64)
    // raise the flags of the previous column now: warp 1 idles during the first 8x8 factor anyway
    // (one fence, then relaxed stores: a release pattern without a second and third membar)
    if (warp == 1 && lane == 0 && (pending_diag >= 0 || pending_tile >= 0)) {
      __threadfence();
      if (pending_diag >= 0) st_relaxed(w.tile + pending_diag, 1);
      if (pending_tile >= 0) st_relaxed(w.tile + pending_tile, 1);
    }
    pending_diag = pending_tile = -1;
    if (kb < NB) {
      // last, partial block: keep the right-hand-side row aside, pad with the identity
      if (tid < NB) rhs_row[tid] = (tid < kb) ? Cs[kb * kCS + tid] : 0.0;
      __syncthreads();
      for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        if (r >= kb || c >= kb) Cs[r * kCS + c] = (r == c) ? 1.0 : 0.0;
      }
      __syncthreads();
    }
    potrf64(Cs, Tm, Iv, deferred ? Tm : nullptr, &s_bad, &hook,
            trace ? trace + 16 * (size_t)(j * T + j) : nullptr);
    deferred = false;
    // Cs := Lpack (diagonal 16x16 blocks <- their inverses)
    for (int idx = tid; idx < 4 * 256; idx += 256) {
      const int b = idx >> 8, r = 16 * b + ((idx >> 4) & 15), c = 16 * b + (idx & 15);
      Cs[r * kCS + c] = Tm[r * kCS + c];
    }
    __syncthreads();
    if (trace && tid == 0) trace[16 * (size_t)(j * T + j) + 2] = global_ns();
    if (kb < NB) {
      // y = rhs L^-T for the right-hand-side row of the last block: column-oriented substitution
      // with the blocked factor (off-diagonal blocks = L, diagonal blocks = inverses)
      if (tid < NB) {
        // v = rhs (thread c owns entry c); four 16-blocks
        double v = rhs_row[tid];
        for (int b = 0; b < 4; ++b) {
          // x_b = v_b inv(L_bb)^T : x_c = sum_{p <= c, p in block} v_p I[c][p]
          __syncwarp();
          if ((tid >> 4) == b) rhs_row[tid] = v;
          __threadfence_block();
          asm volatile("bar.sync 1, 64;");
          double x = 0.0;
          if ((tid >> 4) == b)
            for (int p = 16 * b; p <= tid; ++p) x += rhs_row[p] * Cs[tid * kCS + p];
          asm volatile("bar.sync 1, 64;");
          if ((tid >> 4) == b) rhs_row[tid] = x;
          asm volatile("bar.sync 1, 64;");
          // v_c -= sum_{p in block b} x_p L[c][p] for the later blocks
          if ((tid >> 4) > b)
            for (int p = 16 * b; p < 16 * b + 16; ++p) v -= rhs_row[p] * Cs[tid * kCS + p];
          if ((tid >> 4) == b) v = x;
        }
        if (tid < kb) A[(size_t)n * ld + k0 + tid] = v;
      }
      __syncthreads();
    }
    // publish Lpack_j (the helpers' triangular solves and the back-substitution read it)
    {
      double* lpack = w.lpack + (size_t)j * NB * NB;
      for (int idx = tid; idx < NB * NB / 2; idx += 256) {
        const int r = idx >> 5, c = (idx & 31) * 2;
        *reinterpret_cast<double2*>(lpack + r * NB + c) =
            *reinterpret_cast<const double2*>(Cs + r * kCS + c);
      }
      pending_diag = j * T + j;
    }
    if (!has_row) {
      __syncthreads();
      break;
    }
    // sub-diagonal tile: X = C L^-T, kept in Tm for the update of the next diagonal tile
    if (!hook.issued) {
      if (tid == 0) wait_flag(w.pre + (j + 1) * T + j);
      __syncthreads();
      tile_prefetch(Ps, A + (size_t)(k0 + NB) * ld + k0, ld);
    }
    cp_async_wait<0>();
    __syncthreads();
    if (trace && tid == 0) trace[16 * (size_t)(j * T + j) + 4] = global_ns();
    // Lpack_j is needed by the helpers' triangular solves of this column, which head the chain
    // that produces the walker's inputs of the NEXT column: publish it now (costs warp 1 one
    // membar of latency inside the solve below) instead of at the top of the next iteration
    if (tid == 32 && pending_diag >= 0) {
      __threadfence();
      st_relaxed(w.tile + pending_diag, 1);
    }
    pending_diag = -1;
    double acc[8][2];
    frag_load_smem(acc, Ps);
    // the pre-accumulated next diagonal tile follows into Ps: start the copy before the solve if
    // the helpers are done with it, otherwise right after
    bool next_issued = false;
    if (has_next) {
      // (usually seen already by the watcher inside potrf64: no global load on this path then)
      if (tid == 0) {
        s_poll = s_seen2;
        if (!s_poll) {
          s_poll = ld_relaxed(w.pre + (j + 1) * T + (j + 1)) != 0;
          if (s_poll) fence_acquire();
        }
      }
      __syncthreads();  // (also: every warp has its fragments, Ps may be overwritten)
      if (s_poll) {
        tile_prefetch(Ps, A + (size_t)(k0 + NB) * ld + k0 + NB, ld);
        next_issued = true;
      }
    }
    trsm_core(acc, Tm, Cs);
    frag_store_global(acc, A + (size_t)(k0 + NB) * ld + k0, ld);
    pending_tile = (j + 1) * T + j;
    if (trace && tid == 0) trace[16 * (size_t)(j * T + j) + 5] = global_ns();
    if (!has_next) {
      __syncthreads();
      break;
    }
    if (!next_issued) {
      if (tid == 0) wait_flag(w.pre + (j + 1) * T + (j + 1));
      __syncthreads();
      tile_prefetch(Ps, A + (size_t)(k0 + NB) * ld + k0 + NB, ld);
    }
    // next diagonal tile: C -= X X^T (the one term the helpers could not pre-accumulate), in
    // place in Ps.  Only column blocks 0 and 1 (15 of the 36 lower 8x8 blocks) are needed before
    // its factorisation can start; the others are applied inside potrf64, off the dependency
    // chain.  A partial last block takes all of it now (its right-hand-side row is read first).
    cp_async_wait<0>();
    __syncthreads();  // X complete in Tm, pre-accumulated tile complete in Ps, Lpack stores issued
    if (trace && tid == 0) trace[16 * (size_t)(j * T + j) + 6] = global_ns();
    const bool defer = n - (k0 + NB) >= NB;
    const int nblk = defer ? 15 : 36;
#pragma unroll 1
    for (int e = warp; e < nblk; e += 8) {
      int bi, bj;
      if (defer) {
        bi = e < 8 ? e : e - 7;
        bj = e < 8 ? 0 : 1;
      } else {
        bi = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= e) ++bi;
        while (bi * (bi + 1) / 2 > e) --bi;
        bj = e - bi * (bi + 1) / 2;
      }
      double* cp = Ps + (8 * bi + g) * kCS + 8 * bj + 2 * q;
      double2 cv = *reinterpret_cast<const double2*>(cp);
      double2 c2 = make_double2(0.0, 0.0);
      const double* xa = Tm + (8 * bi + g) * kCS + q;
      const double* xb = Tm + (8 * bj + g) * kCS + q;
#pragma unroll
      for (int kk = 0; kk < NB; kk += 8) {
        dmma_m8n8k4(cv.x, cv.y, -xa[kk], xb[kk]);
        dmma_m8n8k4(c2.x, c2.y, -xa[kk + 4], xb[kk + 4]);
      }
      cv.x += c2.x;
      cv.y += c2.y;
      *reinterpret_cast<double2*>(cp) = cv;
    }
    __syncthreads();
    par ^= 1;  // the updated tile becomes the current one; the old one (Lpack_j, stored) is free
    deferred = defer;
    if (trace && tid == 0) trace[16 * (size_t)(j * T + j) + 7] = global_ns();
  }
#undef Cs
#undef Ps
  // flags of the last column
  if (tid == 0) {
    if (s_bad) atomicExch(status, 1);
    __threadfence();
    if (pending_diag >= 0) st_release(w.tile + pending_diag, 1);
    if (pending_tile >= 0) st_release(w.tile + pending_tile, 1);
  }
}

// ------------------------------------------------------------------------------------------
// The factorisation kernel: persistent CTAs (256 threads); the first one to start is the walker,
// the others pull tile tasks from a ticket.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
chol_factor_kernel(double* __restrict__ A, int ld, int n, double* __restrict__ work_base,
                   int* __restrict__ status, unsigned long long* __restrict__ trace) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_ticket, s_ready, s_role;
  const Work w = work_layout(work_base, n);
  const int T = ld / NB;
  const int ncols = (n + NB - 1) / NB;
  const int ntiles = ncols * T - ncols * (ncols - 1) / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int row = warp * 8 + g;

  if (tid == 0) {
    s_role = atomicCAS(w.flags + 2, 0, 1) == 0 ? 1 : 0;
    if (s_role) {
      st_release(w.flags + 3, (int)smid() + 1);
    } else {
      int ws;
      while ((ws = ld_relaxed(w.flags + 3)) == 0) __nanosleep(40);
      // Shares the walker's SM: leave it alone -- but only ONE CTA may ever leave.  When the
      // other SMs are busy (another stream, another process) the block scheduler refills the
      // slot a leaver frees on this very SM, and if every CTA that lands here left, the grid
      // would drain through that slot and the walker would wait for helpers that no longer
      // exist.  The grid has at least walker + 2 CTAs, so at least one helper always remains;
      // a later CTA that lands on the walker's SM stays and works (slower, never stuck).
      if (ws == (int)smid() + 1 && atomicCAS(w.flags + 2, 1, 2) == 1) s_role = 2;
    }
  }
  __syncthreads();
  if (s_role == 2) return;
  if (s_role == 1) {
    walker(A, ld, n, smem, w, T, ncols, status, trace);
    return;
  }

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
    if (trace && tid == 0) trace[16 * (size_t)(i * T + j)] = global_ns();
    const bool diag = (i == j);
    const bool pre_only = (i - j) <= 1;  // finished by the walker
    // terms the task accumulates: k < j, except that the last one of a diagonal tile needs the
    // sub-diagonal tile of the previous column, which the walker produces and applies itself
    const int nk = diag ? (j > 0 ? j - 1 : 0) : j;

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
    const int nchunks = 2 * nk;
    int issued = 0, known_k = 0;  // chunks issued; operand tiles of columns < known_k are ready
    auto issue = [&](int c) {
      const int k = c >> 1, half = c & 1, st = c % kStages;
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
      // keep up to kStages chunks in flight; block on a flag only when there is nothing to compute
      while (issued < nchunks && issued < c + kStages) {
        const int k = issued >> 1;
        if (k >= known_k) {
          const bool must = (issued == c);
          if (tid == 0) {
            const int* fi = w.tile + i * T + k;
            const int* fj = w.tile + j * T + k;
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
      const int in_flight = issued - c - 1;  // groups younger than chunk c
      if (in_flight >= 2) cp_async_wait<2>();
      else if (in_flight == 1) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncthreads();
      const int st = c % kStages;
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
    if (trace && tid == 0) trace[16 * (size_t)(i * T + j) + 3] = global_ns();
    if (pre_only) {
      // hand the pre-accumulated tile to the walker (in place)
      if (nchunks > 0) frag_store_global(acc, A + (size_t)(i * NB) * ld + j * NB, ld);
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        st_release(w.pre + i * T + j, 1);
      }
      continue;
    }
    // X = C L_jj^-T with the published Lpack_j
    double* Xs = smem;             // [64][kCS]
    double* Lp = smem + 64 * kCS;  // [64][kCS]
    if (tid == 0) wait_flag(w.tile + j * T + j);
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
    trsm_core(acc, Xs, Lp);
    frag_store_global(acc, A + (size_t)(i * NB) * ld + j * NB, ld);
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      st_release(w.tile + i * T + j, 1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward substitution L^T x = y as a dataflow kernel: one CTA per 64-block (descending by
// ticket).  CTA j streams the tiles L[b][j], b > j, through a cp.async double buffer, applies
// y_j -= L[b][j]^T x_b as soon as x_b is published, then x_j = L_jj^-T y_j with the dense inverse
// it assembled from the blocked factor Lpack_j while it was waiting.
// x is its own flag: the host fills it with an all-ones NaN pattern, a consumer spins on the 64
// values it needs, and a producer never stores that pattern (NaNs are made canonical first).  The
// chain x_(j+1) -> x_j is then one global round trip per step — no flag store behind a fence and no
// second dependent load (2.4 -> 1.4 us per step).
// ------------------------------------------------------------------------------------------
constexpr int kBackSmem = 4 * NB * NB * (int)sizeof(double);

__global__ void __launch_bounds__(256)
chol_backsolve_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ work_base,
                      double* __restrict__ x) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double yj[NB], xb[NB], red[4][NB];
  __shared__ int s_ticket;
  const Work w = work_layout(work_base, n);
  const int nblk = (n + NB - 1) / NB;
  const int tid = threadIdx.x, c = tid & 63, part = tid >> 6;
  if (tid == 0) s_ticket = atomicAdd(w.flags + 1, 1);
  __syncthreads();
  const int j = nblk - 1 - s_ticket;
  if (j < 0) return;
  const int k0 = j * NB;
  const int kb = min(NB, n - k0);
  double* Lp = smem;                   // [64][64] Lpack_j
  double* Li = smem + NB * NB;         // [64][64] dense L_jj^-1, built while waiting for x
  double* stage0 = smem + 2 * NB * NB; // two tile stages
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
    const double* src = w.lpack + (size_t)j * NB * NB;
    for (int p = tid; p < NB * 32; p += 256) cp_async16(Lp + 2 * p, src + 2 * p);
    cp_async_commit();
  }
  if (tid < NB) yj[tid] = (tid < kb) ? ld_cg(A + (size_t)n * ld + k0 + tid) : 0.0;
  int issued_b = nblk - 1;  // next tile to issue
  for (int cnt = 0; cnt < 2 && issued_b > j; ++cnt, --issued_b) issue_tile(issued_b, (nblk - 1 - issued_b) & 1);
  // Dense inverse of the diagonal block from the blocked factor (diagonal 16x16 blocks of Lpack
  // are the inverses I_b, off-diagonal blocks are L), by block forward substitution
  //   Linv[bi][bj] = -I_bi * sum_{bk = bj}^{bi-1} L[bi][bk] Linv[bk][bj].
  // Every CTA but the first one of the chain has to wait for x anyway, so this is free.
  if (issued_b < nblk - 2) cp_async_wait<2>();       // Lpack is the oldest group
  else if (issued_b < nblk - 1) cp_async_wait<1>();
  else cp_async_wait<0>();
  __syncthreads();
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 6, cc = idx & 63;
    Li[idx] = ((r >> 4) == (cc >> 4)) ? Lp[idx] : 0.0;
  }
  __syncthreads();
#pragma unroll 1
  for (int dist = 1; dist < 4; ++dist) {
    const int nb = 4 - dist;  // blocks (bi = bj + dist, bj)
    double tmp[3];
    int cnt = 0;
    for (int idx = tid; idx < nb * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, cc = idx & 15;
      double acc = 0.0;
      for (int p2 = 16 * bj; p2 < 16 * bi; ++p2) acc += Lp[(16 * bi + r) * NB + p2] * Li[p2 * NB + 16 * bj + cc];
      tmp[cnt] = acc;
    }
    // tmp -> scratch above the diagonal of Li (unused otherwise): Li[bj-rows][bi-cols]
    cnt = 0;
    for (int idx = tid; idx < nb * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, cc = idx & 15;
      Li[(16 * bj + r) * NB + 16 * bi + cc] = tmp[cnt];
    }
    __syncthreads();
    cnt = 0;
    for (int idx = tid; idx < nb * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, cc = idx & 15;
      double acc = 0.0;
#pragma unroll
      for (int p2 = 0; p2 < 16; ++p2)
        acc += Lp[(16 * bi + r) * NB + 16 * bi + p2] * Li[(16 * bj + p2) * NB + 16 * bi + cc];
      tmp[cnt] = -acc;
    }
    __syncthreads();
    cnt = 0;
    for (int idx = tid; idx < nb * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, cc = idx & 15;
      Li[(16 * bi + r) * NB + 16 * bj + cc] = tmp[cnt];
    }
    __syncthreads();
  }
  for (int b = nblk - 1; b > j; --b) {
    const int st = (nblk - 1 - b) & 1;
    const int kbb = min(NB, n - b * NB);
    if (tid < NB) {
      double v = 0.0;
      if (tid < kbb) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(x + b * NB + tid);
        unsigned long long bits;
        while ((bits = ld_relaxed_u64(src)) == kXPending) {
        }
        v = __longlong_as_double((long long)bits);
      }
      xb[tid] = v;
    }
    // tile b has been issued; at most one younger group is in flight
    if (issued_b < b - 1) cp_async_wait<1>(); else cp_async_wait<0>();
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
    for (int r = part; r < NB; r += 4)
      if (r >= c) s += Li[r * NB + c] * yj[r];  // (L^-1)^T y: lower triangle only (the upper
                                                 // off-diagonal blocks hold scratch)
    red[part][c] = s;
  }
  __syncthreads();
  if (tid < kb) {
    double v = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
    if (v != v) v = __longlong_as_double(0x7ff8000000000000ll);  // never the "pending" pattern
    st_relaxed_u64(reinterpret_cast<unsigned long long*>(x + k0 + tid),
                   (unsigned long long)__double_as_longlong(v));
  }
}

}  // namespace

int chol_ld(int n) { return ((n + 1 + NB - 1) / NB) * NB; }
size_t chol_work_doubles(int n) {
  const size_t nblk = (size_t)(n + NB - 1) / NB;
  return nblk * NB * NB + (work_flag_ints(n) + 1) / 2 + 2;
}

// Factor + solve.  A: ld x ld (see header).  x: n doubles (device).  status: device int, set to
// 1 if a non-positive pivot was met.  Asynchronous on `s`; returns the number of launches.
int chol_solve_bordered(double* A, int n, int ld, double* x, double* work, int* status,
                        cudaStream_t s) {
  static PerDevice<> per_device;
  const auto& dev = per_device.get([](const DeviceFacts&, int&) {
    cudaError_t e = cudaFuncSetAttribute(chol_factor_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kFactorSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(chol_backsolve_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, kBackSmem);
  });
  const int num_sms = dev.facts.num_sms;
  const Work w = work_layout(work, n);
  cudaMemsetAsync(status, 0, sizeof(int), s);
  cudaMemsetAsync(w.flags, 0, sizeof(int) * work_flag_ints(n), s);
  cudaMemsetAsync(x, 0xff, sizeof(double) * (size_t)n, s);  // "pending" (see chol_backsolve_kernel)
  const int T = ld / NB;
  const int ncols = (n + NB - 1) / NB;
  const int ntiles = ncols * T - ncols * (ncols - 1) / 2;
  int grid = 2 * num_sms;
  if (grid > ntiles + 2) grid = ntiles + 2;  // walker + (possibly) the CTA that leaves its SM
  static const char* trace_path = std::getenv("PPSFM_CHOL_TRACE");
  unsigned long long* trace = nullptr;
  if (trace_path) {
    cudaMalloc(&trace, sizeof(unsigned long long) * 16 * (size_t)T * T);
    cudaMemsetAsync(trace, 0, sizeof(unsigned long long) * 16 * (size_t)T * T, s);
  }
  chol_factor_kernel<<<grid, 256, kFactorSmem, s>>>(A, ld, n, work, status, trace);
  chol_backsolve_kernel<<<ncols, 256, kBackSmem, s>>>(A, ld, n, work, x);
  if (trace_path) {  // development aid: dump "i j t0..t15" (ns) per tile
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
