// dense_chol.cu — FP64 dense Cholesky solve of the reduced camera system S dc = rhs.
//
// In the reference this step is hidden inside ceres::Solve (SPARSE_SCHUR / DENSE_SCHUR,
// src/optim/bundle_adjustment.cc:275-286).  The reduced camera matrix of a BA problem with
// hundreds of cameras that all share points is dense, so it is factored densely:
// blocked right-looking Cholesky on the lower triangle, 64-wide panels,
//   panel kernel : every CTA re-factors the 64x64 diagonal block in shared memory (87 kflop,
//                  cheaper than a separate launch + sync) and solves its own 64-row tile
//   update kernel: trailing C_ij -= X_i X_j^T on 64x64 tiles with FP64 tensor-core MMA
//                  (mma.sync.m8n8k4.f64 — tcgen05 has no FP64 kind), the one genuinely dense
//                  contraction of the path
// The right-hand side rides along as an extra matrix row ("bordered" factorisation), which
// yields y = L^-1 rhs for free; the backward substitution L^T x = y runs block by block.
//
// Matrix layout: row-major, leading dimension ld (multiple of 64), rows [0, n) = S (lower
// triangle referenced), row n = rhs^T, rows (n, ld) zero padding.
#include "common.h"
#include "dense_chol.h"

namespace ppsfm {

constexpr int NB = 64;


// ------------------------------------------------------------------------------------------
// Diagonal-block kernel (1 CTA, 256 threads): factors the 64x64 diagonal block and inverts its
// Cholesky factor, so that the panel below becomes a plain matrix product X = A L^-T that runs
// on the FP64 tensor cores.  Blocked in 16-wide sub-blocks: only the 16x16 factorisations (one
// warp, rows in registers, shuffles) and their triangular inverses are sequential; panels and
// trailing updates inside the block use all 256 threads.  A partial last block is padded with
// the identity.  Outputs: L (lower) written back into A, L^-1 (lower, dense 64x64) into `linv`.
// ------------------------------------------------------------------------------------------
__device__ void diag_block(double* __restrict__ A, int ld, int n, int k0,
                           double* __restrict__ linv, int* __restrict__ status,
                           double* dyn_smem) {
  double (*D)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem);  // block -> L
  double (*Tm)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + NB * (NB + 1));  // L^-1
  double (*I16)[16][17] =  // inverses of the four 16x16 diagonal sub-blocks
      reinterpret_cast<double (*)[16][17]>(dyn_smem + 2 * NB * (NB + 1));
  __shared__ int ok;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kb = min(NB, n - k0);
  if (tid == 0) ok = 1;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    double v = 0.0;
    if (r < kb && c <= r) v = A[(size_t)(k0 + r) * ld + k0 + c];
    else if (r >= kb && c == r) v = 1.0;
    D[r][c] = v;
    Tm[r][c] = 0.0;
  }
  __syncthreads();
#pragma unroll 1
  for (int bk = 0; bk < 4; ++bk) {
    const int k1 = 16 * bk;
    {
      // (a)+(b) 16x16 Cholesky with its inverse: lane i (< 16) owns row i of the block (a[]) and
      // row i of the accumulated elimination transform (m[], starts as e_i).  Applying the
      // eliminations of step j (scale row j by 1/l_jj, subtract l_ij x row j from rows i > j) to
      // the identity yields L^-1 — no divisions, no second sequential pass.
      // ALL warps execute this redundantly (only warp 0 stores): inside a warp-specialised
      // branch every shuffle compiles to a ~30-cycle WARPSYNC.COLLECTIVE sequence; in uniform
      // control flow it is a plain SHFL.  The other warps would idle anyway.
      const int i = lane & 15;
      double a[16], m[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        a[c] = (lane < 16 && c <= i) ? D[k1 + i][k1 + c] : 0.0;
        m[c] = (c == i) ? 1.0 : 0.0;
      }
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double ajj = __shfl_sync(0xffffffffu, a[j], j);
        bad = bad || !(ajj > 0.0);
        ajj = (ajj > 0.0) ? ajj : 1.0;
        const double inv = rsqrt(ajj);
        const double lij = (i == j) ? ajj * inv : a[j] * inv;
        a[j] = (i >= j) ? lij : 0.0;
#pragma unroll
        for (int c = 0; c <= j; ++c) {
          const double mc = (i == j) ? m[c] * inv : m[c];
          const double mjc = __shfl_sync(0xffffffffu, mc, j);
          m[c] = (i > j) ? mc - a[j] * mjc : mc;
        }
#pragma unroll
        for (int c = j + 1; c < 16; ++c) {
          const double lcj = __shfl_sync(0xffffffffu, a[j], c);
          a[c] = (i >= c) ? a[c] - a[j] * lcj : a[c];
        }
      }
      __syncthreads();  // all warps are done reading the sub-block
      if (warp == 0) {
        if (bad && lane == 0) ok = 0;
        if (lane < 16) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (c <= i) D[k1 + i][k1 + c] = a[c];
            I16[bk][i][c] = (c <= i) ? m[c] : 0.0;
          }
        }
      }
    }
    __syncthreads();
    const int below = NB - (k1 + 16);  // rows under this sub-block
    if (below > 0) {
      // (c) panel: X[r][c] = sum_{p <= c} A[r][k1+p] * I16[c][p]
      double xv[3];
      int cnt = 0;
      for (int idx = tid; idx < below * 16; idx += 256, ++cnt) {
        const int r = k1 + 16 + idx / 16, c = idx % 16;
        double acc = 0.0;
#pragma unroll
        for (int p2 = 0; p2 < 16; ++p2) acc += D[r][k1 + p2] * I16[bk][c][p2];
        xv[cnt] = acc;
      }
      __syncthreads();
      cnt = 0;
      for (int idx = tid; idx < below * 16; idx += 256, ++cnt)
        D[k1 + 16 + idx / 16][k1 + idx % 16] = xv[cnt];
      __syncthreads();
      // (d) trailing update inside the block (lower part)
      for (int idx = tid; idx < below * below; idx += 256) {
        const int r = k1 + 16 + idx / below, c = k1 + 16 + idx % below;
        if (c > r) continue;
        double acc = 0.0;
#pragma unroll
        for (int p2 = 0; p2 < 16; ++p2) acc += D[r][k1 + p2] * D[c][k1 + p2];
        D[r][c] -= acc;
      }
      __syncthreads();
    }
  }
  // L^-1 by block forward substitution: Linv[bi][bj] = -I16[bi] * sum_{bk=bj}^{bi-1} L[bi][bk] Linv[bk][bj]
  for (int idx = tid; idx < 4 * 256; idx += 256) {
    const int b = idx >> 8, r = (idx >> 4) & 15, c = idx & 15;
    Tm[16 * b + r][16 * b + c] = I16[b][r][c];
  }
  __syncthreads();
#pragma unroll 1
  for (int dist = 1; dist < 4; ++dist) {
    const int nblk = 4 - dist;  // blocks (bi = bj + dist, bj)
    double tmp[3];
    int cnt = 0;
    for (int idx = tid; idx < nblk * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, c = idx & 15;
      double acc = 0.0;
      for (int p2 = 16 * bj; p2 < 16 * bi; ++p2) acc += D[16 * bi + r][p2] * Tm[p2][16 * bj + c];
      tmp[cnt] = acc;
    }
    // tmp -> scratch region above the diagonal of Tm (unused otherwise): Tm[bj-rows][bi-cols]
    cnt = 0;
    for (int idx = tid; idx < nblk * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, c = idx & 15;
      Tm[16 * bj + r][16 * bi + c] = tmp[cnt];
    }
    __syncthreads();
    cnt = 0;
    for (int idx = tid; idx < nblk * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, c = idx & 15;
      double acc = 0.0;
#pragma unroll
      for (int p2 = 0; p2 < 16; ++p2) acc += I16[bi][r][p2] * Tm[16 * bj + p2][16 * bi + c];
      tmp[cnt] = -acc;
    }
    __syncthreads();
    cnt = 0;
    for (int idx = tid; idx < nblk * 256; idx += 256, ++cnt) {
      const int bj = idx >> 8, bi = bj + dist, r = (idx >> 4) & 15, c = idx & 15;
      Tm[16 * bi + r][16 * bj + c] = tmp[cnt];
    }
    __syncthreads();
  }
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    if (r < kb && c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = D[r][c];
    linv[idx] = (c <= r) ? Tm[r][c] : 0.0;
  }
  __syncthreads();
  if (tid == 0 && !ok) atomicExch(status, 1);
}

__global__ void __launch_bounds__(256)
chol_diag_kernel(double* __restrict__ A, int ld, int n, int k0, double* __restrict__ linv,
                 int* __restrict__ status) {
  extern __shared__ __align__(16) double dyn_smem_diag[];
  diag_block(A, ld, n, k0, linv, status, dyn_smem_diag);
}

// Trailing update with FP64 tensor cores: C(ti, tj) -= X_ti X_tj^T for tiles ti >= tj below /
// right of the panel.  One CTA (8 warps) per 64x64 tile; warp w owns rows 8w..8w+7 of the tile and
// all 64 columns as eight m8n8k4 accumulators.
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// One 64x64 output tile per CTA (8 warps; warp w owns rows 8w..8w+7 as eight m8n8k4 accumulators):
//   kPanel == false: trailing update  C(ti, tj) -= X_ti X_tj^T  for tiles ti >= tj
//   kPanel == true : panel            X_t = A_t Linv^T           (A_t overwritten in place)
template <bool kPanel>
__global__ void __launch_bounds__(256)
chol_tile_kernel(double* __restrict__ A, int ld, int k0, int kb, int first_tile_row,
                 const double* __restrict__ linv, int n, double* __restrict__ linv_next,
                 int* __restrict__ status) {
  int ti, tj;
  if (kPanel) {
    ti = blockIdx.x;
    tj = 0;
  } else {
    const int t = blockIdx.x;
    ti = (int)floor((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    tj = t - ti * (ti + 1) / 2;
  }
  const int ri = first_tile_row + ti * NB, rj = first_tile_row + tj * NB;
  extern __shared__ __align__(16) double dyn_smem[];
  double (*Xi)[NB + 4] = reinterpret_cast<double (*)[NB + 4]>(dyn_smem);
  double (*Xj)[NB + 4] = reinterpret_cast<double (*)[NB + 4]>(dyn_smem + NB * (NB + 4));
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx >> 6, c = idx & 63;
    Xi[r][c] = (c < kb && ri + r < ld) ? A[(size_t)(ri + r) * ld + k0 + c] : 0.0;
    if (kPanel)
      Xj[r][c] = linv[idx];
    else
      Xj[r][c] = (c < kb) ? A[(size_t)(rj + r) * ld + k0 + c] : 0.0;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;  // mma fragment coordinates
  double acc[8][2];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
  const int row = warp * 8 + g;  // A fragment: a = A[row = g][k = q]
#pragma unroll 4
  for (int kk = 0; kk < NB; kk += 4) {
    const double a = Xi[row][kk + q];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const double b = Xj[nb * 8 + g][kk + q];  // B fragment (col-major k x n): B[k = q][n = g]
      dmma_m8n8k4(acc[nb][0], acc[nb][1], a, b);
    }
  }
  // C fragment: c0 = C[g][2q], c1 = C[g][2q+1]
  if (kPanel) {
    if (ri + row < ld) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int c = nb * 8 + 2 * q;
        double* dst = A + (size_t)(ri + row) * ld + k0 + c;
        if (c < kb) dst[0] = acc[nb][0];
        if (c + 1 < kb) dst[1] = acc[nb][1];
      }
    }
  } else {
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int c = nb * 8 + 2 * q;
      double* dst = A + (size_t)(ri + row) * ld + rj + c;
      if (ti != tj || c <= row) dst[0] -= acc[nb][0];
      if (ti != tj || c + 1 <= row) dst[1] -= acc[nb][1];
    }
    // Look-ahead: tile (0, 0) is the next diagonal block.  Its CTA factors (and inverts) it right
    // away, overlapped with the rest of this trailing update, so the sequential 64x64
    // factorisation leaves the critical path of the next step.
    if (blockIdx.x == 0 && linv_next != nullptr && first_tile_row < n) {
      __threadfence_block();
      __syncthreads();
      diag_block(A, ld, n, first_tile_row, linv_next, status, dyn_smem);
    }
  }
}

// Backward substitution L^T x = y, one launch per 64-block (descending).  Every CTA recomputes
// x_b = L_bb^-T y_b from the stored inverse (a 64x64 mat-vec); CTA j < b then applies
// y_j -= L[b][j]^T x_b on its own 64 columns, CTA b stores x_b.  y and x are distinct buffers.
__global__ void __launch_bounds__(256)
chol_backsolve_kernel(const double* __restrict__ A, int ld, int n, int b,
                      const double* __restrict__ linv_all, double* __restrict__ y,
                      double* __restrict__ x) {
  __shared__ double yb[NB], xb[NB], red[4][NB];
  const int tid = threadIdx.x, c = tid & 63, part = tid >> 6;
  const int k0 = b * NB;
  const int kb = min(NB, n - k0);
  const double* linv = linv_all + (size_t)b * NB * NB;
  if (tid < NB) yb[tid] = (tid < kb) ? y[k0 + tid] : 0.0;
  __syncthreads();
  {
    double s = 0.0;
    for (int r = part; r < NB; r += 4) s += linv[r * NB + c] * yb[r];  // (L^-1)^T y
    red[part][c] = s;
  }
  __syncthreads();
  if (tid < NB) xb[tid] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
  __syncthreads();
  const int j = blockIdx.x;
  if (j == b) {
    if (tid < kb) x[k0 + tid] = xb[tid];
    return;
  }
  double s = 0.0;
  for (int r = part; r < kb; r += 4) s += A[(size_t)(k0 + r) * ld + j * NB + c] * xb[r];
  red[part][c] = s;
  __syncthreads();
  if (part == 0) y[j * NB + c] -= red[0][c] + red[1][c] + red[2][c] + red[3][c];
}

__global__ void chol_extract_y_kernel(const double* __restrict__ A, int ld, int n,
                                      double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = A[(size_t)n * ld + i];
}

int chol_ld(int n) { return ((n + 1 + NB - 1) / NB) * NB; }
size_t chol_work_doubles(int n) {
  const size_t nblk = (size_t)(n + NB - 1) / NB;
  return nblk * NB * NB + nblk * NB + NB;  // one L^-1 per diagonal block + y
}

// Factor + solve.  A: ld x ld (see header).  x: n doubles (device).  status: device int, set to
// 1 if a non-positive pivot was met.  Asynchronous on `s`; returns the number of launches.
int chol_solve_bordered(double* A, int n, int ld, double* x, double* linv, int* status,
                        cudaStream_t s) {
  constexpr int kTileSmem = 2 * NB * (NB + 4) * (int)sizeof(double);
  constexpr int kDiagSmem = (2 * NB * (NB + 1) + 4 * 16 * 17) * (int)sizeof(double);
  constexpr int kUpdateSmem = kDiagSmem > kTileSmem ? kDiagSmem : kTileSmem;  // look-ahead reuse
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kDiagSmem);
    cudaFuncSetAttribute(chol_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kUpdateSmem);
    cudaFuncSetAttribute(chol_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kTileSmem);
    attr_set = true;
  }
  int launches = 0;
  cudaMemsetAsync(status, 0, sizeof(int), s);
  bool diag_done = false;  // diagonal block k already factored by the previous update kernel
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = (n - k0 < NB) ? (n - k0) : NB;
    const int r0 = k0 + kb;
    double* linv_k = linv + (size_t)(k0 / NB) * NB * NB;
    if (!diag_done) {
      chol_diag_kernel<<<1, 256, kDiagSmem, s>>>(A, ld, n, k0, linv_k, status);
      ++launches;
    }
    diag_done = false;
    // row tiles below the diagonal block (they include the rhs row); for the last, partial
    // block the remaining rows (rhs + padding) start unaligned and fit in one guarded tile
    const int tiles = (ld - r0 + NB - 1) / NB;
    if (tiles > 0) {
      chol_tile_kernel<true><<<tiles, 256, kTileSmem, s>>>(A, ld, k0, kb, r0, linv_k, n, nullptr,
                                                           status);
      ++launches;
    }
    if (r0 < ld && kb == NB) {
      const int nt = (ld - r0) / NB;
      const int ntiles = nt * (nt + 1) / 2;
      if (ntiles > 0) {
        double* linv_next = (r0 < n) ? linv_k + NB * NB : nullptr;
        chol_tile_kernel<false><<<ntiles, 256, kUpdateSmem, s>>>(A, ld, k0, kb, r0, nullptr, n,
                                                                  linv_next, status);
        ++launches;
        diag_done = (linv_next != nullptr);
      }
    }
  }
  const int nblk = (n + NB - 1) / NB;
  double* y = linv + (size_t)nblk * NB * NB;  // y = L^-1 rhs (the bordered row), then consumed
  chol_extract_y_kernel<<<(n + 255) / 256, 256, 0, s>>>(A, ld, n, y);
  ++launches;
  for (int b = nblk - 1; b >= 0; --b) {
    chol_backsolve_kernel<<<b + 1, 256, 0, s>>>(A, ld, n, b, linv, y, x);
    ++launches;
  }
  return launches;
}

}  // namespace ppsfm
