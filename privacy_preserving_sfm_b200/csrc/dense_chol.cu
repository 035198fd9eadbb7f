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

// Panel step, 64 threads per CTA, thread r owns matrix row r in REGISTERS (fully unrolled):
//   1. every CTA re-factors the 64x64 diagonal block (left-looking; row j is broadcast from shared
//      memory, 4 independent accumulators hide the DFMA latency) — cheaper than a launch + sync;
//   2. X L^T = A for its own 64-row tile by per-row forward substitution (no synchronisation).
// A partial last block is padded with the identity.  CTA 0 writes the factored block back.
__global__ void __launch_bounds__(NB)
chol_panel_kernel(double* __restrict__ A, int ld, int n, int k0, int* __restrict__ status) {
  __shared__ double D[NB][NB + 1];
  __shared__ int ok;
  const int r = threadIdx.x;
  const int kb = min(NB, n - k0);
  const int r0 = k0 + kb + blockIdx.x * NB;
  if (r == 0) ok = 1;
  double drow[NB];
  {
    const double* src = A + (size_t)(k0 + r) * ld + k0;
#pragma unroll
    for (int c = 0; c < NB; ++c)
      drow[c] = (r < kb) ? ((c <= r) ? src[c] : 0.0) : ((c == r) ? 1.0 : 0.0);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int p = 0; p < j; ++p) acc[p & 3] += drow[p] * D[j][p];
    double sj = drow[j] - ((acc[0] + acc[1]) + (acc[2] + acc[3]));
    if (r == j) {
      if (!(sj > 0.0)) {
        ok = 0;
        sj = 1.0;
      }
      drow[j] = sqrt(sj);
      D[j][j] = drow[j];
    }
    __syncthreads();
    if (r > j) {
      drow[j] = sj / D[j][j];
      D[r][j] = drow[j];
    }
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    if (r < kb) {
      double* dst = A + (size_t)(k0 + r) * ld + k0;
#pragma unroll
      for (int c = 0; c < NB; ++c)
        if (c <= r) dst[c] = drow[c];
    }
    if (r == 0 && !ok) atomicExch(status, 1);
  }
  if (r0 >= ld) return;
  const bool live = r0 + r < ld;
  double trow[NB];
  {
    const double* src = A + (size_t)(r0 + (live ? r : 0)) * ld + k0;
#pragma unroll
    for (int c = 0; c < NB; ++c) trow[c] = (live && c < kb) ? src[c] : 0.0;
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int p = 0; p < j; ++p) acc[p & 3] += trow[p] * D[j][p];
    trow[j] = (trow[j] - ((acc[0] + acc[1]) + (acc[2] + acc[3]))) / D[j][j];
  }
  if (live) {
    double* dst = A + (size_t)(r0 + r) * ld + k0;
#pragma unroll
    for (int c = 0; c < NB; ++c)
      if (c < kb) dst[c] = trow[c];
  }
}

// Trailing update with FP64 tensor cores: C(ti, tj) -= X_ti X_tj^T for tiles ti >= tj below /
// right of the panel.  One CTA (8 warps) per 64x64 tile; warp w owns rows 8w..8w+7 of the tile and
// all 64 columns as eight m8n8k4 accumulators.
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
chol_update_kernel(double* __restrict__ A, int ld, int k0, int kb, int first_tile_row) {
  // linear tile index -> (ti, tj) with tj <= ti, both relative to first_tile_row
  const int t = blockIdx.x;
  int ti = (int)floor((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int ri = first_tile_row + ti * NB, rj = first_tile_row + tj * NB;
  extern __shared__ __align__(16) double dyn_smem[];
  double (*Xi)[NB + 4] = reinterpret_cast<double (*)[NB + 4]>(dyn_smem);
  double (*Xj)[NB + 4] = reinterpret_cast<double (*)[NB + 4]>(dyn_smem + NB * (NB + 4));
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    Xi[r][c] = (c < kb) ? A[(size_t)(ri + r) * ld + k0 + c] : 0.0;
    Xj[r][c] = (c < kb) ? A[(size_t)(rj + r) * ld + k0 + c] : 0.0;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;  // mma fragment coordinates
  double acc[8][2];
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
  const int row = warp * 8 + g;  // A fragment: a = A[row = g][k = q]
  for (int kk = 0; kk < NB; kk += 4) {
    const double a = Xi[row][kk + q];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const double b = Xj[nb * 8 + g][kk + q];  // B fragment (col-major k x n): B[k = q][n = g]
      dmma_m8n8k4(acc[nb][0], acc[nb][1], a, b);
    }
  }
  // C fragment: c0 = C[g][2q], c1 = C[g][2q+1]
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int c = nb * 8 + 2 * q;
    double* dst = A + (size_t)(ri + row) * ld + rj + c;
    if (ti != tj || c <= row) dst[0] -= acc[nb][0];
    if (ti != tj || c + 1 <= row) dst[1] -= acc[nb][1];
  }
}

// Backward substitution L^T x = y in ONE persistent CTA: for each 64-block from the bottom,
// warp 0 solves the diagonal block column by column (no reductions: after x_i is known the
// remaining entries are updated), then all 1024 threads apply y[0:k0] -= L[k0:k0+kb, 0:k0]^T x_k
// with coalesced reads of the L rows.
__global__ void __launch_bounds__(1024)
chol_backsolve_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ y) {
  __shared__ double D[NB][NB + 1];
  __shared__ double xk[NB];
  const int tid = threadIdx.x;
  const int nblk = (n + NB - 1) / NB;
  for (int b = nblk - 1; b >= 0; --b) {
    const int k0 = b * NB;
    const int kb = min(NB, n - k0);
    for (int idx = tid; idx < NB * NB; idx += 1024) {
      const int r = idx >> 6, c = idx & 63;
      D[r][c] = (r < kb && c <= r) ? A[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
    }
    if (tid < NB) xk[tid] = (tid < kb) ? y[k0 + tid] : 0.0;
    __syncthreads();
    if (tid < 32) {
      double v0 = xk[tid], v1 = xk[tid + 32];  // lane owns entries tid and tid + 32
      for (int i = kb - 1; i >= 0; --i) {
        const double mine = (i < 32) ? v0 : v1;
        const double xi = __shfl_sync(0xffffffffu, mine, i & 31) / D[i][i];
        if (tid == (i & 31)) { if (i < 32) v0 = xi; else v1 = xi; }
        if (tid < i) v0 -= D[i][tid] * xi;
        if (tid + 32 < i) v1 -= D[i][tid + 32] * xi;
      }
      xk[tid] = v0;
      xk[tid + 32] = v1;
    }
    __syncthreads();
    if (tid < kb) y[k0 + tid] = xk[tid];
    for (int c = tid; c < k0; c += 1024) {
      double s = 0.0;
      const double* col = A + (size_t)k0 * ld + c;
#pragma unroll 8
      for (int r = 0; r < kb; ++r) s += col[(size_t)r * ld] * xk[r];
      y[c] -= s;
    }
    __syncthreads();
  }
}

__global__ void chol_extract_y_kernel(const double* __restrict__ A, int ld, int n,
                                      double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = A[(size_t)n * ld + i];
}

int chol_ld(int n) { return ((n + 1 + NB - 1) / NB) * NB; }

// Factor + solve.  A: ld x ld (see header).  x: n doubles (device).  status: device int, set to
// 1 if a non-positive pivot was met.  Asynchronous on `s`; returns the number of launches.
int chol_solve_bordered(double* A, int n, int ld, double* x, int* status, cudaStream_t s) {
  constexpr int kUpdateSmem = 2 * NB * (NB + 4) * (int)sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kUpdateSmem);
    attr_set = true;
  }
  int launches = 0;
  cudaMemsetAsync(status, 0, sizeof(int), s);
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = (n - k0 < NB) ? (n - k0) : NB;
    const int r0 = k0 + kb;
    // row tiles below the diagonal block (they include the rhs row); for the last, partial
    // block the remaining rows (rhs + padding) start unaligned and fit in one guarded tile
    const int tiles = (ld - r0 + NB - 1) / NB;
    chol_panel_kernel<<<tiles > 0 ? tiles : 1, NB, 0, s>>>(A, ld, n, k0, status);
    ++launches;
    if (r0 < ld && kb == NB) {
      const int nt = (ld - r0) / NB;
      const int ntiles = nt * (nt + 1) / 2;
      if (ntiles > 0) {
        chol_update_kernel<<<ntiles, 256, kUpdateSmem, s>>>(A, ld, k0, kb, r0);
        ++launches;
      }
    }
  }
  chol_extract_y_kernel<<<(n + 255) / 256, 256, 0, s>>>(A, ld, n, x);
  ++launches;
  chol_backsolve_kernel<<<1, 1024, 0, s>>>(A, ld, n, x);
  ++launches;
  return launches;
}

}  // namespace ppsfm
