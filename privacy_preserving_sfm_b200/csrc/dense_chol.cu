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

// Unblocked Cholesky of a 64x64 block held in shared memory (row-major, stride NB+1).
// 256 threads.  Returns via *ok (shared) whether all pivots were positive.
__device__ void factor_diag_smem(double (*D)[NB + 1], int kb, int* ok) {
  const int tid = threadIdx.x;
  for (int j = 0; j < kb; ++j) {
    __syncthreads();
    const double d = D[j][j];
    if (tid == 0 && !(d > 0.0)) *ok = 0;
    const double sd = sqrt(d > 0.0 ? d : 1.0);
    __syncthreads();
    // scale column j
    for (int i = j + tid; i < kb; i += blockDim.x) D[i][j] = (i == j) ? sd : D[i][j] / sd;
    __syncthreads();
    // rank-1 update of the trailing lower part
    const int rem = kb - j - 1;
    for (int idx = tid; idx < rem * rem; idx += blockDim.x) {
      const int r = j + 1 + idx / rem, c = j + 1 + idx % rem;
      if (c <= r) D[r][c] -= D[r][j] * D[c][j];
    }
  }
  __syncthreads();
}

// Panel step k: rows of tile `blockIdx.x + first_tile` (below the diagonal block), or the
// diagonal block itself for the CTA that owns it (blockIdx.x == 0 writes it back).
__global__ void __launch_bounds__(256)
chol_panel_kernel(double* __restrict__ A, int ld, int n, int k0, int* __restrict__ status) {
  extern __shared__ __align__(16) double dyn_smem[];
  double (*D)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem);
  double (*T)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dyn_smem + NB * (NB + 1));
  __shared__ int ok;
  const int tid = threadIdx.x;
  const int kb = min(NB, n - k0);
  if (tid == 0) ok = 1;
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    D[r][c] = (r < kb && c < kb && c <= r) ? A[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
  }
  __syncthreads();
  factor_diag_smem(D, kb, &ok);
  if (blockIdx.x == 0) {
    for (int idx = tid; idx < kb * kb; idx += blockDim.x) {
      const int r = idx / kb, c = idx % kb;
      if (c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = D[r][c];
    }
    if (tid == 0 && !ok) atomicExch(status, 1);
  }
  // my row tile: rows r0 .. r0+63 (all < ld), columns k0 .. k0+kb-1
  const int r0 = k0 + kb + blockIdx.x * NB;
  if (r0 >= ld) return;
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    T[r][c] = (c < kb && r0 + r < ld) ? A[(size_t)(r0 + r) * ld + k0 + c] : 0.0;
  }
  __syncthreads();
  // X L^T = T  -> forward substitution along the columns, 4 threads per row
  {
    const int r = tid >> 2, q = tid & 3;
    for (int j = 0; j < kb; ++j) {
      double s = 0.0;
      for (int p = q; p < j; p += 4) s += T[r][p] * D[j][p];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) T[r][j] = (T[r][j] - s) / D[j][j];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    if (c < kb && r0 + r < ld) A[(size_t)(r0 + r) * ld + k0 + c] = T[r][c];
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

// Backward substitution step for block k (descending): every CTA re-solves
// x_k = L_kk^-T y_k in shared memory; CTA j < k applies y_j -= L[k][j]^T x_k, CTA k stores x_k.
__global__ void __launch_bounds__(256)
chol_backsolve_kernel(const double* __restrict__ A, int ld, int n, int k0, double* __restrict__ y) {
  __shared__ double D[NB][NB + 1];
  __shared__ double xk[NB];
  const int tid = threadIdx.x;
  const int kb = min(NB, n - k0);
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx % NB;
    D[r][c] = (r < kb && c <= r) ? A[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
  }
  if (tid < NB) xk[tid] = (tid < kb) ? y[k0 + tid] : 0.0;
  __syncthreads();
  if (tid < 32) {  // one warp: sequential over rows from the bottom
    for (int i = kb - 1; i >= 0; --i) {
      double s = 0.0;
      for (int p = i + 1 + tid; p < kb; p += 32) s += D[p][i] * xk[p];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      if (tid == 0) xk[i] = (xk[i] - s) / D[i][i];
      __syncwarp();
    }
  }
  __syncthreads();
  const int j = blockIdx.x;  // tile column (0 .. k0/NB), the last one stores x_k
  if (j * NB == k0) {
    if (tid < kb) y[k0 + tid] = xk[tid];
    return;
  }
  // y_j[c] -= sum_r L[k0 + r][j*NB + c] * xk[r]
  const int c = tid & 63, part = tid >> 6;  // 4 partial sums per column
  double s = 0.0;
  for (int r = part; r < kb; r += 4) s += A[(size_t)(k0 + r) * ld + j * NB + c] * xk[r];
  __shared__ double red[4][NB];
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

// Factor + solve.  A: ld x ld (see header).  x: n doubles (device).  status: device int, set to
// 1 if a non-positive pivot was met.  Asynchronous on `s`; returns the number of launches.
int chol_solve_bordered(double* A, int n, int ld, double* x, int* status, cudaStream_t s) {
  constexpr int kPanelSmem = 2 * NB * (NB + 1) * (int)sizeof(double);
  constexpr int kUpdateSmem = 2 * NB * (NB + 4) * (int)sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kPanelSmem);
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
    chol_panel_kernel<<<tiles > 0 ? tiles : 1, 256, kPanelSmem, s>>>(A, ld, n, k0, status);
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
  const int nblk = (n + NB - 1) / NB;
  for (int b = nblk - 1; b >= 0; --b) {
    chol_backsolve_kernel<<<b + 1, 256, 0, s>>>(A, ld, n, b * NB, x);
    ++launches;
  }
  return launches;
}

}  // namespace ppsfm
