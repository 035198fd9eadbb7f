// Phase timing of the 64x64 diagonal-block factorisation pieces (clock64, one CTA alone).
#include "../../privacy_preserving_sfm_b200/csrc/dense_chol.cu"
#include <vector>
#include <cstdio>
using namespace ppsfm;
__global__ void __launch_bounds__(256) k_phases(const double* Ain, long long* cyc, double* out) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double rdiag[64];
  double* Cs = smem;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int idx = tid; idx < 64 * 64; idx += 256) Cs[(idx >> 6) * kCS + (idx & 63)] = Ain[idx];
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) factor16(Cs, 0, rdiag, lane);
  long long t1 = clock64();
  __syncthreads();
  long long t2 = clock64();
  // panel
  const int k1 = 0, below = 48;
  if (tid < below) {
    double* rowp = Cs + (k1 + 16 + tid) * kCS + k1;
    double x[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) x[c] = rowp[c];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      x[c] *= rdiag[k1 + c];
#pragma unroll
      for (int p = c + 1; p < 16; ++p) x[p] -= x[c] * Cs[(k1 + p) * kCS + k1 + c];
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) rowp[c] = x[c];
  }
  long long t3 = clock64();
  __syncthreads();
  long long t4 = clock64();
  const int cnt = below * (below + 1) / 2;
  for (int idx = tid; idx < cnt; idx += 256) {
    int r = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= idx) ++r;
    while (r * (r + 1) / 2 > idx) --r;
    const int c = idx - r * (r + 1) / 2;
    const double* xr = Cs + (k1 + 16 + r) * kCS + k1;
    const double* xc = Cs + (k1 + 16 + c) * kCS + k1;
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int p = 0; p < 16; p += 2) { acc0 += xr[p] * xc[p]; acc1 += xr[p + 1] * xc[p + 1]; }
    Cs[(k1 + 16 + r) * kCS + k1 + 16 + c] -= acc0 + acc1;
  }
  long long t5 = clock64();
  __syncthreads();
  long long t6 = clock64();
  if (warp < 4) invert16(Cs, 16 * warp, rdiag, smem + 64 * kCS + (16 * warp) * kCS + 16 * warp, lane);
  long long t7 = clock64();
  if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; cyc[3] = t7 - t6; cyc[4] = t2 - t1; cyc[5] = t4 - t3; }
  out[tid] = Cs[tid * 3];
}
int main() {
  const int n = 64;
  std::vector<double> A(n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[i * n + j] = (i == j ? n + 1.0 : 1.0 / (1 + abs(i - j)));
  double *dA, *dout; long long* dc;
  cudaMalloc(&dA, sizeof(double) * n * n); cudaMalloc(&dout, 8 * 256); cudaMalloc(&dc, 8 * 8);
  cudaMemcpy(dA, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_phases, cudaFuncAttributeMaxDynamicSharedMemorySize, kFactorSmem);
  for (int rep = 0; rep < 2; ++rep) k_phases<<<1, 256, kFactorSmem>>>(dA, dc, dout);
  long long h[8]; cudaMemcpy(h, dc, 64, cudaMemcpyDeviceToHost);
  printf("factor16 %lld | panel %lld | update %lld | invert16 %lld cycles (barriers %lld %lld) %s\n", h[0], h[1], h[2], h[3], h[4], h[5],
         cudaGetErrorString(cudaDeviceSynchronize()));
}
