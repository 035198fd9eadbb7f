// Cycle count of the walker's 8x8 factor step (factor8 of dense_chol.cu) and of one 8x8 tile
// update, single warp, back to back on a block in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o /tmp/f8 scripts/micro/factor8_probe.cu && /tmp/f8
#include <cstdio>

#include "../../privacy_preserving_sfm_b200/csrc/dense_chol.cu"

namespace ppsfm {
namespace {
__global__ void probe(long long* out, int iters) {
  __shared__ double Cs[64 * kCS];
  __shared__ double Iv[64];
  const int lane = threadIdx.x;
  for (int i = lane; i < 64 * kCS; i += 32) Cs[i] = 0.0;
  __syncwarp();
  long long t_f = 0, t_u = 0;
  for (int it = 0; it < iters; ++it) {
    // a fresh SPD 8x8 block
    for (int i = lane; i < 64; i += 32) {
      const int r = i >> 3, c = i & 7;
      Cs[r * kCS + c] = (r == c) ? 9.0 + r : 1.0 / (1.0 + r + c);
    }
    __syncwarp();
    long long c0 = clock64();
    const bool bad = factor8(Cs, 0, Iv, lane);
    __syncwarp();
    long long c1 = clock64();
    t_f += c1 - c0;
    if (bad) t_f += 1000000;
    // one tile update as the walker does it
    const int g = lane >> 2, q = lane & 3;
    c0 = clock64();
    double* cp = Cs + (8 + g) * kCS + 8 + 2 * q;
    double2 cv = *reinterpret_cast<double2*>(cp);
    const double a0 = Cs[(8 + g) * kCS + q], a1 = Cs[(8 + g) * kCS + 4 + q];
    dmma_m8n8k4(cv.x, cv.y, -a0, a0);
    dmma_m8n8k4(cv.x, cv.y, -a1, a1);
    *reinterpret_cast<double2*>(cp) = cv;
    __syncwarp();
    c1 = clock64();
    t_u += c1 - c0;
  }
  if (lane == 0) {
    out[0] = t_f / iters;
    out[1] = t_u / iters;
  }
}
}  // namespace
}  // namespace ppsfm

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  ppsfm::probe<<<1, 32>>>(d, 1000);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("factor8 %lld cycles, 8x8 tile update %lld cycles (%s)\n", h[0], h[1],
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
