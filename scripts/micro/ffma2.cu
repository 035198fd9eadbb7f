// Throughput micro-benchmark of the packed float FMA forms on sm_100a (one CTA per SM, W warps per
// SMSP, 8 independent accumulator chains per thread):
//   nvcc -gencode arch=compute_100a,code=sm_100a ffma2.cu -o ffma2 && ./ffma2
// Reports cycles per warp instruction per SMSP for FFMA (scalar), FFMA2 with three packed
// operands, FFMA2 with a scalar-broadcast multiplicand (the form the score kernel uses) and DFMA.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 4096;
template <int MODE>
__global__ void k(float* out, long long* cyc, float s) {
  float2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(s + i, s - i + threadIdx.x);
  const float ax = s * 0.999f + threadIdx.x * 1e-8f;
  const float2 a = make_float2(ax, MODE == 2 ? ax : ax * 0.999f);  // registers
  const float2 b = make_float2(1e-3f + threadIdx.x * 1e-9f, 2e-3f - threadIdx.x * 1e-9f);  // registers
  double dacc[8];
  for (int i = 0; i < 8; ++i) dacc[i] = s + i;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < N; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) acc[i].x = fmaf(a.x, acc[i].x, b.x);
      if (MODE == 1 || MODE == 2) acc[i] = __ffma2_rn(a, acc[i], b);
      if (MODE == 3) dacc[i] = fma((double)s, dacc[i], (double)b.x);
    }
  }
  const long long t1 = clock64();
  float r = 0;
  for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y + (float)dacc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_smsp) {
  float* out; long long* cyc;
  const int threads = 128 * warps_per_smsp, blocks = 148;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
  k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  printf("%-28s warps/SMSP %d: %.2f cycles per warp instruction per SMSP\n", name, warps_per_smsp,
         avg / (double(N) * 8 * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 4, 8}) {
    run<0>("FFMA", w);
    run<1>("FFMA2 packed operands", w);
    run<2>("FFMA2 broadcast multiplicand", w);
    run<3>("DFMA", w);
  }
  return 0;
}
