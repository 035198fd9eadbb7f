// Latency micro-benchmark (single warp, dependent chains): nvcc -arch=sm_100a lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = seed; sm[lane + 32] = seed * 0.5;
  __syncthreads();
  double x = seed + lane * 1e-9, y = seed * 0.999;
  long long t0, t1;
  const int N = 256;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y + 1e-3;
  t1 = clock64(); if (lane == 0) cyc[0] = (t1 - t0) ; out[0] = x;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); if (lane == 0) cyc[1] = (t1 - t0); out[1] = x;
  // shfl double chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = shfl_d(x, (lane + 1) & 31);
  t1 = clock64(); if (lane == 0) cyc[2] = (t1 - t0); out[2] = x;
  // reciprocal chain
  x = seed + 1.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = 1.0 / x + 0.5;
  t1 = clock64(); if (lane == 0) cyc[3] = (t1 - t0); out[3] = x;
  // rsqrt chain
  x = seed + 1.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 0.5;
  t1 = clock64(); if (lane == 0) cyc[4] = (t1 - t0); out[4] = x;
  // LDS chain (pointer chase through values)
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = (int)sm[idx & 63] & 63;
  t1 = clock64(); if (lane == 0) cyc[5] = (t1 - t0); out[5] = idx;
  // DMMA chain
  double c0 = 0, c1 = 0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(x), "d"(y));
  t1 = clock64(); if (lane == 0) cyc[6] = (t1 - t0); out[6] = c0 + c1;
  // independent DMMA throughput (8 accumulators)
  double a[8][2] = {};
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(a[q][0]), "+d"(a[q][1]) : "d"(x), "d"(y));
  t1 = clock64(); if (lane == 0) cyc[7] = (t1 - t0); 
  double s = 0; for (int q = 0; q < 8; ++q) s += a[q][0] + a[q][1]; out[7] = s;
  // float rsqrt + cvt chain (rsqrtf seeded Newton)
  x = seed + 1.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double r = (double)rsqrtf((float)x); r = r * (1.5 - 0.5 * x * r * r); r = r * (1.5 - 0.5 * x * r * r); x = r + 0.5; }
  t1 = clock64(); if (lane == 0) cyc[8] = (t1 - t0); out[8] = x;
  // DSETP+select chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = (x > 0.7) ? x * y : 1.0;
  t1 = clock64(); if (lane == 0) cyc[9] = (t1 - t0); out[9] = x;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 16 * 8); cudaMalloc(&cyc, 16 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, 1.0000001);
  long long h[16]; cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA", "DMUL", "SHFL.f64", "1.0/x+add", "rsqrt+add", "LDS+cvt", "DMMA dep", "DMMA x8 indep (per 8)", "rsqrtf+2 Newton+add", "DSETP+sel+DMUL"};
  for (int i = 0; i < 10; ++i) printf("%-26s %.1f cycles/op\n", names[i], h[i] / 256.0);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
