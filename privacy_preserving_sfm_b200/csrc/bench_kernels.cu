// bench_kernels.cu — measurement helpers used by bench.py (not on the product path):
//   * FP64 issue-rate micro-benchmarks (the roofline denominator of the scoring kernel: FP64
//     peak is not in MEASURED_PEAKS.json, SURVEY.md §8d asks the builder to measure it)
//   * an L2 flush (writes a buffer larger than the 126 MB L2)
#include "common.h"

namespace {

template <bool kFused>
__global__ void __launch_bounds__(256) fp64_rate_kernel(double* out, int iters, double a,
                                                        double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    if (kFused) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    } else {
      x0 = __dmul_rn(x0, a); x1 = __dadd_rn(x1, b); x2 = __dmul_rn(x2, a); x3 = __dadd_rn(x3, b);
      x4 = __dmul_rn(x4, a); x5 = __dadd_rn(x5, b); x6 = __dmul_rn(x6, a); x7 = __dadd_rn(x7, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// Packed float FMA rate (FFMA2: two FMAs per lane and instruction — the form the float stage of
// the score filter is built from), 8 independent accumulator pairs per thread.
__global__ void __launch_bounds__(256) fp32_rate_kernel(float* out, int iters, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 av = make_float2(a + threadIdx.x * 1e-9f, a - threadIdx.x * 1e-9f);
  const float2 bv = make_float2(b + threadIdx.x * 1e-9f, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], av, bv);
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// HBM write / read micro-benchmarks: STREAM-style copy peaks mix reads and writes 1:1, but the
// Jacobian build writes 3x what it reads, so its ceiling is the write-side bandwidth.
__global__ void __launch_bounds__(256) hbm_write_kernel(double2* __restrict__ out, size_t n,
                                                        double v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = make_double2(v, v + (double)i);
}
__global__ void __launch_bounds__(256) hbm_read_kernel(const double2* __restrict__ in, size_t n,
                                                       double* __restrict__ sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 v = in[i];
    acc += v.x + v.y;
  }
  if (acc == 123.456) sink[0] = acc;
}

}  // namespace

extern "C" {

// Write-only and read-only HBM bandwidth in GB/s over a 2 GiB buffer (best of 3).
int ppsfm_bench_hbm_rw_peak(ppsfm_ctx* ctx, double* write_gbs, double* read_gbs) {
  if (!ctx) return PPSFM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const size_t bytes = (size_t)2 << 30, n = bytes / sizeof(double2);
  double2* buf = nullptr;
  double* sink = nullptr;
  PPSFM_CUDA(ctx, cudaMalloc(&buf, bytes));
  PPSFM_CUDA(ctx, cudaMalloc(&sink, 64));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best[2] = {0, 0};
  const int blocks = ctx->num_sms * 8;
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(a, ctx->stream);
      if (mode == 0) hbm_write_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, n, 1.0 + rep);
      else hbm_read_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, n, sink);
      cudaEventRecord(b, ctx->stream);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      const double gbs = bytes / (ms * 1e-3) / 1e9;
      if (rep > 0 && gbs > best[mode]) best[mode] = gbs;
    }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(buf);
  cudaFree(sink);
  PPSFM_CUDA(ctx, cudaGetLastError());
  if (write_gbs) *write_gbs = best[0];
  if (read_gbs) *read_gbs = best[1];
  return PPSFM_OK;
}

// Returns FP64 instruction throughput in 1e12 thread-instructions per second:
// *dfma_tips for DFMA (x2 = TFLOP/s), *dmuladd_tips for an unfused DMUL/DADD mix.
int ppsfm_bench_fp64_peak(ppsfm_ctx* ctx, double* dfma_tips, double* dmuladd_tips) {
  if (!ctx) return PPSFM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const int blocks = ctx->num_sms * 8, threads = 256, iters = 8192;
  double* d = nullptr;
  PPSFM_CUDA(ctx, cudaMalloc(&d, sizeof(double) * blocks * threads));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best[2] = {0, 0};
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(a, ctx->stream);
      if (mode == 0)
        fp64_rate_kernel<true><<<blocks, threads, 0, ctx->stream>>>(d, iters, 1.0000001, 1e-9);
      else
        fp64_rate_kernel<false><<<blocks, threads, 0, ctx->stream>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(b, ctx->stream);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      const double ops = (double)blocks * threads * iters * 8.0;
      const double tips = ops / (ms * 1e-3) / 1e12;
      if (rep > 0 && tips > best[mode]) best[mode] = tips;
    }
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  PPSFM_CUDA(ctx, cudaGetLastError());
  if (dfma_tips) *dfma_tips = best[0];
  if (dmuladd_tips) *dmuladd_tips = best[1];
  return PPSFM_OK;
}

// Returns the packed float FMA throughput in 1e12 FMAs per second (x2 = TFLOP/s FP32).
int ppsfm_bench_fp32_peak(ppsfm_ctx* ctx, double* ffma_tips) {
  if (!ctx) return PPSFM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const int blocks = ctx->num_sms * 8, threads = 256, iters = 8192;
  float* d = nullptr;
  PPSFM_CUDA(ctx, cudaMalloc(&d, sizeof(float) * blocks * threads));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a, ctx->stream);
    fp32_rate_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.9999999f, 1e-6f);
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double fmas = (double)blocks * threads * iters * 16.0;
    const double tips = fmas / (ms * 1e-3) / 1e12;
    if (rep > 0 && tips > best) best = tips;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  PPSFM_CUDA(ctx, cudaGetLastError());
  if (ffma_tips) *ffma_tips = best;
  return PPSFM_OK;
}

// Evicts L2 by writing `bytes` (>= 256 MB recommended) on the context stream; blocking.
int ppsfm_bench_l2_flush(ppsfm_ctx* ctx, size_t bytes) {
  if (!ctx) return PPSFM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  static thread_local ppsfm::DevBuf buf;
  PPSFM_CUDA(ctx, buf.reserve(bytes));
  PPSFM_CUDA(ctx, cudaMemsetAsync(buf.p, 1, bytes, ctx->stream));
  PPSFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PPSFM_OK;
}

}  // extern "C"
