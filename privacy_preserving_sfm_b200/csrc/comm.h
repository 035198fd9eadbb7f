// comm.h — NCCL all-reduce on the context's stream (comm.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

struct ppsfm_ctx;

namespace ppsfm {
// In-place all-reduce (sum or max) of `count` doubles in device memory; no-op for world == 1.
int CommAllReduce(ppsfm_ctx* ctx, double* dev, size_t count, bool max_op);
// One NCCL group: sum all-reduce of sum_dev[sum_count] and max all-reduce of max_dev[max_count].
int CommAllReduceSumAndMax(ppsfm_ctx* ctx, double* sum_dev, size_t sum_count, double* max_dev,
                           size_t max_count);
// In-place sum all-reduce of `count` 32-bit counts on `stream`.
int CommAllReduceU32(ppsfm_ctx* ctx, unsigned* dev, size_t count, cudaStream_t stream);
}  // namespace ppsfm
