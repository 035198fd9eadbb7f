// comm.cu — NCCL plumbing for the one real exchange step of the path: the
// all-reduce of the reduced camera system (and a few scalars) in sharded bundle adjustment.
//
// libnccl is resolved at run time with dlopen so that a host process that already carries NCCL
// (PyTorch bundles its own libnccl.so.2) shares that copy instead of loading a second one; a
// plain C++ host gets the system library.  Only five entry points are used.
#include <dlfcn.h>

#include <cstring>

#include "common.h"
#include "comm.h"

namespace ppsfm {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclUint32 = 3, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) =
      nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi* Api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);  // already in the process (e.g. via torch)?
    if (api.handle) break;
  }
  if (!api.handle)
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
  if (!api.handle) return &api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
  api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
  api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
  return &api;
}

int NcclFail(ppsfm_ctx* ctx, const char* what, ncclResult_t r) {
  NcclApi* a = Api();
  return fail(ctx, PPSFM_ERR_NCCL, "%s failed: %s", what,
              (a->GetErrorString ? a->GetErrorString(r) : "?"));
}

}  // namespace

int CommAllReduce(ppsfm_ctx* ctx, double* dev, size_t count, bool max_op) {
  if (ctx->world <= 1 || count == 0) return PPSFM_OK;
  NcclApi* a = Api();
  if (!a->ok || !ctx->comm) return fail(ctx, PPSFM_ERR_NCCL, "communicator not initialised");
  const ncclResult_t r = a->AllReduce(dev, dev, count, kNcclFloat64, max_op ? kNcclMax : kNcclSum,
                                      (ncclComm_t)ctx->comm, ctx->stream);
  if (r != 0) return NcclFail(ctx, "ncclAllReduce", r);
  return PPSFM_OK;
}

// A sum all-reduce and a max all-reduce issued as ONE NCCL group (one launch): the cost and the
// gradient max-norm of an accepted LM step.
int CommAllReduceSumAndMax(ppsfm_ctx* ctx, double* sum_dev, size_t sum_count, double* max_dev,
                           size_t max_count) {
  if (ctx->world <= 1) return PPSFM_OK;
  NcclApi* a = Api();
  if (!a->ok || !ctx->comm) return fail(ctx, PPSFM_ERR_NCCL, "communicator not initialised");
  const bool group = a->GroupStart && a->GroupEnd;
  if (group) a->GroupStart();
  ncclResult_t r = a->AllReduce(sum_dev, sum_dev, sum_count, kNcclFloat64, kNcclSum,
                                (ncclComm_t)ctx->comm, ctx->stream);
  if (r == 0)
    r = a->AllReduce(max_dev, max_dev, max_count, kNcclFloat64, kNcclMax, (ncclComm_t)ctx->comm,
                     ctx->stream);
  if (group) {
    const ncclResult_t r2 = a->GroupEnd();
    if (r == 0) r = r2;
  }
  if (r != 0) return NcclFail(ctx, "ncclAllReduce (group)", r);
  return PPSFM_OK;
}

// Sum all-reduce of 32-bit counts on `stream` (the per-wave exchange of a sharded RANSAC call).
int CommAllReduceU32(ppsfm_ctx* ctx, unsigned* dev, size_t count, cudaStream_t stream) {
  if (ctx->world <= 1 || count == 0) return PPSFM_OK;
  NcclApi* a = Api();
  if (!a->ok || !ctx->comm) return fail(ctx, PPSFM_ERR_NCCL, "communicator not initialised");
  const ncclResult_t r = a->AllReduce(dev, dev, count, kNcclUint32, kNcclSum,
                                      (ncclComm_t)ctx->comm, stream);
  if (r != 0) return NcclFail(ctx, "ncclAllReduce", r);
  return PPSFM_OK;
}

}  // namespace ppsfm

using namespace ppsfm;

extern "C" {

int ppsfm_comm_get_unique_id(ppsfm_ctx* ctx, char* id128) {
  if (!id128) return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  NcclApi* a = Api();
  if (!a->ok) return fail(ctx, PPSFM_ERR_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  const ncclResult_t r = a->GetUniqueId(&id);
  if (r != 0) return NcclFail(ctx, "ncclGetUniqueId", r);
  std::memcpy(id128, id.internal, 128);
  return PPSFM_OK;
}

int ppsfm_comm_init(ppsfm_ctx* ctx, int world_size, int rank, const char* id128) {
  if (!ctx || !id128 || world_size < 1 || rank < 0 || rank >= world_size)
    return fail(ctx, PPSFM_ERR_INVALID, "bad communicator arguments");
  if (world_size == 1) {
    ctx->world = 1;
    ctx->rank = 0;
    return PPSFM_OK;
  }
  NcclApi* a = Api();
  if (!a->ok) return fail(ctx, PPSFM_ERR_NCCL, "libnccl.so.2 could not be loaded");
  cudaSetDevice(ctx->device);
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  ncclComm_t comm = nullptr;
  const ncclResult_t r = a->CommInitRank(&comm, world_size, id, rank);
  if (r != 0) return NcclFail(ctx, "ncclCommInitRank", r);
  ctx->comm = comm;
  ctx->world = world_size;
  ctx->rank = rank;
  return PPSFM_OK;
}

void ppsfm_comm_destroy(ppsfm_ctx* ctx) {
  if (!ctx || !ctx->comm) return;
  NcclApi* a = Api();
  if (a->ok) a->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->world = 1;
  ctx->rank = 0;
}

int ppsfm_comm_rank(const ppsfm_ctx* ctx) { return ctx ? ctx->rank : 0; }
int ppsfm_comm_world_size(const ppsfm_ctx* ctx) { return ctx ? ctx->world : 1; }

// Test / bench helper: in-place sum all-reduce of a HOST buffer through the device.
int ppsfm_comm_allreduce_sum_host(ppsfm_ctx* ctx, double* host, size_t count) {
  if (!ctx || !host) return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  cudaSetDevice(ctx->device);
  double* d = nullptr;
  PPSFM_CUDA(ctx, cudaMalloc(&d, sizeof(double) * (count ? count : 1)));
  int rc = PPSFM_OK;
  auto body = [&]() -> int {
    PPSFM_CUDA(ctx, cudaMemcpyAsync(d, host, sizeof(double) * count, cudaMemcpyHostToDevice,
                                    ctx->stream));
    const int r = CommAllReduce(ctx, d, count, false);
    if (r != PPSFM_OK) return r;
    PPSFM_CUDA(ctx, cudaMemcpyAsync(host, d, sizeof(double) * count, cudaMemcpyDeviceToHost,
                                    ctx->stream));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PPSFM_OK;
  };
  rc = body();
  cudaFree(d);
  return rc;
}

}  // extern "C"
