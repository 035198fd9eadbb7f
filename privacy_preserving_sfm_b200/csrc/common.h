// common.h — context, error handling and device-buffer helpers shared by the translation units
// of libppsfm_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "../../include/ppsfm_b200.h"

namespace ppsfm {

// Growable device / pinned-host buffers (never shrink; freed with the context).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace ppsfm

// Correspondence set resident in HBM.
struct ppsfm_corr {
  size_t n = 0;
  double* corr6 = nullptr;    // n x 6 interleaved (l0,l1,l2,X0,X1,X2): 48 B per correspondence
  float* corr6f = nullptr;    // ceil(n/2) x 12: the same rounded to float, two correspondences
                              // a, b per record (l0a,l0b,l1a,l1b,l2a,l2b,X0a,X0b,...) — score filter
  uint8_t* aligned = nullptr; // n bytes
  double* bounds = nullptr;   // 3 doubles: max |X_k|, max(|l_0|, |l_1|), max |l_2| (score filter)
  bool owns = true;           // false: storage belongs to the context's scratch buffers
};

struct ppsfm_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {nullptr};
  std::string last_error;
  std::mt19937 prng{0u};  // util/random.h:46 kDefaultPRNGSeed = 0

  // RANSAC scratch
  ppsfm::DevBuf d_samples, d_models, d_num_models, d_cmodels, d_msrc, d_K, d_part_cnt,
      d_part_sum, d_cnt, d_sum, d_eidx, d_emodels, d_rbuf, d_esum, d_ecnt, d_mask, d_tmp_corr,
      d_tmp_aligned, d_corr6, d_aligned, d_bounds, d_corr6f;
  ppsfm::PinBuf h_samples, h_num_models, h_cnt, h_sum, h_eidx, h_emodels, h_esum, h_ecnt, h_mask,
      h_K, h_stage;
  ppsfm_ransac_timing timing{};
  // RANSAC wave pipeline (ransac_host.cu RansacResident): per-wave buffers and events, a
  // high-priority stream for the sample copies / solve kernels / exact kernels
  static constexpr int kWaveSlots = 4;
  struct WaveSlot {
    ppsfm::DevBuf d_samples, d_models, d_num_models, d_off, d_part_cnt, d_cnt, d_list;
    ppsfm::PinBuf h_samples, h_off, h_cnt;
    cudaEvent_t ev[6] = {nullptr};  // solve begin / end, score begin / end, results on the host,
                                    // counts exchanged (sharded call)
    cudaStream_t solve_stream = nullptr;  // high priority: the waves' solve kernels overlap
  } wave[kWaveSlots];
  ppsfm::DevBuf d_best_lb;  // best inlier count of the waves scored so far (exact pruning)
  cudaStream_t stream_hi = nullptr;
  cudaStream_t stream_copy = nullptr;  // result copies, so that scoring kernels run back to back
  cudaEvent_t ev_sync = nullptr;
  // index-order support + mask of the current best model, computed ahead of the end of the loop
  ppsfm::DevBuf d_fmodel, d_frbuf, d_fmask, d_fcnt, d_fsum;
  ppsfm::PinBuf h_fmask, h_fres;
  cudaEvent_t ev_final[3] = {nullptr};  // kernel begin / end, results on the host

  // multi-GPU (comm.cu): NCCL communicator, one rank per context
  void* comm = nullptr;
  int rank = 0, world = 1;
};

namespace ppsfm {

inline int fail(ppsfm_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->last_error = buf;
  return code;
}

// Development knob read from the environment (kernel variants under measurement); the defaults
// are the shipped configuration.
inline int tune_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

// Per-device facts and one-time per-device setup.  cudaFuncSetAttribute and the SM count belong to
// a DEVICE, and the ABI allows one context per device (and per host thread) in one process, so
// nothing of this kind may live in a process-wide static: each call site owns a PerDevice<> table
// indexed by the current device, initialised under a std::once_flag per entry.
constexpr int kMaxDevices = 64;
struct DeviceFacts {
  int num_sms = 0;
  int max_smem_optin = 0;
};
template <typename Extra = int>
struct PerDevice {
  struct Entry {
    std::once_flag once;
    DeviceFacts facts;
    Extra extra{};
    cudaError_t err = cudaSuccess;
  };
  Entry entries[kMaxDevices];
  // init(facts, extra) runs once per device (with that device current) and returns a cudaError_t
  template <typename Init>
  Entry& get(Init&& init) {
    int dev = 0;
    cudaGetDevice(&dev);
    Entry& e = entries[(dev >= 0 && dev < kMaxDevices) ? dev : 0];
    std::call_once(e.once, [&] {
      cudaDeviceGetAttribute(&e.facts.num_sms, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&e.facts.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      e.err = init(e.facts, e.extra);
    });
    return e;
  }
};

#define PPSFM_CUDA(ctx, expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::ppsfm::fail((ctx), PPSFM_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__,    \
                           #expr, cudaGetErrorString(_e));                                 \
  } while (0)

}  // namespace ppsfm
