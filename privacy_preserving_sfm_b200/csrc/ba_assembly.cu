// ba_assembly.cu — problem assembly on the device.
//
// What BundleAdjuster::SetUp / AddImageToProblem / AddPointToProblem do with a ceres::Problem
// (src/optim/bundle_adjustment.cc:326-542), on the SoA boundary: validate every observation
// (index ranges; CHECK_NEAR(line.head<2>().norm(), 1.0, 1e-6), :374), drop residual blocks whose
// parameter blocks are all constant (Ceres removes them from the program), keep this rank's
// points, order the kept observations point-major (stable: input order inside a track) and build
// the camera-major index.  The host only numbers the camera blocks (O(images)); the O(observations)
// work — two stable radix sorts and a few gathers — runs in HBM right after the raw arrays have
// been uploaded, instead of ~45 ms of host loops at 2 M observations.
#include <cub/device/device_radix_sort.cuh>

#include <cstdint>
#include <initializer_list>

#include "ba_kernels.h"
#include "common.h"

namespace ppsfm {

namespace {

// key = LOCAL point index for kept observations, P_local for dropped ones (they sort to the end)
__global__ void asm_classify_kernel(BaRaw raw, int rank, int world, uint32_t* __restrict__ keys,
                                    int* __restrict__ vals, uint8_t* __restrict__ cam_used,
                                    unsigned long long* __restrict__ err /* [2] first bad obs */) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= raw.O) return;
  const int ci = raw.obs_image[o], pi = raw.obs_point[o];
  uint32_t key = (uint32_t)raw.P_local;
  if (ci < 0 || ci >= raw.C || pi < 0 || pi >= raw.P) {
    atomicMin(&err[0], (unsigned long long)o);
  } else {
    const double l0 = raw.obs_line[3 * o], l1 = raw.obs_line[3 * o + 1];
    const double n2 = l0 * l0 + l1 * l1;
    if (!(n2 > 0.999998 && n2 < 1.000002) && fabs(sqrt(n2) - 1.0) > 1e-6)
      atomicMin(&err[1], (unsigned long long)o);
    // (bit 4 of the device copy of pose_flags: the image's camera has variable intrinsics)
    const bool cc = raw.pose_flags[ci] & 1, pc = raw.point_const[pi] != 0;
    const bool ic = raw.pose_flags[ci] & 16;
    if (!(cc && pc && !ic)) {
      cam_used[ci] = 1;  // global property: identical on every rank
      if (world == 1) key = (uint32_t)pi;
      else if ((pi % world) == rank) key = (uint32_t)(pi / world);
    }
  }
  keys[o] = key;
  vals[o] = (int)o;
}

// start[q] = first sorted position whose key is >= q, q in [0, nkeys]
__global__ void asm_lower_bound_kernel(const uint32_t* __restrict__ keys, int64_t n, int nkeys,
                                       int64_t* __restrict__ start) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > nkeys) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < (uint32_t)q) lo = mid + 1; else hi = mid;
  }
  start[q] = lo;
}

__global__ void asm_gather_kernel(BaRaw raw, const uint32_t* __restrict__ keys,
                                  const int* __restrict__ vals, int64_t K, int* __restrict__ obs_cam,
                                  int* __restrict__ obs_pt, double* __restrict__ obs_line) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int64_t o = vals[k];
  obs_cam[k] = raw.obs_image[o];
  obs_pt[k] = (int)keys[k];
  obs_line[k] = raw.obs_line[3 * o];
  obs_line[K + k] = raw.obs_line[3 * o + 1];
  obs_line[2 * K + k] = raw.obs_line[3 * o + 2];
}

// a point takes part on this rank only if it is variable and has kept observations here
__global__ void asm_point_var_kernel(BaRaw raw, const int64_t* __restrict__ pt_start,
                                     uint8_t* __restrict__ pt_var) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= raw.P_local) return;
  const int64_t pg = (int64_t)p * raw.world + raw.rank;  // the caller's index of local point p
  pt_var[p] = (!raw.point_const[pg] && pt_start[p + 1] > pt_start[p]) ? 1 : 0;
}

__global__ void asm_camera_keys_kernel(const int* __restrict__ obs_cam,
                                       const int* __restrict__ cam_block, int64_t K, int NB,
                                       uint32_t* __restrict__ keys, int* __restrict__ vals) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int b = cam_block[obs_cam[k]];
  keys[k] = b >= 0 ? (uint32_t)b : (uint32_t)NB;
  vals[k] = (int)k;
}

int bits_for(uint32_t max_value) {
  int b = 1;
  while (b < 32 && (1ull << b) <= max_value) ++b;
  return b;
}

// stable sort of (keys, vals); returns pointers to the sorted arrays (inside the double buffers)
cudaError_t sort_pairs(uint32_t* k0, uint32_t* k1, int* v0, int* v1, int64_t n, int bits,
                       cudaStream_t s, const uint32_t** ks, const int** vs) {
  cub::DoubleBuffer<uint32_t> kb(k0, k1);
  cub::DoubleBuffer<int> vb(v0, v1);
  size_t bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, kb, vb, n, 0, bits, s);
  if (e != cudaSuccess) return e;
  void* tmp = nullptr;
  e = cudaMallocAsync(&tmp, bytes < 16 ? 16 : bytes, s);
  if (e != cudaSuccess) return e;
  e = cub::DeviceRadixSort::SortPairs(tmp, bytes, kb, vb, n, 0, bits, s);
  cudaFreeAsync(tmp, s);
  *ks = kb.Current();
  *vs = vb.Current();
  return e;
}

}  // namespace

// Point-major assembly.  Fills d.K, d.obs_cam, d.obs_pt, d.obs_line, d.pt_start, d.pt_var;
// cam_used_host[C] (which images have a kept observation, on any rank) and the first observation
// with an index / line-normal violation (-1 = none) come back to the host.
cudaError_t ba_assemble_points(BaDev& d, const BaRaw& raw, int rank, int world,
                               void* (*alloc)(void*, size_t), void* alloc_ctx, cudaStream_t s,
                               uint8_t* cam_used_host, int64_t* first_bad_index,
                               int64_t* first_bad_norm) {
  const int64_t O = raw.O;
  const int P = raw.P_local, C = raw.C;  // per-point arrays are local
  cudaError_t e = cudaSuccess;
  auto tmp_alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMallocAsync(p, bytes < 16 ? 16 : bytes, s);
  };
  uint32_t *k0 = nullptr, *k1 = nullptr;
  int *v0 = nullptr, *v1 = nullptr;
  uint8_t* cam_used = nullptr;
  unsigned long long* err = nullptr;
  tmp_alloc((void**)&k0, sizeof(uint32_t) * (size_t)O);
  tmp_alloc((void**)&k1, sizeof(uint32_t) * (size_t)O);
  tmp_alloc((void**)&v0, sizeof(int) * (size_t)O);
  tmp_alloc((void**)&v1, sizeof(int) * (size_t)O);
  tmp_alloc((void**)&cam_used, (size_t)C);
  tmp_alloc((void**)&err, 2 * sizeof(unsigned long long));
  if (e != cudaSuccess) return e;
  cudaMemsetAsync(cam_used, 0, (size_t)(C > 0 ? C : 1), s);
  cudaMemsetAsync(err, 0xff, 2 * sizeof(unsigned long long), s);
  const uint32_t* ks = k0;
  const int* vs = v0;
  if (O > 0) {
    asm_classify_kernel<<<(unsigned)((O + 255) / 256), 256, 0, s>>>(raw, rank, world, k0, v0,
                                                                    cam_used, err);
    e = sort_pairs(k0, k1, v0, v1, O, bits_for((uint32_t)P), s, &ks, &vs);
    if (e != cudaSuccess) return e;
  }
  d.pt_start = (int64_t*)alloc(alloc_ctx, sizeof(int64_t) * ((size_t)P + 2));
  d.pt_var = (uint8_t*)alloc(alloc_ctx, (size_t)(P > 0 ? P : 1));
  if (!d.pt_start || !d.pt_var) return cudaErrorMemoryAllocation;
  // start[P] = number of kept observations (dropped ones carry key P), start[P+1] = O
  asm_lower_bound_kernel<<<(P + 2 + 255) / 256, 256, 0, s>>>(ks, O, P + 1, d.pt_start);
  int64_t K = 0;
  unsigned long long herr[2];
  e = cudaMemcpyAsync(&K, d.pt_start + P, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(herr, err, sizeof(herr), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && C > 0)
    e = cudaMemcpyAsync(cam_used_host, cam_used, (size_t)C, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
  *first_bad_index = herr[0] == ~0ull ? -1 : (int64_t)herr[0];
  *first_bad_norm = herr[1] == ~0ull ? -1 : (int64_t)herr[1];
  d.K = K;
  d.obs_cam = (int*)alloc(alloc_ctx, sizeof(int) * (size_t)(K > 0 ? K : 1));
  d.obs_pt = (int*)alloc(alloc_ctx, sizeof(int) * (size_t)(K > 0 ? K : 1));
  d.obs_line = (double*)alloc(alloc_ctx, sizeof(double) * 3 * (size_t)(K > 0 ? K : 1));
  if (!d.obs_cam || !d.obs_pt || !d.obs_line) return cudaErrorMemoryAllocation;
  if (K > 0 && *first_bad_index < 0)
    asm_gather_kernel<<<(unsigned)((K + 255) / 256), 256, 0, s>>>(raw, ks, vs, K, d.obs_cam,
                                                                  d.obs_pt, d.obs_line);
  if (P > 0) asm_point_var_kernel<<<(P + 255) / 256, 256, 0, s>>>(raw, d.pt_start, d.pt_var);
  for (void* p : {(void*)k0, (void*)k1, (void*)v0, (void*)v1, (void*)cam_used, (void*)err})
    cudaFreeAsync(p, s);
  return cudaGetLastError();
}

// Camera-major index (d.cam_block resident): d.cam_obs = kept observations grouped by camera
// block (stable: ascending observation index inside a block), d.cam_start[NB + 1].
cudaError_t ba_assemble_cameras(BaDev& d, void* (*alloc)(void*, size_t), void* alloc_ctx,
                                cudaStream_t s) {
  const int64_t K = d.K;
  const int NB = d.NB;
  d.cam_start = (int64_t*)alloc(alloc_ctx, sizeof(int64_t) * ((size_t)NB + 2));
  d.cam_obs = (int*)alloc(alloc_ctx, sizeof(int) * (size_t)(K > 0 ? K : 1));
  if (!d.cam_start || !d.cam_obs) return cudaErrorMemoryAllocation;
  cudaError_t e = cudaSuccess;
  uint32_t *k0 = nullptr, *k1 = nullptr;
  int* v1 = nullptr;
  auto tmp_alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMallocAsync(p, bytes < 16 ? 16 : bytes, s);
  };
  tmp_alloc((void**)&k0, sizeof(uint32_t) * (size_t)K);
  tmp_alloc((void**)&k1, sizeof(uint32_t) * (size_t)K);
  tmp_alloc((void**)&v1, sizeof(int) * (size_t)K);
  if (e != cudaSuccess) return e;
  const uint32_t* ks = k0;
  const int* vs = d.cam_obs;
  if (K > 0) {
    asm_camera_keys_kernel<<<(unsigned)((K + 255) / 256), 256, 0, s>>>(d.obs_cam, d.cam_block, K,
                                                                       NB, k0, d.cam_obs);
    e = sort_pairs(k0, k1, d.cam_obs, v1, K, bits_for((uint32_t)NB), s, &ks, &vs);
    if (e != cudaSuccess) return e;
    if (vs != d.cam_obs)
      e = cudaMemcpyAsync(d.cam_obs, vs, sizeof(int) * (size_t)K, cudaMemcpyDeviceToDevice, s);
  }
  asm_lower_bound_kernel<<<(NB + 2 + 255) / 256, 256, 0, s>>>(ks, K, NB + 1, d.cam_start);
  for (void* p : {(void*)k0, (void*)k1, (void*)v1}) cudaFreeAsync(p, s);
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace ppsfm
