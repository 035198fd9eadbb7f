// ba_host.cu — host side of the line-reprojection bundle adjustment: problem assembly on the
// SoA boundary, the trust-region Levenberg-Marquardt loop and the C-ABI entry points.
//
// Replaces BundleAdjuster::Solve (src/optim/bundle_adjustment.cc:260-320) incl. the part the
// reference delegates to ceres::Solve (:306), and RefineAbsolutePoseFromLines
// (src/estimators/pose.cc:96-213).  The minimiser follows Ceres' documented trust-region LM:
// Jacobi column scaling from the initial Jacobian, LM diagonal clamp(diag(J^T J)) / radius,
// Schur elimination of the points, step-quality ratio, radius update
// (SURVEY.md Appendix A); the control flow runs on the host, every array stays in HBM.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "ba_kernels.h"
#include "comm.h"
#include "common.h"
#include "dense_chol.h"

namespace ppsfm {

struct BaState {
  ppsfm_ctx* ctx = nullptr;
  BaDev d;
  ppsfm_ba_options opt{};
  std::vector<void*> allocs;
  // host copies needed to write results back in the caller's order
  int C = 0, P = 0;       // P: points resident on this rank (all of them on a single GPU)
  int P_global = 0;       // points of the caller's problem
  int64_t num_obs_in = 0;
  double *q0 = nullptr, *t0 = nullptr, *X0 = nullptr;  // initial state (for reset)
  double* Xg = nullptr;              // sharded solve: whole-problem displacement buffer (download)
  // intrinsics refinement: candidate parameters, initial parameters, overflow flag of the
  // intrinsics Schur kernel, number of variable intrinsics
  double *cam_params_n = nullptr, *img_params_n = nullptr;
  double *cam_params0 = nullptr, *img_params0 = nullptr;
  int* intr_overflow = nullptr;
  int intr_eff = 0;
  std::vector<double> points_in;     // sharded solve: the caller's points as handed in
  PinBuf h_scalars;
  int64_t launches = 0;
  double lin_ms = 0, lin_launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // Jacobian-build timing without a host sync per launch: a ring of event pairs, read back when
  // the solve is over
  static constexpr int kLinEvents = 64;
  cudaEvent_t lin_ev[2 * kLinEvents] = {nullptr};
  int lin_pending = 0;
  cudaEvent_t evp[4] = {nullptr, nullptr, nullptr, nullptr};  // phase marks of one LM iteration
  // multi-GPU
  int rank = 0, world = 1;
  void* comm = nullptr;  // ncclComm_t
};

namespace {

template <typename T>
cudaError_t DevAlloc(BaState* st, T** p, size_t count) {
  void* v = nullptr;
  // stream-ordered pool allocation: after the first solve the context's pool serves these from
  // cached HBM (no cudaMalloc / cudaFree on the per-call path)
  cudaError_t e = cudaMallocAsync(&v, std::max<size_t>(1, count) * sizeof(T), st->ctx->stream);
  if (e == cudaSuccess) {
    st->allocs.push_back(v);
    *p = static_cast<T*>(v);
  }
  return e;
}

template <typename T>
cudaError_t Upload(BaState* st, T** p, const std::vector<T>& h) {
  cudaError_t e = DevAlloc(st, p, h.size());
  if (e != cudaSuccess) return e;
  if (!h.empty())
    e = cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice,
                        st->ctx->stream);
  return e;
}

// allocation callback for build_schur_lists
// Parameter groups of the camera models (FocalLengthIdxs / PrincipalPointIdxs / ExtraParamsIdxs,
// src/base/camera_models.h:597-846) as bit masks over Camera::Params().
const unsigned kFocalMask[11] = {0x1, 0x3, 0x1, 0x1, 0x3, 0x3, 0x3, 0x3, 0x1, 0x1, 0x3};
const unsigned kPrincipalMask[11] = {0x6, 0xc, 0x6, 0x6, 0xc, 0xc, 0xc, 0xc, 0x6, 0x6, 0xc};
const unsigned kExtraMask[11] = {0x0, 0x0, 0x8, 0x18, 0xf0, 0xf0, 0xff0, 0x10, 0x8, 0x18, 0xff0};
const int kNumParams[11] = {3, 4, 4, 5, 8, 8, 12, 5, 4, 5, 12};

// BundleAdjuster::ParameterizeCameras (bundle_adjustment.cc:490-528): the variable parameters of
// camera `cam` (0 = constant camera).
unsigned VariableIntrinsics(const ppsfm_ba_problem* pb, const ppsfm_ba_options* opt, int cam) {
  const int m = pb->camera_model[cam];
  if (m < 0 || m > 10) return 0;
  unsigned mask = 0;
  if (opt->refine_focal_length) mask |= kFocalMask[m];
  if (opt->refine_principal_point) mask |= kPrincipalMask[m];
  if (opt->refine_extra_params) mask |= kExtraMask[m];
  if (pb->camera_const && pb->camera_const[cam]) mask = 0;
  return mask;
}

void* StateAlloc(void* state, size_t bytes) {
  char* p = nullptr;
  return DevAlloc(static_cast<BaState*>(state), &p, bytes) == cudaSuccess ? p : nullptr;
}

double Secs(std::chrono::steady_clock::time_point a) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
}

}  // namespace

void BaFree(BaState* st) {
  if (!st) return;
  for (void* p : st->allocs) cudaFreeAsync(p, st->ctx->stream);
  cudaStreamSynchronize(st->ctx->stream);
  st->h_scalars.release();
  for (auto& e : st->lin_ev)
    if (e) cudaEventDestroy(e);
  if (st->ev0) cudaEventDestroy(st->ev0);
  if (st->ev1) cudaEventDestroy(st->ev1);
  for (auto& e : st->evp)
    if (e) cudaEventDestroy(e);
  delete st;
}

// Assembly: what BundleAdjuster::SetUp / AddImageToProblem / AddPointToProblem /
// ParameterizeCameras / ParameterizePoints (bundle_adjustment.cc:326-542) do with a ceres::Problem,
// on the SoA boundary: drop residual blocks whose parameter blocks are all constant, number the
// variable camera blocks, record the per-dimension tangent masks (constant pose, constant
// tvec components = SubsetParameterization), order observations point-major and build the
// camera-major index.
int BaCreate(ppsfm_ctx* ctx, const ppsfm_ba_problem* pb, const ppsfm_ba_options* opt,
             int rank, int world, BaState** out) {
  if (!ctx || !pb || !opt || !out) return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  const int C = pb->num_images, P_global = pb->num_points;
  const int64_t O = pb->num_obs;
  if (C < 0 || P_global < 0 || O < 0) return fail(ctx, PPSFM_ERR_INVALID, "negative size");
  // sharded solve: the points p with p % world == rank live here, numbered p / world; every
  // per-point array (and every per-point kernel) is sized by the local count
  const int P = world <= 1 ? P_global
                           : (P_global > rank ? (P_global - rank + world - 1) / world : 0);
  for (int i = 0; i < C; ++i) {
    const int cam = pb->image_camera[i];
    if (cam < 0 || cam >= pb->num_cameras)
      return fail(ctx, PPSFM_ERR_INVALID, "image %d references a missing camera", i);
    const int m = pb->camera_model[cam];
    if (m < 0 || m > 10)
      return fail(ctx, PPSFM_ERR_INVALID,
                  "camera model id %d unknown (COLMAP ids 0..10)", m);
  }
  static const bool timing = tune_int("PPSFM_BA_TIMING", 0) != 0;
  const auto t_create = std::chrono::steady_clock::now();
  BaState* st = new BaState();
  st->ctx = ctx;
  st->opt = *opt;
  st->C = C;
  st->P = P;
  st->P_global = P_global;
  st->num_obs_in = O;
  st->rank = rank;
  st->world = world;
  cudaStream_t s = ctx->stream;

  BaDev& d = st->d;
  d.C = C; d.P = P;
  cudaError_t e = cudaSuccess;
#define BA_TRY(x) do { if (e == cudaSuccess) e = (x); } while (0)
  // ---- raw arrays to HBM, assembly on the device (ba_assembly.cu)
  BaRaw raw;
  raw.C = C; raw.P = P_global; raw.O = O;
  raw.P_local = P; raw.rank = world <= 1 ? 0 : rank; raw.world = world <= 1 ? 1 : world;
  std::vector<uint8_t> flags_h(std::max(C, 1), 0), pconst_h(std::max(P_global, 1), 0);
  if (pb->pose_flags) std::copy(pb->pose_flags, pb->pose_flags + C, flags_h.begin());
  if (pb->point_const)
    for (int i = 0; i < P_global; ++i) pconst_h[i] = pb->point_const[i] ? 1 : 0;
  // images whose camera has variable intrinsics keep their residuals even with a constant pose
  // and a constant point (bit 4 of the device copy of the flags)
  std::vector<unsigned> intr_candidate(std::max(pb->num_cameras, 1), 0u);
  for (int c = 0; c < pb->num_cameras; ++c) intr_candidate[c] = VariableIntrinsics(pb, opt, c);
  for (int i = 0; i < C; ++i) {
    flags_h[i] &= 0x0f;
    if (intr_candidate[pb->image_camera[i]]) flags_h[i] |= 16;
  }
  std::vector<void*> raw_tmp;
  auto raw_upload = [&](const void* src, size_t bytes) -> void* {
    void* p = nullptr;
    BA_TRY(cudaMallocAsync(&p, std::max<size_t>(bytes, 16), s));
    if (e == cudaSuccess) raw_tmp.push_back(p);
    if (e == cudaSuccess && bytes > 0)
      e = cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, s);
    return p;
  };
  raw.obs_image = (const int*)raw_upload(pb->obs_image, sizeof(int32_t) * (size_t)O);
  raw.obs_point = (const int*)raw_upload(pb->obs_point, sizeof(int32_t) * (size_t)O);
  raw.obs_line = (const double*)raw_upload(pb->obs_line, sizeof(double) * 3 * (size_t)O);
  raw.pose_flags = (const uint8_t*)raw_upload(flags_h.data(), (size_t)C);
  raw.point_const = (const uint8_t*)raw_upload(pconst_h.data(), (size_t)P_global);
  std::vector<uint8_t> cam_used(std::max(C, 1), 0);
  int64_t bad_index = -1, bad_norm = -1;
  BA_TRY(ba_assemble_points(d, raw, rank, world, &StateAlloc, st, s, cam_used.data(), &bad_index,
                            &bad_norm));
  if (e == cudaSuccess && (bad_index >= 0 || bad_norm >= 0)) {
    for (void* p : raw_tmp) cudaFreeAsync(p, s);
    BaFree(st);
    // the reference walks the observations in order: report the first violation it would meet
    if (bad_index >= 0 && (bad_norm < 0 || bad_index <= bad_norm))
      return fail(ctx, PPSFM_ERR_INVALID, "observation %lld references a missing image/point",
                  (long long)bad_index);
    return fail(ctx, PPSFM_ERR_INVALID, "observation %lld: line normal is not unit length",
                (long long)bad_norm);
  }
  for (void* p : raw_tmp) cudaFreeAsync(p, s);
  const int64_t K = d.K;
  std::vector<int> cam_block(C, -1), block_img;
  std::vector<uint8_t> cam_mask(C, 0);
  std::vector<double> q(pb->qvecs, pb->qvecs + 4 * (size_t)C), t(pb->tvecs, pb->tvecs + 3 * (size_t)C);
  for (int i = 0; i < C; ++i) {
    if ((flags_h[i] & 1) || !cam_used[i]) continue;
    cam_block[i] = (int)block_img.size();
    block_img.push_back(i);
    const uint8_t f = flags_h[i];
    uint8_t m = 0x07;
    for (int k = 0; k < 3; ++k)
      if (!(f & (2 << k))) m |= (uint8_t)(8 << k);
    cam_mask[i] = m;
    // image.NormalizeQvec()  (bundle_adjustment.cc:355)
    double nrm = 0;
    for (int k = 0; k < 4; ++k) nrm += q[4 * i + k] * q[4 * i + k];
    nrm = std::sqrt(nrm);
    if (nrm > 0) for (int k = 0; k < 4; ++k) q[4 * i + k] /= nrm;
  }
  const int NB = (int)block_img.size();
  std::vector<int> img_model(C);
  std::vector<double> img_params(12 * (size_t)C, 0.0);
  for (int i = 0; i < C; ++i) {
    const int cam = pb->image_camera[i];
    img_model[i] = pb->camera_model[cam];
    if (img_model[i] >= 5) d.has_ext_models = 1;
    std::memcpy(&img_params[12 * (size_t)i], pb->camera_params + 12 * (size_t)cam, 12 * sizeof(double));
  }
  // intrinsics blocks: cameras of the images that are part of the problem, in image order
  std::vector<int> cam_intr_block(std::max(pb->num_cameras, 1), -1), cam_nparams(std::max(pb->num_cameras, 1), 0);
  std::vector<unsigned> intr_mask;
  for (int c = 0; c < pb->num_cameras; ++c) {
    const int m = pb->camera_model[c];
    cam_nparams[c] = (m >= 0 && m <= 10) ? kNumParams[m] : 0;
  }
  for (int i = 0; i < C; ++i) {
    const int cam = pb->image_camera[i];
    if (cam_used[i] && intr_candidate[cam] && cam_intr_block[cam] < 0) {
      cam_intr_block[cam] = (int)intr_mask.size();
      intr_mask.push_back(intr_candidate[cam]);
      st->intr_eff += __builtin_popcount(intr_candidate[cam]);
    }
  }
  const int NCv = (int)intr_mask.size();
  if (NCv > kMaxVarCams) {
    BaFree(st);
    return fail(ctx, PPSFM_ERR_INVALID, "%d cameras with variable intrinsics (limit %d)", NCv,
                kMaxVarCams);
  }
  d.NB = NB; d.NCv = NCv; d.n = 6 * NB + kIntrW * NCv; d.ld = chol_ld(d.n);
  BA_TRY(Upload(st, &d.cam_block, cam_block));
  BA_TRY(ba_assemble_cameras(d, &StateAlloc, st, s));
  if (timing) {
    cudaStreamSynchronize(s);
    std::fprintf(stderr, "[ba] upload + device assembly %.2f ms\n", 1e3 * Secs(t_create));
  }
  std::vector<double> X(3 * (size_t)P);
  for (int p = 0; p < P; ++p) {
    const size_t pg = (size_t)p * raw.world + raw.rank;
    for (int k = 0; k < 3; ++k) X[3 * (size_t)p + k] = pb->points[3 * pg + k];
  }
  if (world > 1) {
    BA_TRY(DevAlloc(st, &st->Xg, 3 * (size_t)std::max(P_global, 1)));
    st->points_in.assign(pb->points, pb->points + 3 * (size_t)P_global);
  }
  BA_TRY(Upload(st, &d.block_img, block_img));
  BA_TRY(Upload(st, &d.cam_mask, cam_mask));
  BA_TRY(Upload(st, &d.img_model, img_model));
  BA_TRY(Upload(st, &d.img_params, img_params));
  {
    d.num_cameras = pb->num_cameras;
    std::vector<int> img_cam(pb->image_camera, pb->image_camera + C);
    std::vector<int> cam_model(pb->camera_model, pb->camera_model + pb->num_cameras);
    std::vector<double> cam_params(pb->camera_params,
                                   pb->camera_params + 12 * (size_t)pb->num_cameras);
    BA_TRY(Upload(st, &d.img_cam, img_cam));
    BA_TRY(Upload(st, &d.cam_model, cam_model));
    BA_TRY(Upload(st, &d.cam_params, cam_params));
  }
  BA_TRY(Upload(st, &d.q, q));
  BA_TRY(Upload(st, &d.t, t));
  BA_TRY(Upload(st, &d.X, X));
  BA_TRY(Upload(st, &st->q0, q));
  BA_TRY(Upload(st, &st->t0, t));
  BA_TRY(Upload(st, &st->X0, X));
  BA_TRY(DevAlloc(st, &d.qn, 4 * (size_t)C));
  BA_TRY(DevAlloc(st, &d.tn, 3 * (size_t)C));
  BA_TRY(DevAlloc(st, &d.Xn, 3 * (size_t)P));
  BA_TRY(DevAlloc(st, &d.J, ba_j_doubles(K)));
  BA_TRY(DevAlloc(st, &d.cam_scale, 6 * (size_t)NB));
  BA_TRY(DevAlloc(st, &d.pt_scale, 3 * (size_t)P));
  // U and g_c (and the intrinsics blocks U_ii, U_ic, g_i behind them) in ONE allocation: the
  // sharded solve sums them over the ranks with one all-reduce
  BA_TRY(DevAlloc(st, &d.U, 42 * (size_t)NB + intr_normal_doubles(NB, NCv)));
  d.gc = d.U + 36 * (size_t)NB;
  if (NCv > 0) {
    d.Uii = d.U + 42 * (size_t)NB;
    d.Uic = d.Uii + (size_t)NCv * kIntrW * kIntrW;
    d.gi = d.Uic + (size_t)NB * kIntrW * 6;
    BA_TRY(Upload(st, &d.cam_intr_block, cam_intr_block));
    BA_TRY(Upload(st, &d.intr_mask, intr_mask));
    BA_TRY(Upload(st, &d.cam_nparams, cam_nparams));
    BA_TRY(DevAlloc(st, &d.intr_scale, (size_t)NCv * kIntrW));
    BA_TRY(DevAlloc(st, &d.Ji, 2 * (size_t)kIntrW * (size_t)std::max<int64_t>(K, 1)));
    BA_TRY(DevAlloc(st, &st->cam_params_n, 12 * (size_t)pb->num_cameras));
    BA_TRY(DevAlloc(st, &st->img_params_n, 12 * (size_t)std::max(C, 1)));
    std::vector<double> cp0(pb->camera_params, pb->camera_params + 12 * (size_t)pb->num_cameras);
    BA_TRY(Upload(st, &st->cam_params0, cp0));
    BA_TRY(Upload(st, &st->img_params0, img_params));
    BA_TRY(DevAlloc(st, &st->intr_overflow, 1));
    BA_TRY(cudaMemsetAsync(st->intr_overflow, 0, sizeof(int), s));
  }
  BA_TRY(DevAlloc(st, &d.Upart, 4 * 27 * (size_t)NB));
  BA_TRY(DevAlloc(st, &d.V, 6 * (size_t)P));
  BA_TRY(DevAlloc(st, &d.gp, 3 * (size_t)P));
  BA_TRY(DevAlloc(st, &d.Vinv, 6 * (size_t)P));
  if (timing) {
    cudaStreamSynchronize(s);
    std::fprintf(stderr, "[ba] + uploads %.2f ms\n", 1e3 * Secs(t_create));
  }
  BA_TRY(DevAlloc(st, &d.S, (size_t)d.ld * d.ld));
  if (world > 1) BA_TRY(DevAlloc(st, &d.Spacked, packed_lower_doubles(d.n) + 1));
  // zeroed once: the assembly overwrites every lower block and the rhs row each iteration, the
  // factorisation keeps the padding rows zero
  BA_TRY(cudaMemsetAsync(d.S, 0, sizeof(double) * (size_t)d.ld * d.ld, s));
  BA_TRY(build_schur_lists(d, &StateAlloc, st, s));
  BA_TRY(DevAlloc(st, &d.dc, (size_t)std::max(1, d.n)));
  BA_TRY(DevAlloc(st, &d.dp, 3 * (size_t)P));
  BA_TRY(DevAlloc(st, &d.u, 2 * (size_t)K));
  BA_TRY(DevAlloc(st, &d.chol_status, 1));
  BA_TRY(DevAlloc(st, &d.chol_work, chol_work_doubles(d.n)));
  d.num_partials = (int)std::max<int64_t>((K + 255) / 256, (P + 255) / 256) + 1;
  BA_TRY(DevAlloc(st, &d.partials, 3 * (size_t)d.num_partials));
  BA_TRY(DevAlloc(st, &d.scalars, kNumScalars));
  BA_TRY(cudaMemsetAsync(d.scalars, 0, sizeof(double) * kNumScalars, s));
  // (+ two ints behind the scalars: the Cholesky status and the intrinsics overflow flag, so that
  // their device -> host copies are truly asynchronous — a copy to pageable memory blocks the host
  // until the stream has drained, i.e. until the factorisation is over)
  BA_TRY(st->h_scalars.reserve(sizeof(double) * kNumScalars + 2 * sizeof(int)));
  BA_TRY(cudaEventCreate(&st->ev0));
  BA_TRY(cudaEventCreate(&st->ev1));
  for (auto& ev : st->evp) BA_TRY(cudaEventCreate(&ev));
  BA_TRY(cudaStreamSynchronize(s));
#undef BA_TRY
  if (e != cudaSuccess) {
    BaFree(st);
    return fail(ctx, PPSFM_ERR_CUDA, "BA setup: %s", cudaGetErrorString(e));
  }
  if (timing) std::fprintf(stderr, "[ba] + structure build, create total %.2f ms\n", 1e3 * Secs(t_create));
  *out = st;
  return PPSFM_OK;
}

namespace {

// Sum-reduction across ranks (NCCL all-reduce over NVLink); identity for a single GPU.
int AllReduceSum(BaState* st, double* dev, size_t count);

struct Scalars {
  double v[kNumScalars];
};

int FetchScalars(BaState* st, Scalars* out) {
  ppsfm_ctx* ctx = st->ctx;
  PPSFM_CUDA(ctx, cudaMemcpyAsync(st->h_scalars.p, st->d.scalars, sizeof(double) * kNumScalars,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  PPSFM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::memcpy(out->v, st->h_scalars.p, sizeof(out->v));
  return PPSFM_OK;
}

}  // namespace

// The LM loop (mirrors ceres::internal::TrustRegionMinimizer with LevenbergMarquardtStrategy).
int BaRun(BaState* st, ppsfm_ba_summary* sum) {
  ppsfm_ctx* ctx = st->ctx;
  cudaStream_t s = ctx->stream;
  const ppsfm_ba_options& opt = st->opt;
  BaDev& d = st->d;
  const auto t_start = std::chrono::steady_clock::now();
  std::memset(sum, 0, sizeof(*sum));
  sum->num_residuals = 2 * st->num_obs_in;
  sum->num_residuals_reduced = 2 * d.K;  // this rank's share when sharded
  st->launches = 0;
  st->lin_ms = 0;
  st->lin_launches = 0;
  if (st->num_obs_in == 0) return PPSFM_NO_SOLUTION;  // bundle_adjustment.cc:269-271
  const BaLoss loss{opt.loss_type, opt.loss_scale};
  Scalars sc;

  auto drain_lin_events = [&]() {  // (the stream must have passed the recorded events)
    for (int i = 0; i < st->lin_pending; ++i) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, st->lin_ev[2 * i], st->lin_ev[2 * i + 1]) == cudaSuccess) {
        st->lin_ms += ms;
        st->lin_launches += 1;
      }
    }
    st->lin_pending = 0;
  };
  auto linearize = [&](const double* q, const double* t, const double* X, bool jac) -> int {
    int slot = -1;
    if (jac) {
      if (st->lin_pending == BaState::kLinEvents) {
        PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
        drain_lin_events();
      }
      slot = st->lin_pending++;
      for (int k = 0; k < 2; ++k)
        if (!st->lin_ev[2 * slot + k]) PPSFM_CUDA(ctx, cudaEventCreate(&st->lin_ev[2 * slot + k]));
      PPSFM_CUDA(ctx, cudaEventRecord(st->lin_ev[2 * slot], s));
    }
    st->launches += launch_linearize(d, q, t, X, jac, loss, s);
    if (jac) PPSFM_CUDA(ctx, cudaEventRecord(st->lin_ev[2 * slot + 1], s));
    if (jac) st->launches += launch_intr_jacobian(d, q, t, X, loss, s);
    return PPSFM_OK;
  };
  auto normal_equations = [&]() -> int {
    st->launches += launch_normal_equations(d, s);
    st->launches += launch_intr_normal(d, s);
    if (st->world > 1) {  // cameras are replicated: U, g_c are sums over all ranks' observations
      // (U, g_c and the intrinsics blocks are contiguous)
      const int rc = AllReduceSum(st, d.U, 42 * (size_t)d.NB + intr_normal_doubles(d.NB, d.NCv));
      if (rc != PPSFM_OK) return rc;
    }
    return PPSFM_OK;
  };
  auto cost_and_gradient = [&](double* cost, double* gmax) -> int {
    st->launches += launch_gradient_max_norm(d, s);
    st->launches += launch_intr_gradient(d, s);
    if (st->world > 1) {
      const int rc = CommAllReduceSumAndMax(st->ctx, d.scalars + kCost, 1, d.scalars + kGradMax, 1);
      if (rc != PPSFM_OK) return rc;
    }
    int rc = FetchScalars(st, &sc);
    if (rc != PPSFM_OK) return rc;
    *cost = sc.v[kCost];
    *gmax = sc.v[kGradMax];
    return PPSFM_OK;
  };

  int eff = 0;
  {
    std::vector<uint8_t> mask(d.C), pv(d.P);
    PPSFM_CUDA(ctx, cudaMemcpy(mask.data(), d.cam_mask, d.C, cudaMemcpyDeviceToHost));
    PPSFM_CUDA(ctx, cudaMemcpy(pv.data(), d.pt_var, d.P, cudaMemcpyDeviceToHost));
    for (uint8_t m : mask) eff += __builtin_popcount(m);
    for (uint8_t v : pv) eff += v ? 3 : 0;
  }
  eff += st->intr_eff;
  sum->num_effective_parameters_reduced = eff;

  // unit scales, first linearisation
  {
    std::vector<double> ones(std::max<size_t>(6 * (size_t)d.NB, 3 * (size_t)d.P), 1.0);
    PPSFM_CUDA(ctx, cudaMemcpyAsync(d.cam_scale, ones.data(), sizeof(double) * 6 * d.NB,
                                    cudaMemcpyHostToDevice, s));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(d.pt_scale, ones.data(), sizeof(double) * 3 * d.P,
                                    cudaMemcpyHostToDevice, s));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
  }
  st->launches += launch_intr_scales(d, false, s);
  int rc = linearize(d.q, d.t, d.X, true);
  if (rc != PPSFM_OK) return rc;
  rc = normal_equations();
  if (rc != PPSFM_OK) return rc;
  if (opt.jacobi_scaling) {
    st->launches += launch_jacobi_scales(d, s);
    st->launches += launch_intr_scales(d, true, s);
    rc = linearize(d.q, d.t, d.X, true);
    if (rc != PPSFM_OK) return rc;
    rc = normal_equations();
    if (rc != PPSFM_OK) return rc;
  }
  double cost = 0, gmax = 0;
  rc = cost_and_gradient(&cost, &gmax);
  if (rc != PPSFM_OK) return rc;
  sum->initial_cost = cost;

  int tl = 0;
  auto trace = [&](double c, double radius, int acc) {
    if (tl < PPSFM_BA_MAX_TRACE) {
      sum->trace_cost[tl] = c;
      sum->trace_radius[tl] = radius;
      sum->trace_accepted[tl] = acc;
      ++tl;
    }
  };
  double radius = opt.initial_trust_region_radius;
  double decrease_factor = 2.0;
  trace(cost, radius, 1);
  sum->termination_type = 1;
  int invalid = 0;
  bool done = eff == 0 || gmax <= opt.gradient_tolerance;
  if (done) sum->termination_type = 0;
  double solver_s = 0;

  static const bool gap_trace = tune_int("PPSFM_BA_GAPS", 0) != 0;
  double g_enq1 = 0, g_wait1 = 0, g_enq2 = 0, g_wait2 = 0;
  int g_n = 0;
  for (int iter = 0; !done && iter < opt.max_num_iterations; ++iter) {
    const auto t_lin = std::chrono::steady_clock::now();
    // --- reduced camera system + dense Cholesky
    PPSFM_CUDA(ctx, cudaEventRecord(st->evp[0], s));
    st->launches += launch_build_reduced_system(d, radius, opt.min_lm_diagonal,
                                                opt.max_lm_diagonal, st->rank == 0, s);
    st->launches += launch_intr_reduced_rows(d, radius, opt.min_lm_diagonal, opt.max_lm_diagonal,
                                             st->rank == 0, st->intr_overflow, s);
    if (st->world > 1) {
      // rank 0 contributed blockdiag(U + D) and -g_c; every rank its points' Schur products
      // (only the lower triangle and the rhs row are referenced: pack, reduce half the bytes)
      launch_pack_lower(d.S, d.n, d.ld, d.Spacked, s);
      rc = AllReduceSum(st, d.Spacked, packed_lower_doubles(d.n));
      if (rc != PPSFM_OK) return rc;
      launch_unpack_lower(d.S, d.n, d.ld, d.Spacked, s);
      st->launches += 2;
    }
    PPSFM_CUDA(ctx, cudaEventRecord(st->evp[1], s));
    int* h_flags = reinterpret_cast<int*>(st->h_scalars.as<double>() + kNumScalars);
    h_flags[0] = h_flags[1] = 0;  // [0] Cholesky failed, [1] intrinsics overflow (pinned memory)
    if (d.NCv > 0)
      PPSFM_CUDA(ctx, cudaMemcpyAsync(h_flags + 1, st->intr_overflow, sizeof(int),
                                      cudaMemcpyDeviceToHost, s));
    if (d.n > 0) {
      st->launches += chol_solve_bordered(d.S, d.n, d.ld, d.dc, d.chol_work, d.chol_status, s);
      PPSFM_CUDA(ctx, cudaGetLastError());  // a refused launch must not read as "factorised"
      PPSFM_CUDA(ctx, cudaMemcpyAsync(h_flags, d.chol_status, sizeof(int),
                                      cudaMemcpyDeviceToHost, s));
    }
    PPSFM_CUDA(ctx, cudaEventRecord(st->evp[2], s));
    st->launches += launch_backsubstitute_and_update(d, st->rank == 0, s);
    st->launches += launch_intr_update(d, st->cam_params_n, st->img_params_n, st->rank == 0, s);
    // candidate cost
    if (d.NCv > 0) {
      BaDev dn = d;  // the candidate intrinsics
      dn.cam_params = st->cam_params_n;
      dn.img_params = st->img_params_n;
      st->launches += launch_linearize(dn, d.qn, d.tn, d.Xn, false, loss, s);
    } else {
      st->launches += launch_linearize(d, d.qn, d.tn, d.Xn, false, loss, s);
    }
    if (st->world > 1) {
      // candidate cost, model cost change, step^2, x^2: four adjacent scalars, one all-reduce
      static_assert(kCost == 0 && kModelChange == 1 && kStepSq == 2 && kXSq == 3, "adjacent");
      rc = AllReduceSum(st, d.scalars + kCost, 4);
      if (rc != PPSFM_OK) return rc;
    }
    PPSFM_CUDA(ctx, cudaEventRecord(st->evp[3], s));
    const double t_enq1 = Secs(t_lin);
    rc = FetchScalars(st, &sc);
    if (rc != PPSFM_OK) return rc;
    solver_s += Secs(t_lin);
    g_enq1 += t_enq1;
    g_wait1 += Secs(t_lin) - t_enq1;
    const int chol_failed = h_flags[0], intr_overflow = h_flags[1];  // (the stream is drained)
    if (intr_overflow)
      return fail(ctx, PPSFM_ERR_INVALID,
                  "a point is seen by more distinct cameras with variable intrinsics than supported");
    {
      float m01 = 0, m12 = 0, m23 = 0;
      cudaEventElapsedTime(&m01, st->evp[0], st->evp[1]);
      cudaEventElapsedTime(&m12, st->evp[1], st->evp[2]);
      cudaEventElapsedTime(&m23, st->evp[2], st->evp[3]);
      sum->schur_time_s += 1e-3 * m01;
      sum->cholesky_time_s += 1e-3 * m12;
      sum->backsub_time_s += 1e-3 * m23;
    }
    const double model_cost_change = sc.v[kModelChange];
    const double cost_new = sc.v[kCost];
    const double step_norm = std::sqrt(sc.v[kStepSq]), x_norm = std::sqrt(sc.v[kXSq]);

    if (chol_failed || !(model_cost_change > 0.0)) {
      ++sum->num_unsuccessful_steps;
      if (++invalid >= opt.max_num_consecutive_invalid_steps) {
        sum->termination_type = 2;
        trace(cost, radius, 0);
        break;
      }
      radius = radius / decrease_factor;  // LevenbergMarquardtStrategy::StepIsInvalid
      decrease_factor *= 2.0;
      trace(cost, radius, 0);
      continue;
    }
    invalid = 0;
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      sum->termination_type = 0;
      break;
    }
    const double cost_change = cost - cost_new;
    if (std::fabs(cost_change) <= opt.function_tolerance * cost) {
      sum->termination_type = 0;
      break;
    }
    const double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > opt.min_relative_decrease) {
      const double tmp = 2.0 * relative_decrease - 1.0;
      radius = radius / std::max(1.0 / 3.0, 1.0 - tmp * tmp * tmp);
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      std::swap(d.q, d.qn);
      std::swap(d.t, d.tn);
      std::swap(d.X, d.Xn);
      if (d.NCv > 0) {
        std::swap(d.cam_params, st->cam_params_n);
        std::swap(d.img_params, st->img_params_n);
      }
      const auto t_acc = std::chrono::steady_clock::now();
      rc = linearize(d.q, d.t, d.X, true);
      if (rc != PPSFM_OK) return rc;
      rc = normal_equations();
      if (rc != PPSFM_OK) return rc;
      const double t_enq2 = Secs(t_acc);
      rc = cost_and_gradient(&cost, &gmax);
      if (rc != PPSFM_OK) return rc;
      g_enq2 += t_enq2;
      g_wait2 += Secs(t_acc) - t_enq2;
      ++g_n;
      ++sum->num_successful_steps;
      trace(cost, radius, 1);
      if (gmax <= opt.gradient_tolerance) {
        sum->termination_type = 0;
        done = true;
      }
    } else {
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      ++sum->num_unsuccessful_steps;
      trace(cost, radius, 0);
      if (radius < opt.min_trust_region_radius) {
        sum->termination_type = 0;
        done = true;
      }
    }
  }
  if (gap_trace && g_n > 0)
    std::fprintf(stderr, "[ba gaps] per iteration (us): enqueue solve batch %.0f, wait %.0f | "
                 "enqueue linearise batch %.0f, wait (incl. gradient launch) %.0f\n",
                 1e6 * g_enq1 / g_n, 1e6 * g_wait1 / g_n, 1e6 * g_enq2 / g_n, 1e6 * g_wait2 / g_n);
  PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
  PPSFM_CUDA(ctx, cudaGetLastError());
  drain_lin_events();
  sum->final_cost = cost;
  sum->final_gradient_max_norm = gmax;
  sum->trace_len = tl;
  sum->total_time_s = Secs(t_start);
  sum->linear_solver_time_s = solver_s;
  sum->jacobian_time_s = st->lin_ms * 1e-3;
  sum->jacobian_launches = (int32_t)st->lin_launches;
  sum->kernel_launches = st->launches;
  return PPSFM_OK;
}

int BaDownload(BaState* st, const ppsfm_ba_problem* pb) {
  ppsfm_ctx* ctx = st->ctx;
  cudaStream_t s = ctx->stream;
  const BaDev& d = st->d;
  // poses of every image (constant ones come back unchanged); points owned by this rank
  PPSFM_CUDA(ctx, cudaMemcpyAsync(pb->qvecs, d.q, sizeof(double) * 4 * d.C, cudaMemcpyDeviceToHost, s));
  PPSFM_CUDA(ctx, cudaMemcpyAsync(pb->tvecs, d.t, sizeof(double) * 3 * d.C, cudaMemcpyDeviceToHost, s));
  if (d.NCv > 0)  // Camera::Params() of the refined cameras (replicated on every rank)
    PPSFM_CUDA(ctx, cudaMemcpyAsync(pb->camera_params, d.cam_params,
                                    sizeof(double) * 12 * d.num_cameras, cudaMemcpyDeviceToHost, s));
  if (st->world > 1) {
    // every rank moved only its own points (local index p = the caller's p * world + rank):
    // scatter the displacements into a zeroed whole-problem buffer and sum it over the ranks, so
    // that all ranks return the complete point set (once per solve, not per iteration)
    const size_t ng = 3 * (size_t)st->P_global;
    PPSFM_CUDA(ctx, cudaMemsetAsync(st->Xg, 0, sizeof(double) * std::max<size_t>(ng, 1), s));
    launch_scatter_displacement(st->Xg, d.X, st->X0, d.P, st->world, st->rank, s);
    const int rc = CommAllReduce(ctx, st->Xg, ng, false);
    if (rc != PPSFM_OK) return rc;
    std::vector<double> disp(ng);
    PPSFM_CUDA(ctx, cudaMemcpyAsync(disp.data(), st->Xg, sizeof(double) * ng, cudaMemcpyDeviceToHost, s));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
    for (size_t i = 0; i < ng; ++i) pb->points[i] = st->points_in[i] + disp[i];
    return PPSFM_OK;
  }
  PPSFM_CUDA(ctx, cudaMemcpyAsync(pb->points, d.X, sizeof(double) * 3 * d.P, cudaMemcpyDeviceToHost, s));
  PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
  return PPSFM_OK;
}

int BaReset(BaState* st) {
  ppsfm_ctx* ctx = st->ctx;
  cudaStream_t s = ctx->stream;
  BaDev& d = st->d;
  PPSFM_CUDA(ctx, cudaMemcpyAsync(d.q, st->q0, sizeof(double) * 4 * d.C, cudaMemcpyDeviceToDevice, s));
  PPSFM_CUDA(ctx, cudaMemcpyAsync(d.t, st->t0, sizeof(double) * 3 * d.C, cudaMemcpyDeviceToDevice, s));
  PPSFM_CUDA(ctx, cudaMemcpyAsync(d.X, st->X0, sizeof(double) * 3 * d.P, cudaMemcpyDeviceToDevice, s));
  if (d.NCv > 0) {
    PPSFM_CUDA(ctx, cudaMemcpyAsync(d.cam_params, st->cam_params0,
                                    sizeof(double) * 12 * d.num_cameras, cudaMemcpyDeviceToDevice, s));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(d.img_params, st->img_params0, sizeof(double) * 12 * d.C,
                                    cudaMemcpyDeviceToDevice, s));
  }
  PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
  return PPSFM_OK;
}

namespace {
int AllReduceSum(BaState* st, double* dev, size_t count) {
  return CommAllReduce(st->ctx, dev, count, false);
}
}  // namespace

}  // namespace ppsfm

// ============================================================================================
// C-ABI
// ============================================================================================
using namespace ppsfm;

extern "C" {

void ppsfm_ba_options_default(ppsfm_ba_options* o) {
  if (!o) return;
  // BundleAdjustmentOptions() (src/optim/bundle_adjustment.h:49-93) + ceres::Solver::Options
  o->loss_type = 0;
  o->loss_scale = 1.0;
  o->max_num_iterations = 100;
  o->function_tolerance = 0.0;
  o->gradient_tolerance = 0.0;
  o->parameter_tolerance = 0.0;
  o->max_num_consecutive_invalid_steps = 10;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->jacobi_scaling = 1;
  o->num_threads = -1;
  o->refine_focal_length = 0;     // intrinsics constant: what the mapper of the reference runs
  o->refine_principal_point = 0;  // with (controllers/incremental_mapper.h:81-83) and the default
  o->refine_extra_params = 0;     // of the pose refinement (estimators/pose.h:84-101)
}

int ppsfm_ba_create(ppsfm_ctx* ctx, const ppsfm_ba_problem* problem,
                    const ppsfm_ba_options* options, ppsfm_ba** out) {
  if (!ctx) return PPSFM_ERR_INVALID;
  cudaSetDevice(ctx->device);
  BaState* st = nullptr;
  const int rc = BaCreate(ctx, problem, options, ctx->rank, ctx->world, &st);
  if (out) *out = reinterpret_cast<ppsfm_ba*>(st);
  return rc;
}

int ppsfm_ba_run(ppsfm_ba* ba, ppsfm_ba_summary* summary) {
  BaState* st = reinterpret_cast<BaState*>(ba);
  if (!st || !summary) return PPSFM_ERR_INVALID;
  cudaSetDevice(st->ctx->device);
  return BaRun(st, summary);
}

int ppsfm_ba_reset(ppsfm_ba* ba) {
  BaState* st = reinterpret_cast<BaState*>(ba);
  if (!st) return PPSFM_ERR_INVALID;
  cudaSetDevice(st->ctx->device);
  return BaReset(st);
}

int ppsfm_ba_download(ppsfm_ba* ba, const ppsfm_ba_problem* problem) {
  BaState* st = reinterpret_cast<BaState*>(ba);
  if (!st || !problem) return PPSFM_ERR_INVALID;
  cudaSetDevice(st->ctx->device);
  return BaDownload(st, problem);
}

void ppsfm_ba_free(ppsfm_ba* ba) {
  BaState* st = reinterpret_cast<BaState*>(ba);
  if (!st) return;
  cudaSetDevice(st->ctx->device);
  BaFree(st);
}

int ppsfm_ba_solve(ppsfm_ctx* ctx, const ppsfm_ba_problem* problem,
                   const ppsfm_ba_options* options, ppsfm_ba_summary* summary) {
  static const bool timing = tune_int("PPSFM_BA_TIMING", 0) != 0;
  const auto t0 = std::chrono::steady_clock::now();
  ppsfm_ba* ba = nullptr;
  int rc = ppsfm_ba_create(ctx, problem, options, &ba);
  if (rc != PPSFM_OK) return rc;
  const double t_create = Secs(t0);
  rc = ppsfm_ba_run(ba, summary);
  const double t_run = Secs(t0);
  if (rc == PPSFM_OK) {
    const int rc2 = ppsfm_ba_download(ba, problem);
    if (rc2 != PPSFM_OK) rc = rc2;
  }
  const double t_down = Secs(t0);
  ppsfm_ba_free(ba);
  if (timing)
    std::fprintf(stderr, "[ba] solve: create %.2f, run %.2f, download %.2f, free %.2f ms\n",
                 1e3 * t_create, 1e3 * (t_run - t_create), 1e3 * (t_down - t_run),
                 1e3 * (Secs(t0) - t_down));
  return rc;
}

// Host-only: how a problem is dealt to `rank` of `world` ranks (no GPU needed; used by the
// world_size-2 gloo tests).  out = {kept observations on this rank, points owned by this rank
// that have kept observations, camera blocks (identical on every rank), kept observations in total}.
int ppsfm_ba_shard_stats(const ppsfm_ba_problem* pb, int rank, int world, int64_t* out) {
  if (!pb || !out || world < 1 || rank < 0 || rank >= world) return PPSFM_ERR_INVALID;
  const int C = pb->num_images, P = pb->num_points;
  std::vector<uint8_t> cam_used(C, 0), pt_has(P, 0);
  int64_t local = 0, total = 0;
  for (int64_t o = 0; o < pb->num_obs; ++o) {
    const int ci = pb->obs_image[o], pi = pb->obs_point[o];
    if (ci < 0 || ci >= C || pi < 0 || pi >= P) return PPSFM_ERR_INVALID;
    const bool cc = pb->pose_flags && (pb->pose_flags[ci] & 1);
    const bool pc = pb->point_const && pb->point_const[pi];
    if (cc && pc) continue;
    cam_used[ci] = 1;
    ++total;
    if (world == 1 || (pi % world) == rank) {
      ++local;
      pt_has[pi] = 1;
    }
  }
  int64_t blocks = 0, pts = 0;
  for (int i = 0; i < C; ++i)
    if (cam_used[i] && !(pb->pose_flags && (pb->pose_flags[i] & 1))) ++blocks;
  for (int i = 0; i < P; ++i) pts += pt_has[i];
  out[0] = local; out[1] = pts; out[2] = blocks; out[3] = total;
  return PPSFM_OK;
}

// Test hook: residuals and tangent-space Jacobian blocks of every observation at the current
// state, unscaled, with the loss correction of `options` (row-major 2, 2x6, 2x3 per observation,
// in the order of the input observations; all-constant blocks come back as zeros).
int ppsfm_ba_linearize(ppsfm_ctx* ctx, const ppsfm_ba_problem* problem,
                       const ppsfm_ba_options* options, double* residuals, double* jac_cam,
                       double* jac_point, double* cost) {
  ppsfm_ba* ba = nullptr;
  int rc = ppsfm_ba_create(ctx, problem, options, &ba);
  if (rc != PPSFM_OK) return rc;
  BaState* st = reinterpret_cast<BaState*>(ba);
  BaDev& d = st->d;
  cudaStream_t s = ctx->stream;
  auto body = [&]() -> int {
    std::vector<double> ones(std::max<size_t>(6 * (size_t)d.NB, 3 * (size_t)d.P), 1.0);
    PPSFM_CUDA(ctx, cudaMemcpy(d.cam_scale, ones.data(), sizeof(double) * 6 * d.NB,
                               cudaMemcpyHostToDevice));
    PPSFM_CUDA(ctx, cudaMemcpy(d.pt_scale, ones.data(), sizeof(double) * 3 * d.P,
                               cudaMemcpyHostToDevice));
    launch_linearize(d, d.q, d.t, d.X, true, BaLoss{options->loss_type, options->loss_scale}, s);
    const int64_t K = d.K;
    std::vector<double> J(ba_j_doubles(K));
    std::vector<int> oc(K), op(K);
    PPSFM_CUDA(ctx, cudaMemcpyAsync(J.data(), d.J, sizeof(double) * J.size(), cudaMemcpyDeviceToHost, s));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(oc.data(), d.obs_cam, sizeof(int) * K, cudaMemcpyDeviceToHost, s));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(op.data(), d.obs_pt, sizeof(int) * K, cudaMemcpyDeviceToHost, s));
    double c = 0;
    PPSFM_CUDA(ctx, cudaMemcpyAsync(&c, d.scalars + kCost, sizeof(double), cudaMemcpyDeviceToHost, s));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
    PPSFM_CUDA(ctx, cudaGetLastError());
    if (cost) *cost = c;
    // kept observations are point-major and stable in input order: replay the same filter
    const int64_t O = problem->num_obs;
    std::vector<int64_t> next(problem->num_points + 1, 0);
    std::vector<int64_t> pt_start(problem->num_points + 1, 0);
    auto kept = [&](int64_t o) {
      const int ci = problem->obs_image[o], pi = problem->obs_point[o];
      const bool cc = problem->pose_flags && (problem->pose_flags[ci] & 1);
      const bool pc = problem->point_const && problem->point_const[pi];
      const bool ic = VariableIntrinsics(problem, options, problem->image_camera[ci]) != 0;
      return !(cc && pc && !ic);
    };
    for (int64_t o = 0; o < O; ++o)
      if (kept(o)) pt_start[problem->obs_point[o] + 1]++;
    for (int i = 0; i < problem->num_points; ++i) pt_start[i + 1] += pt_start[i];
    std::copy(pt_start.begin(), pt_start.end(), next.begin());
    for (int64_t o = 0; o < O; ++o) {
      double* ro = residuals + 2 * o;
      double* co = jac_cam + 12 * o;
      double* po = jac_point + 6 * o;
      if (!kept(o)) {
        std::fill(ro, ro + 2, 0.0);
        std::fill(co, co + 12, 0.0);
        std::fill(po, po + 6, 0.0);
        continue;
      }
      const int64_t k = next[problem->obs_point[o]]++;
      for (int i = 0; i < 2; ++i) ro[i] = J[ba_jidx(i, k)];
      for (int i = 0; i < 12; ++i) co[i] = J[ba_jidx(2 + i, k)];
      for (int i = 0; i < 6; ++i) po[i] = J[ba_jidx(14 + i, k)];
    }
    return PPSFM_OK;
  };
  rc = body();
  ppsfm_ba_free(ba);
  return rc;
}

// Test hook: solves A x = b for a dense SPD matrix (row-major n x n, host) with the reduced-
// camera-system solver (dense_chol.cu).  Returns PPSFM_NO_SOLUTION if A is not positive definite.
int ppsfm_dense_cholesky_solve(ppsfm_ctx* ctx, const double* A, int n, const double* b,
                               double* x) {
  if (!ctx || !A || !b || !x || n <= 0) return fail(ctx, PPSFM_ERR_INVALID, "bad argument");
  cudaSetDevice(ctx->device);
  const int ld = chol_ld(n);
  std::vector<double> h((size_t)ld * ld, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) h[(size_t)i * ld + j] = A[(size_t)i * n + j];
  for (int j = 0; j < n; ++j) h[(size_t)n * ld + j] = b[j];
  double *dA = nullptr, *dx = nullptr, *dwork = nullptr;
  int* dst = nullptr;
  PPSFM_CUDA(ctx, cudaMalloc(&dA, sizeof(double) * h.size()));
  PPSFM_CUDA(ctx, cudaMalloc(&dwork, sizeof(double) * chol_work_doubles(n)));
  PPSFM_CUDA(ctx, cudaMalloc(&dx, sizeof(double) * n));
  PPSFM_CUDA(ctx, cudaMalloc(&dst, sizeof(int)));
  cudaStream_t s = ctx->stream;
  int status = 0;
  auto body = [&]() -> int {
    PPSFM_CUDA(ctx, cudaMemcpyAsync(dA, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, s));
    chol_solve_bordered(dA, n, ld, dx, dwork, dst, s);
    PPSFM_CUDA(ctx, cudaMemcpyAsync(x, dx, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    PPSFM_CUDA(ctx, cudaMemcpyAsync(&status, dst, sizeof(int), cudaMemcpyDeviceToHost, s));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
    PPSFM_CUDA(ctx, cudaGetLastError());
    return PPSFM_OK;
  };
  int rc = body();
  cudaFree(dA);
  cudaFree(dx);
  cudaFree(dwork);
  cudaFree(dst);
  if (rc == PPSFM_OK && status) rc = PPSFM_NO_SOLUTION;
  return rc;
}

// RefineAbsolutePoseFromLines (src/estimators/pose.cc:96-213).  refine_focal_length /
// refine_extra_params select the variable groups of camera->Params() (pose.cc:149-183; the
// principal point always stays fixed); camera_params is updated in place when one is set.
int ppsfm_refine_absolute_pose_from_lines_ex(ppsfm_ctx* ctx, const uint8_t* inlier_mask,
                                             const double* lines, const double* points, size_t n,
                                             int camera_model, double* camera_params,
                                             int refine_focal_length, int refine_extra_params,
                                             double gradient_tolerance, int max_num_iterations,
                                             double loss_function_scale, double* qvec,
                                             double* tvec, ppsfm_ba_summary* summary) {
  if (!ctx || !inlier_mask || !lines || !points || !qvec || !tvec || !camera_params)
    return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  if (camera_model < 0 || camera_model > 10)
    return fail(ctx, PPSFM_ERR_INVALID, "camera model %d not supported", camera_model);
  // AbsolutePoseRefinementOptions::Check (pose.h:103-107)
  if (gradient_tolerance < 0 || max_num_iterations < 0 || loss_function_scale < 0)
    return fail(ctx, PPSFM_ERR_INVALID, "AbsolutePoseRefinementOptions::Check failed");
  std::vector<int32_t> oi, op;
  std::vector<double> ol, pts;
  for (size_t i = 0; i < n; ++i) {
    if (!inlier_mask[i]) continue;
    oi.push_back(0);
    op.push_back((int32_t)(pts.size() / 3));
    for (int k = 0; k < 3; ++k) {
      ol.push_back(lines[3 * i + k]);
      pts.push_back(points[3 * i + k]);
    }
  }
  const int np = (int)(pts.size() / 3);
  ppsfm_ba_summary local;
  ppsfm_ba_summary* sum = summary ? summary : &local;
  std::memset(sum, 0, sizeof(*sum));
  if (np == 0) return PPSFM_OK;  // empty problem: Ceres converges immediately, "usable"
  // *qvec = NormalizeQuaternion(*qvec)  (pose.cc:143)
  const double nrm = std::sqrt(qvec[0] * qvec[0] + qvec[1] * qvec[1] + qvec[2] * qvec[2] + qvec[3] * qvec[3]);
  if (nrm > 0) for (int k = 0; k < 4; ++k) qvec[k] /= nrm;
  std::vector<uint8_t> pc(np, 1);
  uint8_t flags = 0;
  int32_t icam = 0, model = camera_model;
  double params[12] = {0};
  const int np_model = kNumParams[camera_model];
  for (int k = 0; k < np_model; ++k) params[k] = camera_params[k];
  ppsfm_ba_problem pb;
  pb.num_images = 1; pb.qvecs = qvec; pb.tvecs = tvec; pb.pose_flags = &flags;
  pb.image_camera = &icam; pb.num_cameras = 1; pb.camera_model = &model;
  pb.camera_params = params; pb.num_points = np; pb.points = pts.data();
  pb.point_const = pc.data(); pb.num_obs = np; pb.obs_image = oi.data();
  pb.obs_point = op.data(); pb.obs_line = ol.data();
  pb.camera_const = nullptr;
  ppsfm_ba_options o;
  ppsfm_ba_options_default(&o);
  o.loss_type = 2;  // ceres::CauchyLoss(options.loss_function_scale)  (pose.cc:106-107)
  o.loss_scale = loss_function_scale;
  o.gradient_tolerance = gradient_tolerance;
  o.max_num_iterations = max_num_iterations;
  o.function_tolerance = 1e-6;   // ceres::Solver::Options defaults: pose.cc:187-190 overrides
  o.parameter_tolerance = 1e-8;  // only gradient_tolerance / max_num_iterations / solver type
  o.refine_focal_length = refine_focal_length ? 1 : 0;
  o.refine_extra_params = refine_extra_params ? 1 : 0;
  const int rc = ppsfm_ba_solve(ctx, &pb, &o, sum);
  if (rc != PPSFM_OK) return rc;
  if (o.refine_focal_length || o.refine_extra_params)
    for (int k = 0; k < np_model; ++k) camera_params[k] = params[k];
  return sum->termination_type != 2 ? PPSFM_OK : PPSFM_NO_SOLUTION;  // IsSolutionUsable()
}

// The same with constant intrinsics (refine_focal_length = refine_extra_params = false, the
// defaults of pose.h:84-101).
int ppsfm_refine_absolute_pose_from_lines(ppsfm_ctx* ctx, const uint8_t* inlier_mask,
                                          const double* lines, const double* points, size_t n,
                                          int camera_model, const double* camera_params,
                                          double gradient_tolerance, int max_num_iterations,
                                          double loss_function_scale, double* qvec, double* tvec,
                                          ppsfm_ba_summary* summary) {
  if (!camera_params) return fail(ctx, PPSFM_ERR_INVALID, "null argument");
  double params[12] = {0};
  if (camera_model >= 0 && camera_model <= 10)
    for (int k = 0; k < kNumParams[camera_model]; ++k) params[k] = camera_params[k];
  return ppsfm_refine_absolute_pose_from_lines_ex(ctx, inlier_mask, lines, points, n, camera_model,
                                                  params, 0, 0, gradient_tolerance,
                                                  max_num_iterations, loss_function_scale, qvec,
                                                  tvec, summary);
}

}  // extern "C"
