// init_kernels.cu — GPU scoring of the candidate models of the four-view initialisation
// (SURVEY.md §8 f4): FourView2dEstimator (src/init/sfm2d.cc:302-444) and PlanarOffsetEstimator
// (src/init/initializer.cc:219-333) under ransac_lib::LocallyOptimizedMSAC.
//
// What is data parallel in that loop is per candidate model: EVERY track is triangulated with the
// model's cameras (a 3x2 or 4x3 least-squares solve) and evaluated (EvaluateModelOnPoint); a
// minimal sample of the 2-D solver yields up to 16 models, the driver scores each against all
// tracks, 1000+ samples per run.  The LO-MSAC control flow (sampling, minimal solvers — a chain
// of small SVDs —, local optimisation, adaptive stopping) stays on the host; per sample ONE launch
// scores all its models:
//   init_score_kernel   CTA / model, thread / track: triangulate + evaluate with the arithmetic of
//                       cpp/ppsfm_init_math.h (shared with the host estimators, --fmad=false),
//                       min(error, threshold) to HBM; then ONE thread adds them up in track
//                       order — the MSAC score the host loop would compute, bit for bit, so the
//                       sequence of accepted models, and with it every later random draw, is that
//                       of the host-only run.
// Candidate models stay "lazy" (cameras only); the host triangulates the tracks of the few models
// that become the best so far (cpp/ppsfm_init.h: ScoreModels / Materialize).
#include <cstring>

#include "../cpp/ppsfm_init_math.h"
#include "common.h"
#include "init_kernels.h"

namespace ppsfm {
namespace {

constexpr int kMaxModels = 32;  // per launch (a 2-D sample yields <= 16)

// obs: 2-D: four views x n x 2 unit observations; 3-D: four views x n x 3 lines.
template <bool k3d>
__global__ void __launch_bounds__(256)
init_score_kernel(const double* __restrict__ obs, int n, const double* __restrict__ cams,
                  double threshold, int threshold_first, double* __restrict__ err,
                  double* __restrict__ scores) {
  constexpr int kCam = k3d ? 48 : 24, kObs = k3d ? 3 : 2;
  __shared__ double cam[kCam];
  const int m = blockIdx.x;
  if (threadIdx.x < kCam) cam[threadIdx.x] = cams[(size_t)m * kCam + threadIdx.x];
  __syncthreads();
  double* e_out = err + (size_t)m * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double o[4][kObs];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int c = 0; c < kObs; ++c) o[v][c] = obs[((size_t)v * n + i) * kObs + c];
    double e;
    if (k3d) {
      double X[3];
      init::hd::triangulate3d_point(cam, o[0], o[1], o[2], o[3], X);
      e = init::hd::planar_offset_error(cam, o[0], o[1], o[2], o[3], X);
    } else {
      double X[2];
      init::hd::triangulate2d_point(cam, cam + 6, cam + 12, o[0], o[1], o[2], X);
      e = init::hd::fourview2d_error(cam, o[0], o[1], o[2], o[3], X);
    }
    // std::min(e, threshold) resp. std::min(threshold, e): the same but for a NaN error
    e_out[i] = threshold_first ? ((e < threshold) ? e : threshold) : ((threshold < e) ? threshold : e);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // the host's summation order
    double s = 0.0;
#pragma unroll 8
    for (int i = 0; i < n; ++i) s += e_out[i];
    scores[m] = s;
  }
}

class GpuBatchScorer : public init::BatchScorer {
 public:
  GpuBatchScorer(ppsfm_ctx* ctx, bool is3d) : ctx_(ctx), is3d_(is3d) {}
  ~GpuBatchScorer() override {
    cudaSetDevice(ctx_->device);
    d_obs_.release();
    d_cams_.release();
    d_err_.release();
    d_scores_.release();
    h_cams_.release();
    h_scores_.release();
  }
  // obs: four arrays of n x (2 | 3) doubles
  cudaError_t Upload(const double* const* obs, int n) {
    n_ = n;
    const size_t per = (size_t)n * (is3d_ ? 3 : 2);
    cudaError_t e = d_obs_.reserve(sizeof(double) * 4 * per + 16);
    if (e == cudaSuccess) e = d_cams_.reserve(sizeof(double) * 48 * kMaxModels);
    if (e == cudaSuccess) e = d_err_.reserve(sizeof(double) * (size_t)kMaxModels * (n > 0 ? n : 1));
    if (e == cudaSuccess) e = d_scores_.reserve(sizeof(double) * kMaxModels);
    if (e == cudaSuccess) e = h_cams_.reserve(sizeof(double) * 48 * kMaxModels);
    if (e == cudaSuccess) e = h_scores_.reserve(sizeof(double) * kMaxModels);
    for (int v = 0; v < 4 && e == cudaSuccess && per > 0; ++v)
      e = cudaMemcpyAsync(d_obs_.as<double>() + v * per, obs[v], sizeof(double) * per,
                          cudaMemcpyHostToDevice, ctx_->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx_->stream);
    return e;
  }
  bool ScoreFourView2d(const double* cams, int num_models, double threshold, bool threshold_first,
                       double* scores) const override {
    return !is3d_ && Score(cams, num_models, threshold, threshold_first, scores);
  }
  bool ScorePlanarOffset(const double* cams, int num_models, double threshold,
                         bool threshold_first, double* scores) const override {
    return is3d_ && Score(cams, num_models, threshold, threshold_first, scores);
  }
  int64_t launches() const { return launches_; }

 private:
  bool Score(const double* cams, int num_models, double threshold, bool threshold_first,
             double* scores) const {
    if (n_ <= 0) return false;
    const int kc = is3d_ ? 48 : 24;
    cudaStream_t s = ctx_->stream;
    for (int m0 = 0; m0 < num_models; m0 += kMaxModels) {
      const int mc = num_models - m0 < kMaxModels ? num_models - m0 : kMaxModels;
      std::memcpy(h_cams_.p, cams + (size_t)m0 * kc, sizeof(double) * kc * mc);
      cudaError_t e = cudaMemcpyAsync(d_cams_.p, h_cams_.p, sizeof(double) * kc * mc,
                                      cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess) return false;
      if (is3d_)
        init_score_kernel<true><<<mc, 256, 0, s>>>(d_obs_.as<double>(), n_, d_cams_.as<double>(),
                                                   threshold, threshold_first ? 1 : 0,
                                                   d_err_.as<double>(), d_scores_.as<double>());
      else
        init_score_kernel<false><<<mc, 256, 0, s>>>(d_obs_.as<double>(), n_, d_cams_.as<double>(),
                                                    threshold, threshold_first ? 1 : 0,
                                                    d_err_.as<double>(), d_scores_.as<double>());
      ++launches_;
      e = cudaMemcpyAsync(h_scores_.p, d_scores_.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) return false;  // the caller falls back to nothing: see init_host.cu
      std::memcpy(scores + m0, h_scores_.p, sizeof(double) * mc);
    }
    return true;
  }
  ppsfm_ctx* ctx_;
  bool is3d_;
  int n_ = 0;
  DevBuf d_obs_, d_cams_, d_err_, d_scores_;
  PinBuf h_cams_, h_scores_;
  mutable int64_t launches_ = 0;
};

}  // namespace

GpuScorerFactory::GpuScorerFactory(ppsfm_ctx* ctx) : ctx_(ctx) {}
GpuScorerFactory::~GpuScorerFactory() {
  for (init::BatchScorer* s : made_) delete s;
}
const init::BatchScorer* GpuScorerFactory::Make(const double* const* obs, int n, bool is3d) {
  GpuBatchScorer* s = new GpuBatchScorer(ctx_, is3d);
  made_.push_back(s);
  error_ = s->Upload(obs, n);
  return error_ == cudaSuccess ? s : nullptr;
}
const init::BatchScorer* GpuScorerFactory::FourView2d(const double* const* x, int n) {
  return Make(x, n, false);
}
const init::BatchScorer* GpuScorerFactory::PlanarOffset(const double* const* lines, int n) {
  return Make(lines, n, true);
}
int64_t GpuScorerFactory::launches() const {
  int64_t t = 0;
  for (const init::BatchScorer* s : made_) t += static_cast<const GpuBatchScorer*>(s)->launches();
  return t;
}

}  // namespace ppsfm
