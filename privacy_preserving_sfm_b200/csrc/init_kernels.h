// init_kernels.h — device-side scoring of the four-view initialisation's candidate models.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "../cpp/ppsfm_init_math.h"

struct ppsfm_ctx;

namespace ppsfm {

// Hands the estimators of cpp/ppsfm_init.h their GPU scorers (observations uploaded once per
// estimator) and owns them.
class GpuScorerFactory : public init::BatchScorerFactory {
 public:
  explicit GpuScorerFactory(ppsfm_ctx* ctx);
  ~GpuScorerFactory() override;
  const init::BatchScorer* FourView2d(const double* const* x, int n) override;
  const init::BatchScorer* PlanarOffset(const double* const* lines, int n) override;
  cudaError_t error() const { return error_; }
  int64_t launches() const;

 private:
  const init::BatchScorer* Make(const double* const* obs, int n, bool is3d);
  ppsfm_ctx* ctx_;
  std::vector<init::BatchScorer*> made_;
  cudaError_t error_ = cudaSuccess;
};

}  // namespace ppsfm
