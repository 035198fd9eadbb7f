// Stand-in for <ceres/ceres.h> + <ceres/rotation.h> — TEST INFRASTRUCTURE ONLY (oracle/build_ref.sh).
// Ceres is absent in this image (and unpinned upstream).  The reference's cost functors
// (src/base/cost_functions.h) and camera models (src/base/camera_models.h) need four things from
// it: the scalar functions ceres::sqrt / atan / tan / sin / cos, the dual-number type their
// templates are instantiated with by AutoDiffCostFunction (ceres::Jet), UnitQuaternionRotatePoint,
// and the CostFunction / AutoDiffCostFunction class names of their Create() factories.  These are
// restated here from Ceres' published sources (1.14 - 2.0; from memory, the library is not in
// the image): what tests/test_ref_cost.py pins is the reference's functor and camera-model TEXT,
// not Ceres.
#pragma once
#include <cmath>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include <glog/logging.h>  // (real Ceres headers pull glog in; cost_functions.h relies on it for CHECK_NEAR)

namespace ceres {

using std::abs;
using std::atan;
using std::cos;
using std::sin;
using std::sqrt;
using std::tan;

// ceres/jet.h: a + sum_i v[i] e_i, e_i e_j = 0
template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a() {
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(const T& value) : a(value) {  // NOLINT (implicit like ceres::Jet)
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(int value) : a(value) {  // NOLINT: T(2) in the functors
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(const T& value, int k) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = T();
    v[k] = T(1);
  }
  Jet& operator+=(const Jet& y) { return *this = *this + y; }
  Jet& operator-=(const Jet& y) { return *this = *this - y; }
  Jet& operator*=(const Jet& y) { return *this = *this * y; }
  Jet& operator/=(const Jet& y) { return *this = *this / y; }
};

template <typename T, int N>
Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
template <typename T, int N>
Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = -f.a;
  for (int i = 0; i < N; ++i) r.v[i] = -f.v[i];
  return r;
}
template <typename T, int N>
Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r;
  r.a = f.a + g.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] + g.v[i];
  return r;
}
template <typename T, int N>
Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r;
  r.a = f.a - g.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] - g.v[i];
  return r;
}
template <typename T, int N>
Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r;
  r.a = f.a * g.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return r;
}
template <typename T, int N>
Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  // ceres/jet.h: one reciprocal, b/c = b * c^-1 and (db - (b/c) dc) * c^-1
  Jet<T, N> r;
  const T g_a_inverse = T(1.0) / g.a;
  r.a = f.a * g_a_inverse;
  for (int i = 0; i < N; ++i) r.v[i] = (f.v[i] - r.a * g.v[i]) * g_a_inverse;
  return r;
}
// mixed with scalars
template <typename T, int N>
Jet<T, N> operator+(const Jet<T, N>& f, T s) { return f + Jet<T, N>(s); }
template <typename T, int N>
Jet<T, N> operator+(T s, const Jet<T, N>& f) { return Jet<T, N>(s) + f; }
template <typename T, int N>
Jet<T, N> operator-(const Jet<T, N>& f, T s) { return f - Jet<T, N>(s); }
template <typename T, int N>
Jet<T, N> operator-(T s, const Jet<T, N>& f) { return Jet<T, N>(s) - f; }
template <typename T, int N>
Jet<T, N> operator*(const Jet<T, N>& f, T s) { return f * Jet<T, N>(s); }
template <typename T, int N>
Jet<T, N> operator*(T s, const Jet<T, N>& f) { return Jet<T, N>(s) * f; }
template <typename T, int N>
Jet<T, N> operator/(const Jet<T, N>& f, T s) { return f / Jet<T, N>(s); }
template <typename T, int N>
Jet<T, N> operator/(T s, const Jet<T, N>& f) { return Jet<T, N>(s) / f; }

#define MINICERES_CMP(op)                                                                 \
  template <typename T, int N>                                                            \
  bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; }         \
  template <typename T, int N>                                                            \
  bool operator op(const Jet<T, N>& f, T s) { return f.a op s; }                          \
  template <typename T, int N>                                                            \
  bool operator op(T s, const Jet<T, N>& f) { return s op f.a; }
MINICERES_CMP(<)
MINICERES_CMP(<=)
MINICERES_CMP(>)
MINICERES_CMP(>=)
MINICERES_CMP(==)
MINICERES_CMP(!=)
#undef MINICERES_CMP

template <typename T, int N>
Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
template <typename T, int N>
Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::sqrt(f.a);
  const T two_a_inverse = T(1.0) / (T(2.0) * r.a);
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * two_a_inverse;
  return r;
}
template <typename T, int N>
Jet<T, N> atan(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::atan(f.a);
  const T tmp = T(1.0) / (T(1.0) + f.a * f.a);
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * tmp;
  return r;
}
template <typename T, int N>
Jet<T, N> tan(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::tan(f.a);
  const T tmp = T(1.0) + r.a * r.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * tmp;
  return r;
}
template <typename T, int N>
Jet<T, N> sin(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::sin(f.a);
  const T c = std::cos(f.a);
  for (int i = 0; i < N; ++i) r.v[i] = c * f.v[i];
  return r;
}
template <typename T, int N>
Jet<T, N> cos(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::cos(f.a);
  const T s = -std::sin(f.a);
  for (int i = 0; i < N; ++i) r.v[i] = s * f.v[i];
  return r;
}

// ceres/rotation.h (1.14 - 2.0): rotation by a UNIT quaternion (w, x, y, z)
template <typename T>
inline void UnitQuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  const T t2 = q[0] * q[1];
  const T t3 = q[0] * q[2];
  const T t4 = q[0] * q[3];
  const T t5 = -q[1] * q[1];
  const T t6 = q[1] * q[2];
  const T t7 = q[1] * q[3];
  const T t8 = -q[2] * q[2];
  const T t9 = q[2] * q[3];
  const T t1 = -q[3] * q[3];
  result[0] = T(2) * ((t8 + t1) * pt[0] + (t6 - t4) * pt[1] + (t3 + t7) * pt[2]) + pt[0];
  result[1] = T(2) * ((t4 + t6) * pt[0] + (t5 + t1) * pt[1] + (t9 - t2) * pt[2]) + pt[1];
  result[2] = T(2) * ((t7 - t3) * pt[0] + (t2 + t9) * pt[1] + (t5 + t8) * pt[2]) + pt[2];
}

// ---------------------------------------------------------------------------------------------
// The solver-facing classes, as a RECORDER: ceres::Problem keeps what the reference's
// BundleAdjuster::SetUp adds to it (residual blocks with their cost-function kind and parameter
// blocks, constant blocks, parameterisations), ceres::Solve stores the options it was handed and
// solves nothing.  oracle/ref/ref_ba_setup.cc reads the record back; tests/test_ref_ba_setup.py
// compares it with what the oracle and the product assemble for the same configuration.
// ---------------------------------------------------------------------------------------------
#define CERES_VERSION_MAJOR 2
#define CERES_VERSION_MINOR 0

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR,
                        SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum PreconditionerType { IDENTITY, JACOBI, SCHUR_JACOBI, CLUSTER_JACOBI, CLUSTER_TRIDIAGONAL };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual int Kind() const = 0;  // 0 trivial, 1 soft-L1, 2 Cauchy (the recorder's numbering)
  virtual double Scale() const { return 0.0; }
};
class TrivialLoss : public LossFunction {
 public:
  int Kind() const override { return 0; }
};
class SoftLOneLoss : public LossFunction {
  double a_;

 public:
  explicit SoftLOneLoss(double a) : a_(a) {}
  int Kind() const override { return 1; }
  double Scale() const override { return a_; }
};
class CauchyLoss : public LossFunction {
  double a_;

 public:
  explicit CauchyLoss(double a) : a_(a) {}
  int Kind() const override { return 2; }
  double Scale() const override { return a_; }
};

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
  virtual const std::vector<int>* ConstantIndices() const { return nullptr; }
};
class QuaternionParameterization : public LocalParameterization {
 public:
  int GlobalSize() const override { return 4; }
  int LocalSize() const override { return 3; }
};
class SubsetParameterization : public LocalParameterization {
  int size_;
  std::vector<int> constant_;

 public:
  SubsetParameterization(int size, const std::vector<int>& constant_parameters)
      : size_(size), constant_(constant_parameters) {}
  int GlobalSize() const override { return size_; }
  int LocalSize() const override { return size_ - static_cast<int>(constant_.size()); }
  const std::vector<int>* ConstantIndices() const override { return &constant_; }
};

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual int NumResiduals() const = 0;
  virtual const std::vector<int>& ParameterBlockSizes() const = 0;
  virtual const void* FunctorAddress() const = 0;
};
template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
  std::unique_ptr<Functor> functor_;
  std::vector<int> sizes_;

 public:
  explicit AutoDiffCostFunction(Functor* functor) : functor_(functor), sizes_{Ns...} {}
  int NumResiduals() const override { return kNumResiduals; }
  const std::vector<int>& ParameterBlockSizes() const override { return sizes_; }
  const void* FunctorAddress() const override { return functor_.get(); }
  const Functor& functor() const { return *functor_; }
};

class Problem {
 public:
  struct ResidualBlock {
    std::unique_ptr<CostFunction> cost;
    const LossFunction* loss;
    std::vector<double*> blocks;
  };
  std::vector<ResidualBlock> residual_blocks;
  std::vector<double*> constant_blocks;
  std::vector<std::pair<double*, std::unique_ptr<LocalParameterization>>> parameterizations;
  std::vector<std::unique_ptr<const LossFunction>> owned_losses;

  template <typename... Ts>
  void AddResidualBlock(CostFunction* cost, LossFunction* loss, Ts*... blocks) {
    ResidualBlock rb;
    rb.cost.reset(cost);
    rb.loss = loss;
    rb.blocks = {blocks...};
    bool owned = false;
    for (const auto& l : owned_losses) owned = owned || l.get() == loss;
    if (!owned && loss != nullptr) owned_losses.emplace_back(loss);  // Ceres takes ownership
    residual_blocks.push_back(std::move(rb));
  }
  void SetParameterBlockConstant(double* block) { constant_blocks.push_back(block); }
  void SetParameterization(double* block, LocalParameterization* p) {
    parameterizations.emplace_back(block, std::unique_ptr<LocalParameterization>(p));
  }
  int NumResidualBlocks() const { return static_cast<int>(residual_blocks.size()); }
  int NumResiduals() const {
    int n = 0;
    for (const auto& rb : residual_blocks) n += rb.cost->NumResiduals();
    return n;
  }
};

class Solver {
 public:
  struct Options {
    double function_tolerance = 1e-6;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    bool minimizer_progress_to_stdout = false;
    int max_num_iterations = 50;
    int max_linear_solver_iterations = 500;
    int max_num_consecutive_invalid_steps = 5;
    int max_consecutive_nonmonotonic_steps = 5;
    int num_threads = 1;
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    PreconditionerType preconditioner_type = JACOBI;
    bool IsValid(std::string*) const { return true; }
  };
  struct Summary {
    int num_residuals_reduced = 0;
    int num_effective_parameters_reduced = 0;
    int num_successful_steps = 0;
    int num_unsuccessful_steps = 0;
    double total_time_in_seconds = 0.0;
    double initial_cost = 0.0;
    double final_cost = 0.0;
    TerminationType termination_type = NO_CONVERGENCE;
    std::string BriefReport() const { return "recorder: nothing solved"; }
    std::string FullReport() const { return BriefReport(); }
    bool IsSolutionUsable() const { return false; }
  };
};

// what the last ceres::Solve call was handed (one per thread): the options, and a structural
// snapshot of the problem (some callers build it on their stack: estimators/pose.cc)
struct SolveRecord {
  Solver::Options options;
  int num_residual_blocks = 0;
  int num_residuals = 0;
  std::vector<std::vector<int>> block_sizes;        // per residual block
  std::vector<std::vector<const double*>> blocks;   // per residual block: parameter pointers
  std::vector<std::vector<double>> block3_values;   // per residual block: values of its 3-vectors
  std::vector<int> loss_kind;
  std::vector<double> loss_scale;
  std::vector<const double*> constant_blocks;
  struct Param {
    const double* block;
    int global_size, local_size;
    std::vector<int> constant;
  };
  std::vector<Param> parameterizations;
};
inline SolveRecord& LastSolveRecord() {
  static thread_local SolveRecord r;
  return r;
}
inline Solver::Options& LastSolveOptions() { return LastSolveRecord().options; }
inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  SolveRecord& r = LastSolveRecord();
  r = SolveRecord();
  r.options = options;
  r.num_residual_blocks = problem->NumResidualBlocks();
  r.num_residuals = problem->NumResiduals();
  for (const auto& rb : problem->residual_blocks) {
    r.block_sizes.push_back(rb.cost->ParameterBlockSizes());
    r.blocks.emplace_back(rb.blocks.begin(), rb.blocks.end());
    std::vector<double> v3;
    for (size_t b = 0; b < rb.blocks.size(); ++b)
      if (rb.cost->ParameterBlockSizes()[b] == 3)
        for (int k = 0; k < 3; ++k) v3.push_back(rb.blocks[b][k]);
    r.block3_values.push_back(v3);
    r.loss_kind.push_back(rb.loss ? rb.loss->Kind() : -1);
    r.loss_scale.push_back(rb.loss ? rb.loss->Scale() : 0.0);
  }
  r.constant_blocks.assign(problem->constant_blocks.begin(), problem->constant_blocks.end());
  for (const auto& p : problem->parameterizations) {
    SolveRecord::Param q;
    q.block = p.first;
    q.global_size = p.second->GlobalSize();
    q.local_size = p.second->LocalSize();
    if (p.second->ConstantIndices()) q.constant = *p.second->ConstantIndices();
    r.parameterizations.push_back(q);
  }
  summary->num_residuals_reduced = problem->NumResiduals();
}

}  // namespace ceres
