// minieigen.h — TEST INFRASTRUCTURE ONLY.  A stand-in for the handful of Eigen 3 calls the
// reference's P6L / re3q3 / RANSAC / cost-functor sources make, so that those sources compile
// HERE from where they lie under /root/reference (oracle/build_ref.sh -> oracle/_ref/), with no
// Eigen in the image.  It is NOT Eigen: fixed-size, eager (every expression is evaluated to a
// value at once), no vectorisation, and only the members those files use.
//
// Arithmetic contract (what makes the comparison with the restatement meaningful):
//   * element-wise operators do one IEEE operation per element, products sum left to right
//     (k = 0, 1, 2, ...) — what Eigen's unrolled, non-vectorised small fixed-size kernels do;
//   * determinant(), lu()/partialPivLu().solve(), EigenSolver: the restatements of
//     oracle/eigen_restated.h (the SAME functions the oracle calls) — the Eigen boundary stays
//     unpinned, everything else in the compiled reference sources is the reference's own text;
//   * setRandom() / Quaternion::UnitRandom() return FIXED generic values (the ones the oracle
//     uses) instead of C rand(): the two degenerate fallbacks stay reproducible.
#pragma once
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <type_traits>
#include <vector>

#include "../../eigen_restated.h"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

constexpr int Dynamic = -1;
enum StorageOptions { ColMajor = 0, RowMajor = 1 };
enum TransformTraits { Affine = 2 };

template <class T>
struct aligned_allocator {
  using value_type = T;
  aligned_allocator() = default;
  template <class U>
  aligned_allocator(const aligned_allocator<U>&) {}
  T* allocate(std::size_t n) { return static_cast<T*>(::operator new(n * sizeof(T))); }
  void deallocate(T* p, std::size_t) { ::operator delete(p); }
  template <class U>
  struct rebind {
    using other = aligned_allocator<U>;
  };
  bool operator==(const aligned_allocator&) const { return true; }
  bool operator!=(const aligned_allocator&) const { return false; }
};

template <class S, int R, int C, int Opt = ColMajor>
class Matrix;
template <class Xpr, int BR, int BC>
class Block;
template <class M>
class Map;

template <class D>
struct traits;
template <class S, int R, int C, int Opt>
struct traits<Matrix<S, R, C, Opt>> {
  using Scalar = S;
  static constexpr int Rows = R, Cols = C;
};
template <class Xpr, int BR, int BC>
struct traits<Block<Xpr, BR, BC>> {
  using Scalar = typename traits<std::remove_const_t<Xpr>>::Scalar;
  static constexpr int Rows = BR, Cols = BC;
};
template <class M>
struct traits<Map<M>> {
  using Scalar = typename traits<std::remove_const_t<M>>::Scalar;
  static constexpr int Rows = traits<std::remove_const_t<M>>::Rows;
  static constexpr int Cols = traits<std::remove_const_t<M>>::Cols;
};

template <class Target>
class CommaInitializer;
template <class S>
class PartialPivLU3;

// ---------------------------------------------------------------------------------------------
// Everything dense derives from this: Derived supplies coeff(i, j) (and coeffRef for writes).
// ---------------------------------------------------------------------------------------------
template <class D>
class MatrixBase {
 public:
  using Scalar = typename traits<D>::Scalar;
  static constexpr int Rows = traits<D>::Rows, Cols = traits<D>::Cols;
  using Plain = Matrix<Scalar, Rows, Cols>;

  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }

  Scalar operator()(int i, int j) const { return derived().coeff(i, j); }
  Scalar& operator()(int i, int j) { return derived().coeffRef(i, j); }
  // vectors (either orientation)
  Scalar operator()(int i) const { return Cols == 1 ? derived().coeff(i, 0) : derived().coeff(0, i); }
  Scalar& operator()(int i) { return Cols == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
  Scalar operator[](int i) const { return (*this)(i); }
  Scalar& operator[](int i) { return (*this)(i); }
  Scalar x() const { return (*this)(0); }
  Scalar y() const { return (*this)(1); }
  Scalar z() const { return (*this)(2); }

  Plain eval() const {
    Plain r;
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) r.coeffRef(i, j) = derived().coeff(i, j);
    return r;
  }
  Matrix<Scalar, Cols, Rows> transpose() const {
    Matrix<Scalar, Cols, Rows> r;
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) r.coeffRef(j, i) = derived().coeff(i, j);
    return r;
  }

  template <class O>
  D& assign(const MatrixBase<O>& o) {
    constexpr bool same = O::Rows == Rows && O::Cols == Cols;
    // Eigen transposes implicitly when a row vector is assigned to a column vector (and back)
    constexpr bool vec_t = O::Rows == Cols && O::Cols == Rows && (Rows == 1 || Cols == 1);
    static_assert(same || vec_t, "minieigen: size mismatch");
    const typename O::Plain v = o.eval();  // (aliasing-safe)
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) = same ? v.coeff(i, j) : v.coeff(j, i);
    return derived();
  }
  template <class O>
  D& operator+=(const MatrixBase<O>& o) {
    const typename O::Plain v = o.eval();
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) = derived().coeff(i, j) + v.coeff(i, j);
    return derived();
  }
  template <class O>
  D& operator-=(const MatrixBase<O>& o) {
    const typename O::Plain v = o.eval();
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) = derived().coeff(i, j) - v.coeff(i, j);
    return derived();
  }
  D& operator/=(Scalar s) {
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) = derived().coeff(i, j) / s;
    return derived();
  }
  D& operator*=(Scalar s) {
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) derived().coeffRef(i, j) = derived().coeff(i, j) * s;
    return derived();
  }
  template <class O>
  void swap(MatrixBase<O>&& o) { swap(o); }
  template <class O>
  void swap(MatrixBase<O>& o) {
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) {
        const Scalar t = derived().coeff(i, j);
        derived().coeffRef(i, j) = o.derived().coeff(i, j);
        o.derived().coeffRef(i, j) = t;
      }
  }

  // Eigen's "random" fills of the two degenerate fallbacks, fixed (see the header comment).
  D& setRandom() {
    static_assert(std::is_same<Scalar, double>::value && Rows == 3 && (Cols == 3 || Cols == 1),
                  "minieigen: setRandom() only for the reference's 3x3 / 3x1 uses");
    if (Cols == 3) {  // absolute_pose.cc:131 — the oracle's kMixA
      static const double k[3][3] = {{0.680375, -0.211234, 0.566198},
                                     {0.596880, 0.823295, -0.604897},
                                     {-0.329554, 0.536459, -0.444451}};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) derived().coeffRef(i, j) = k[i][j];
    } else {  // re3q3.h:42 — normalised by the caller
      static const double k[3] = {0.3, -0.5, 0.8};
      for (int i = 0; i < 3; ++i) derived().coeffRef(i, 0) = k[i];
    }
    return derived();
  }
  Scalar squaredNorm() const {
    Scalar s = derived().coeff(0, 0) * derived().coeff(0, 0);
    for (int j = 0; j < Cols; ++j)
      for (int i = (j == 0 ? 1 : 0); i < Rows; ++i) s = s + derived().coeff(i, j) * derived().coeff(i, j);
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Scalar sum() const {
    Scalar s = derived().coeff(0, 0);
    for (int j = 0; j < Cols; ++j)
      for (int i = (j == 0 ? 1 : 0); i < Rows; ++i) s = s + derived().coeff(i, j);
    return s;
  }
  Plain normalized() const { return eval() / norm(); }
  bool hasNaN() const {
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i)
        if (derived().coeff(i, j) != derived().coeff(i, j)) return true;
    return false;
  }
  void normalize() { *this /= norm(); }
  template <class O>
  Scalar dot(const MatrixBase<O>& o) const {  // vectors of either orientation
    static_assert((Cols == 1 || Rows == 1) && (O::Cols == 1 || O::Rows == 1) &&
                      Rows * Cols == O::Rows * O::Cols,
                  "minieigen: dot of two vectors of one length");
    Scalar s = (*this)(0) * o(0);
    for (int i = 1; i < Rows * Cols; ++i) s = s + (*this)(i) * o(i);
    return s;
  }
  // (x, 1) and x / x_last for column vectors
  Matrix<Scalar, Rows + 1, 1> homogeneous() const {
    static_assert(Cols == 1, "minieigen: homogeneous() of a column vector");
    Matrix<Scalar, Rows + 1, 1> r;
    for (int i = 0; i < Rows; ++i) r.coeffRef(i, 0) = derived().coeff(i, 0);
    r.coeffRef(Rows, 0) = Scalar(1);
    return r;
  }
  Matrix<Scalar, Rows - 1, 1> hnormalized() const {
    static_assert(Cols == 1 && Rows > 1, "minieigen: hnormalized() of a column vector");
    Matrix<Scalar, Rows - 1, 1> r;
    for (int i = 0; i < Rows - 1; ++i) r.coeffRef(i, 0) = derived().coeff(i, 0) / derived().coeff(Rows - 1, 0);
    return r;
  }
  int rows() const { return Rows; }
  int cols() const { return Cols; }
  template <class T>
  Matrix<T, Rows, Cols> cast() const {
    Matrix<T, Rows, Cols> r;
    for (int j = 0; j < Cols; ++j)
      for (int i = 0; i < Rows; ++i) r.coeffRef(i, j) = static_cast<T>(derived().coeff(i, j));
    return r;
  }
  // Members that only OFF-PATH functions of the compiled reference files name
  // (DecomposeProjectionMatrix, ComputeClosestRotationMatrix, DecomposeMatrixRQ): they exist so
  // that those translation units compile, and abort if anything ever runs them.
  struct OffPathReverse {
    Plain reverse() const {
      std::abort();
      return Plain();
    }
  };
  OffPathReverse rowwise() const { return OffPathReverse(); }
  OffPathReverse colwise() const { return OffPathReverse(); }
  struct OffPathTriangular {
    template <class B>
    typename B::Plain solve(const MatrixBase<B>&) const {
      std::abort();
      return typename B::Plain();
    }
  };
  template <int Mode>
  OffPathTriangular triangularView() const { return OffPathTriangular(); }

  // Matrix3d::determinant()
  Scalar determinant() const {
    static_assert(std::is_same<Scalar, double>::value && Rows == 3 && Cols == 3,
                  "minieigen: determinant() is 3x3 only");
    double m[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) m[i][j] = derived().coeff(i, j);
    return eigen_restated::Det3(m);
  }
  // Matrix2d::inverse() (adjugate times 1 / det, Eigen's size-2 kernel).  Only
  // BaseCameraModel::IterativeUndistortion names it; nothing on the tested path evaluates it.
  Plain inverse() const {
    static_assert(Rows == 2 && Cols == 2, "minieigen: inverse() is 2x2 only");
    const D& m = derived();
    const Scalar invdet = Scalar(1) / (m.coeff(0, 0) * m.coeff(1, 1) - m.coeff(1, 0) * m.coeff(0, 1));
    Plain r;
    r.coeffRef(0, 0) = m.coeff(1, 1) * invdet;
    r.coeffRef(1, 0) = -m.coeff(1, 0) * invdet;
    r.coeffRef(0, 1) = -m.coeff(0, 1) * invdet;
    r.coeffRef(1, 1) = m.coeff(0, 0) * invdet;
    return r;
  }
  PartialPivLU3<Scalar> partialPivLu() const { return PartialPivLU3<Scalar>(eval()); }
  PartialPivLU3<Scalar> lu() const { return partialPivLu(); }  // Eigen 3: lu() == partialPivLu()

  template <int BR, int BC>
  Block<D, BR, BC> block(int i, int j) { return Block<D, BR, BC>(derived(), i, j); }
  template <int BR, int BC>
  Block<const D, BR, BC> block(int i, int j) const { return Block<const D, BR, BC>(derived(), i, j); }
  template <int N>
  Block<D, Rows, N> leftCols() { return Block<D, Rows, N>(derived(), 0, 0); }
  template <int N>
  Block<const D, Rows, N> leftCols() const { return Block<const D, Rows, N>(derived(), 0, 0); }
  template <int N>
  Block<D, Rows, N> rightCols() { return Block<D, Rows, N>(derived(), 0, Cols - N); }
  template <int N>
  Block<const D, Rows, N> rightCols() const { return Block<const D, Rows, N>(derived(), 0, Cols - N); }
  // x.array() == y.array() ... .all(): only what util/matrix.h's IsNaN / IsInf write
  struct BoolArray {
    bool all_;
    bool all() const { return all_; }
  };
  struct ArrayView {
    Plain v;
    BoolArray operator==(const ArrayView& o) const {
      bool a = true;
      for (int j = 0; j < Cols; ++j)
        for (int i = 0; i < Rows; ++i) a = a && (v.coeff(i, j) == o.v.coeff(i, j));
      return BoolArray{a};
    }
  };
  ArrayView array() const { return ArrayView{eval()}; }
  // head<N>() of a column vector
  template <int N>
  Block<D, N, 1> head() { return Block<D, N, 1>(derived(), 0, 0); }
  template <int N>
  Block<const D, N, 1> head() const { return Block<const D, N, 1>(derived(), 0, 0); }
  Block<D, Rows, 1> col(int j) { return Block<D, Rows, 1>(derived(), 0, j); }
  Block<const D, Rows, 1> col(int j) const { return Block<const D, Rows, 1>(derived(), 0, j); }
  Block<D, 1, Cols> row(int i) { return Block<D, 1, Cols>(derived(), i, 0); }
  Block<const D, 1, Cols> row(int i) const { return Block<const D, 1, Cols>(derived(), i, 0); }
};

// ---------------------------------------------------------------------------------------------
template <class S, int R, int C, int Opt>
class Matrix : public MatrixBase<Matrix<S, R, C, Opt>> {
  static_assert(R > 0 && C > 0, "minieigen: fixed sizes only");
  S d_[R * C];  // column-major like Eigen's default

 public:
  using Base = MatrixBase<Matrix>;
  Matrix() {}
  template <class O>
  Matrix(const MatrixBase<O>& o) { Base::assign(o); }
  Matrix(const S& x, const S& y) {
    static_assert(R * C == 2, "minieigen: 2-vector constructor");
    d_[0] = x; d_[1] = y;
  }
  Matrix(const S& x, const S& y, const S& z) {
    static_assert(R * C == 3, "minieigen: 3-vector constructor");
    d_[0] = x; d_[1] = y; d_[2] = z;
  }
  Matrix(const S& x, const S& y, const S& z, const S& w) {
    static_assert(R * C == 4, "minieigen: 4-vector constructor");
    d_[0] = x; d_[1] = y; d_[2] = z; d_[3] = w;
  }
  template <class O>
  Matrix& operator=(const MatrixBase<O>& o) { return Base::assign(o); }

  static Matrix Zero() {
    Matrix m;
    for (int i = 0; i < R * C; ++i) m.d_[i] = S(0);
    return m;
  }
  static Matrix Identity() {
    Matrix m = Zero();
    for (int i = 0; i < (R < C ? R : C); ++i) m.coeffRef(i, i) = S(1);
    return m;
  }

  // a 1 x 1 product is a scalar (std::acos(a.transpose() * b))
  template <int RR = R, int CC = C, std::enable_if_t<RR == 1 && CC == 1, int> = 0>
  operator S() const { return d_[0]; }

  const S& coeff(int i, int j) const { return d_[j * R + i]; }
  S& coeffRef(int i, int j) { return d_[j * R + i]; }
  const S* data() const { return d_; }
  S* data() { return d_; }

  // comma initialisation: `M << a, b, c, ...;`
  template <class T>
  CommaInitializer<Matrix> operator<<(const T& first) {
    CommaInitializer<Matrix> ci(*this);
    ci, first;
    return ci;
  }
};

template <class Xpr, int BR, int BC>
class Block : public MatrixBase<Block<Xpr, BR, BC>> {
  Xpr& x_;
  int i0_, j0_;

 public:
  using Base = MatrixBase<Block>;
  using S = typename Base::Scalar;
  Block(Xpr& x, int i0, int j0) : x_(x), i0_(i0), j0_(j0) {}
  Block(const Block&) = default;
  S coeff(int i, int j) const { return x_.coeff(i0_ + i, j0_ + j); }
  S& coeffRef(int i, int j) { return x_.coeffRef(i0_ + i, j0_ + j); }
  template <class O>
  Block& operator=(const MatrixBase<O>& o) { return Base::assign(o); }
  Block& operator=(const Block& o) { return Base::assign(o); }
};

template <class M>
class Map : public MatrixBase<Map<M>> {
  using Base = MatrixBase<Map>;
  using S = typename Base::Scalar;
  using Ptr = std::conditional_t<std::is_const<M>::value, const S*, S*>;
  Ptr p_;

 public:
  explicit Map(Ptr p) : p_(p) {}
  S coeff(int i, int j) const { return p_[j * Base::Rows + i]; }
  S& coeffRef(int i, int j) { return const_cast<S*>(p_)[j * Base::Rows + i]; }
  template <class O>
  Map& operator=(const MatrixBase<O>& o) { return Base::assign(o); }
};

// `M << ...`: scalars and blocks, filled left to right, top to bottom (Eigen's CommaInitializer)
template <class Target>
class CommaInitializer {
  Target& t_;
  int row_ = 0, col_ = 0, h_ = 1;
  void Advance(int h, int w) {
    h_ = h;
    col_ += w;
    if (col_ >= Target::Cols) {
      col_ = 0;
      row_ += h_;
    }
  }

 public:
  explicit CommaInitializer(Target& t) : t_(t) {}
  template <class T, std::enable_if_t<std::is_arithmetic<T>::value, int> = 0>
  CommaInitializer& operator,(const T& s) {
    t_.coeffRef(row_, col_) = static_cast<typename Target::Scalar>(s);
    Advance(1, 1);
    return *this;
  }
  template <class O>
  CommaInitializer& operator,(const MatrixBase<O>& b) {
    for (int j = 0; j < O::Cols; ++j)
      for (int i = 0; i < O::Rows; ++i) t_.coeffRef(row_ + i, col_ + j) = b.derived().coeff(i, j);
    Advance(O::Rows, O::Cols);
    return *this;
  }
};

// ---------------------------------------------------------------------------------------------
// operators (eager)
// ---------------------------------------------------------------------------------------------
template <class A, class B>
Matrix<typename A::Scalar, A::Rows, B::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  static_assert(A::Cols == B::Rows, "minieigen: inner dimensions");
  Matrix<typename A::Scalar, A::Rows, B::Cols> r;
  for (int j = 0; j < B::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) {
      typename A::Scalar s = a.derived().coeff(i, 0) * b.derived().coeff(0, j);
      for (int k = 1; k < A::Cols; ++k) s = s + a.derived().coeff(i, k) * b.derived().coeff(k, j);
      r.coeffRef(i, j) = s;
    }
  return r;
}
template <class A>
typename A::Plain operator*(const MatrixBase<A>& a, typename A::Scalar s) {
  typename A::Plain r;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) r.coeffRef(i, j) = a.derived().coeff(i, j) * s;
  return r;
}
template <class A>
typename A::Plain operator*(typename A::Scalar s, const MatrixBase<A>& a) {
  typename A::Plain r;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) r.coeffRef(i, j) = s * a.derived().coeff(i, j);
  return r;
}
template <class A>
typename A::Plain operator/(const MatrixBase<A>& a, typename A::Scalar s) {
  typename A::Plain r;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) r.coeffRef(i, j) = a.derived().coeff(i, j) / s;
  return r;
}
template <class A, class B>
typename A::Plain operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  static_assert(A::Rows == B::Rows && A::Cols == B::Cols, "minieigen: size mismatch");
  typename A::Plain r;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) r.coeffRef(i, j) = a.derived().coeff(i, j) + b.derived().coeff(i, j);
  return r;
}
template <class A, class B>
typename A::Plain operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  static_assert(A::Rows == B::Rows && A::Cols == B::Cols, "minieigen: size mismatch");
  typename A::Plain r;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) r.coeffRef(i, j) = a.derived().coeff(i, j) - b.derived().coeff(i, j);
  return r;
}
template <class A>
typename A::Plain operator-(const MatrixBase<A>& a) {
  typename A::Plain r;
  for (int j = 0; j < A::Cols; ++j)
    for (int i = 0; i < A::Rows; ++i) r.coeffRef(i, j) = -a.derived().coeff(i, j);
  return r;
}

// ---------------------------------------------------------------------------------------------
// PartialPivLU<Matrix3d>::solve  (eigen_restated::SolvePartialPiv3)
// ---------------------------------------------------------------------------------------------
template <class S>
class PartialPivLU3 {
  Matrix<S, 3, 3> a_;

 public:
  explicit PartialPivLU3(const Matrix<S, 3, 3>& a) : a_(a) {}
  template <class B>
  typename B::Plain solve(const MatrixBase<B>& rhs) const {
    static_assert(std::is_same<S, double>::value && B::Rows == 3, "minieigen: 3x3 double systems");
    double A[3][3], X[3][B::Cols];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) A[i][j] = a_.coeff(i, j);
      for (int j = 0; j < B::Cols; ++j) X[i][j] = rhs.derived().coeff(i, j);
    }
    eigen_restated::SolvePartialPiv3<B::Cols>(A, X);
    typename B::Plain r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < B::Cols; ++j) r.coeffRef(i, j) = X[i][j];
    return r;
  }
};

// ---------------------------------------------------------------------------------------------
// EigenSolver<Matrix<double,8,8>>: eigenvalues only, of an upper-Hessenberg input (the companion
// matrix of re3q3.h:152-165) — eigen_restated::Hqr8.
// ---------------------------------------------------------------------------------------------
template <class M>
class EigenSolver {
  Matrix<std::complex<double>, 8, 1> ev_;

 public:
  EigenSolver(const M& m, bool compute_eigenvectors = true) {
    static_assert(M::Rows == 8 && M::Cols == 8, "minieigen: EigenSolver is 8x8 only");
    if (compute_eigenvectors) std::abort();  // not provided
    eigen_restated::Hqr8 h;
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) {
        h.T[i][j] = m.coeff(i, j);
        if (i > j + 1 && m.coeff(i, j) != 0.0) std::abort();  // Hessenberg inputs only
      }
    h.Reduce();
    double re[8], im[8];
    h.Eigenvalues(re, im);
    for (int i = 0; i < 8; ++i) ev_.coeffRef(i, 0) = std::complex<double>(re[i], im[i]);
  }
  const Matrix<std::complex<double>, 8, 1>& eigenvalues() const { return ev_; }
};

enum { Lower = 1, Upper = 2 };
enum { ComputeFullU = 4, ComputeThinU = 8, ComputeFullV = 16, ComputeThinV = 32 };

// off-path (see MatrixBase): util/matrix.h DecomposeMatrixRQ, projection.cc
template <class M>
class HouseholderQR {
 public:
  explicit HouseholderQR(const M&) { std::abort(); }
  M householderQ() const { return M(); }
  M matrixQR() const { return M(); }
};
template <class M>
class JacobiSVD {
 public:
  JacobiSVD(const M&, unsigned = 0) { std::abort(); }
  M matrixU() const { return M(); }
  M matrixV() const { return M(); }
};

// The ONE dynamically sized matrix of the compiled sources: the n x 4 system of
// TriangulateMultiViewPoint (src/base/triangulation.cc:41-57) — `A(n, 4)`, `A.row(i) = ...`,
// `JacobiSVD<decltype(A)> svd(A, ComputeFullV)`, `svd.matrixV().col(3)`.
template <class S, int Opt>
class Matrix<S, Dynamic, 4, Opt> {
  std::vector<S> d_;  // row-major n x 4
  int rows_;

 public:
  Matrix(std::size_t rows, int cols) : d_(rows * 4), rows_(static_cast<int>(rows)) {
    if (cols != 4) std::abort();
  }
  int rows() const { return rows_; }
  int cols() const { return 4; }
  const std::vector<S>& storage() const { return d_; }
  struct RowRef {
    S* p;
    template <class O>
    RowRef& operator=(const MatrixBase<O>& o) {
      static_assert(O::Rows == 1 && O::Cols == 4, "minieigen: a 1 x 4 row");
      for (int j = 0; j < 4; ++j) p[j] = o.derived().coeff(0, j);
      return *this;
    }
  };
  RowRef row(int i) { return RowRef{d_.data() + 4 * static_cast<std::size_t>(i)}; }
};
// JacobiSVD of that system: only the right singular vector of the smallest singular value is
// provided (column 3 of matrixV(), the one the reference reads), by the one-sided Jacobi
// iteration of eigen_restated::NullVectorNx4 — the function the oracle uses.  Eigen's own
// JacobiSVD (two-sided, QR-preconditioned) gives the same vector up to rounding and sign.
template <class S, int Opt>
class JacobiSVD<Matrix<S, Dynamic, 4, Opt>> {
  Matrix<S, 4, 4> v_;

 public:
  JacobiSVD(const Matrix<S, Dynamic, 4, Opt>& a, unsigned = 0) {
    std::vector<double> w(a.storage().begin(), a.storage().end());
    double x[4];
    eigen_restated::NullVectorNx4(w, a.rows(), x);
    v_ = Matrix<S, 4, 4>::Zero();
    for (int i = 0; i < 4; ++i) v_.coeffRef(i, 3) = x[i];
  }
  const Matrix<S, 4, 4>& matrixV() const { return v_; }
};

// ---------------------------------------------------------------------------------------------
// Geometry stubs: only what the reference's headers name.
// ---------------------------------------------------------------------------------------------
template <class S>
class Quaternion {
  S w_, x_, y_, z_;
  bool fixed_generic_ = false;  // the UnitRandom() stand-in

 public:
  Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
  Quaternion(S w, S x, S y, S z) : w_(w), x_(x), y_(y), z_(z) {}
  // Quaterniond(Matrix3d): eigen_restated::QuaternionFromRotationMatrix
  template <class O>
  explicit Quaternion(const MatrixBase<O>& rot) {
    static_assert(O::Rows == 3 && O::Cols == 3 && std::is_same<S, double>::value,
                  "minieigen: quaternion from a 3x3 double matrix");
    const Matrix<double, 3, 3> r = rot.eval();
    double q[4];
    eigen_restated::QuaternionFromRotationMatrix(r.data(), q);
    w_ = q[0]; x_ = q[1]; y_ = q[2]; z_ = q[3];
  }
  S w() const { return w_; }
  S x() const { return x_; }
  S y() const { return y_; }
  S z() const { return z_; }
  // Eigen's quaternion product and rotation of a vector (Quaternion.h: quat_product,
  // _transformVector: v + w * 2 (u x v) + u x 2 (u x v), u = (x, y, z))
  Quaternion operator*(const Quaternion& b) const {
    return Quaternion(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_,
                      w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                      w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_,
                      w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  template <class O>
  Matrix<S, 3, 1> operator*(const MatrixBase<O>& vec) const {
    static_assert(O::Rows == 3 && O::Cols == 1, "minieigen: quaternion times a 3-vector");
    const S v0 = vec.derived().coeff(0, 0), v1 = vec.derived().coeff(1, 0), v2 = vec.derived().coeff(2, 0);
    S u0 = y_ * v2 - z_ * v1, u1 = z_ * v0 - x_ * v2, u2 = x_ * v1 - y_ * v0;
    u0 = u0 + u0; u1 = u1 + u1; u2 = u2 + u2;
    Matrix<S, 3, 1> r;
    r.coeffRef(0, 0) = v0 + w_ * u0 + (y_ * u2 - z_ * u1);
    r.coeffRef(1, 0) = v1 + w_ * u1 + (z_ * u0 - x_ * u2);
    r.coeffRef(2, 0) = v2 + w_ * u2 + (x_ * u1 - y_ * u0);
    return r;
  }
  Quaternion slerp(const S&, const Quaternion&) const {  // off-path (InterpolatePose)
    std::abort();
    return *this;
  }
  // re3q3.h:41 draws a random rotation; fixed here: the rotation of the oracle's kVarChangeA
  static Quaternion UnitRandom() {
    Quaternion q;
    q.fixed_generic_ = true;
    return q;
  }
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> r;
    if (fixed_generic_) {
      static const double k[3][3] = {
          {-0.45264637943155561, -0.88862107060359552, 0.073917846740986226},
          {0.19122225569950785, -0.015767801546650473, 0.98142010645776834},
          {-0.8709450637742292, 0.45837099527970276, 0.1770613638646179}};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.coeffRef(i, j) = S(k[i][j]);
      return r;
    }
    const S tx = S(2) * x_, ty = S(2) * y_, tz = S(2) * z_;
    const S twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const S txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const S tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    r.coeffRef(0, 0) = S(1) - (tyy + tzz); r.coeffRef(0, 1) = txy - twz; r.coeffRef(0, 2) = txz + twy;
    r.coeffRef(1, 0) = txy + twz; r.coeffRef(1, 1) = S(1) - (txx + tzz); r.coeffRef(1, 2) = tyz - twx;
    r.coeffRef(2, 0) = txz - twy; r.coeffRef(2, 1) = tyz + twx; r.coeffRef(2, 2) = S(1) - (txx + tyy);
    return r;
  }
};
template <class S, int Dim, int Mode>
class Transform {};

// the one dynamic type the compiled headers name (base/polynomial.h, included but unused by
// absolute_pose.cc): size() and element access only
class VectorXd {
  std::vector<double> v_;

 public:
  typedef std::ptrdiff_t Index;
  VectorXd() {}
  explicit VectorXd(Index n) : v_(static_cast<std::size_t>(n)) {}
  Index size() const { return static_cast<Index>(v_.size()); }
  double operator()(Index i) const { return v_[static_cast<std::size_t>(i)]; }
  double& operator()(Index i) { return v_[static_cast<std::size_t>(i)]; }
};

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<float, 3, Affine> Affine3f;

}  // namespace Eigen
