// TEST INFRASTRUCTURE: a minimal stand-in for the TensorFlow-0.12 C++ op API and for the Eigen matrix subset that
// /root/reference/hungarian.cc uses, so that THE REFERENCE'S OWN hungarian.cc compiles unmodified with plain g++
// (oracle/Makefile -> oracle/_ref/libhungarian_ref.so) and can be called as the checker of oracle/hungarian_ref.c.
// Nothing here implements any part of the algorithm: containers, accessors, logging and registration macros only.
#pragma once
#include <sys/param.h>  // MIN / MAX: hungarian.cc uses MIN without defining it (it reaches it through TF's includes)

#include <algorithm>
#include <cstring>
#include <deque>
#include <iostream>
#include <limits>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------------------- Eigen
namespace Eigen {
enum { ColMajor = 0, RowMajor = 1 };

template <typename T, int R, int C, int O>
class Matrix;

template <typename M>
class Block {  // a view of rows [i, i+r) x cols [j, j+c)
 public:
  Block(M *m, int i, int j, int r, int c) : m_(m), i_(i), j_(j), r_(r), c_(c) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  float operator()(int a, int b) const { return (*m_)(i_ + a, j_ + b); }
  template <typename Other>
  Block &operator=(const Other &o) {
    for (int a = 0; a < r_; ++a)
      for (int b = 0; b < c_; ++b) (*m_)(i_ + a, j_ + b) = o(a, b);
    return *this;
  }
  Block &operator=(const Block &o) {
    for (int a = 0; a < r_; ++a)
      for (int b = 0; b < c_; ++b) (*m_)(i_ + a, j_ + b) = o(a, b);
    return *this;
  }

 private:
  M *m_;
  int i_, j_, r_, c_;
};

template <typename M>
class Map {  // row-major view of external storage
 public:
  Map(float *p, long long r, long long c) : p_(p), r_((int)r), c_((int)c) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  float operator()(int a, int b) const { return p_[(size_t)a * c_ + b]; }

 private:
  float *p_;
  int r_, c_;
};

template <typename T, int R, int C, int O>
class Matrix {
  static_assert(O == RowMajor, "the reference uses a row-major dynamic float matrix");

 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(int r, int c) : r_(r), c_(c), d_((size_t)r * c) {}
  template <typename Other>
  Matrix(const Other &o) : r_(o.rows()), c_(o.cols()), d_((size_t)o.rows() * o.cols()) {
    for (int a = 0; a < r_; ++a)
      for (int b = 0; b < c_; ++b) (*this)(a, b) = o(a, b);
  }
  static Matrix Zero(int r, int c) { return Constant(r, c, T(0)); }
  static Matrix Constant(int r, int c, T v) {
    Matrix m(r, c);
    std::fill(m.d_.begin(), m.d_.end(), v);
    return m;
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int outerSize() const { return r_; }  // row-major: outer = rows, inner = cols
  int innerSize() const { return c_; }
  T &operator()(int a, int b) { return d_[(size_t)a * c_ + b]; }
  const T &operator()(int a, int b) const { return d_[(size_t)a * c_ + b]; }
  T maxCoeff() const { return *std::max_element(d_.begin(), d_.end()); }
  Block<Matrix> block(int i, int j, int r, int c) { return Block<Matrix>(this, i, j, r, c); }
  Block<const Matrix> block(int i, int j, int r, int c) const { return Block<const Matrix>(this, i, j, r, c); }

  class RowwiseOp {
   public:
    explicit RowwiseOp(const Matrix *m) : m_(m) {}
    Matrix maxCoeff() const {
      Matrix out(m_->rows(), 1);
      for (int a = 0; a < m_->rows(); ++a) {
        T v = (*m_)(a, 0);
        for (int b = 1; b < m_->cols(); ++b) v = std::max(v, (*m_)(a, b));
        out(a, 0) = v;
      }
      return out;
    }

   private:
    const Matrix *m_;
  };
  RowwiseOp rowwise() const { return RowwiseOp(this); }

 private:
  int r_, c_;
  std::vector<T> d_;
};

template <typename T, int R, int C, int O>
std::ostream &operator<<(std::ostream &os, const Matrix<T, R, C, O> &m) {
  for (int a = 0; a < m.rows(); ++a) {
    for (int b = 0; b < m.cols(); ++b) os << m(a, b) << ' ';
    os << '\n';
  }
  return os;
}
template <typename M>
std::ostream &operator<<(std::ostream &os, const Block<M> &m) {
  for (int a = 0; a < m.rows(); ++a) {
    for (int b = 0; b < m.cols(); ++b) os << m(a, b) << ' ';
    os << '\n';
  }
  return os;
}
}  // namespace Eigen

// ----------------------------------------------------------------------------------------------------- tensorflow
namespace tensorflow {

struct FatalError : public std::runtime_error {  // LOG(FATAL) aborts the process in TensorFlow; here it throws
  explicit FatalError(const std::string &m) : std::runtime_error(m) {}
};

class LogSink {
 public:
  explicit LogSink(bool fatal) : fatal_(fatal) {}
  ~LogSink() noexcept(false) {
    if (fatal_) throw FatalError(ss_.str());
  }
  template <typename T>
  LogSink &operator<<(const T &v) {
    if (fatal_) ss_ << v;
    return *this;
  }
  LogSink &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }

 private:
  bool fatal_;
  std::ostringstream ss_;
};
enum { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };

class Status {
 public:
  bool ok() const { return true; }
};

class TensorShape {
 public:
  int dims() const { return (int)d_.size(); }
  long long dim_size(int i) const { return d_[i]; }
  void AddDim(long long n) { d_.push_back(n); }
  long long num_elements() const {
    long long n = 1;
    for (size_t i = 0; i < d_.size(); ++i) n *= d_[i];
    return n;
  }

 private:
  std::vector<long long> d_;
};

struct StringPieceStub {
  const char *p;
  const char *data() const { return p; }
};

template <int N>
class Accessor {  // row-major element accessor of rank N (2 or 3)
 public:
  Accessor(float *p, const TensorShape &s) : p_(p), s_(s) {}
  float &operator()(long long i, long long j) const { return p_[i * s_.dim_size(1) + j]; }
  float &operator()(long long i, long long j, long long k) const {
    return p_[(i * s_.dim_size(1) + j) * s_.dim_size(2) + k];
  }

 private:
  float *p_;
  TensorShape s_;
};

class Tensor {
 public:
  Tensor() {}
  explicit Tensor(const TensorShape &s) : shape_(s), d_((size_t)s.num_elements()) {}
  const TensorShape &shape() const { return shape_; }
  StringPieceStub tensor_data() const { return StringPieceStub{reinterpret_cast<const char *>(d_.data())}; }
  template <typename T>
  Accessor<2> matrix() { return Accessor<2>(d_.data(), shape_); }
  template <typename T, int N>
  Accessor<N> tensor() { return Accessor<N>(d_.data(), shape_); }
  template <typename T, int N>
  Accessor<N> tensor() const { return Accessor<N>(const_cast<float *>(d_.data()), shape_); }
  float *raw() { return d_.data(); }
  const float *raw() const { return d_.data(); }

 private:
  TensorShape shape_;
  std::vector<float> d_;
};

class OpKernelConstruction {};

class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  std::deque<Tensor> outputs;  // a deque: growing it keeps the Tensor* handed out by allocate_output valid
  const Tensor &input(int i) const { return inputs[i]; }
  Status allocate_output(int idx, const TensorShape &shape, Tensor **out) {
    if ((int)outputs.size() <= idx) outputs.resize(idx + 1);
    outputs[idx] = Tensor(shape);
    *out = &outputs[idx];
    return Status();
  }
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction *) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext *context) = 0;
};

struct OpDefBuilderStub {
  explicit OpDefBuilderStub(const char *) {}
  OpDefBuilderStub &Input(const char *) { return *this; }
  OpDefBuilderStub &Output(const char *) { return *this; }
};

}  // namespace tensorflow

#define LOG(severity) ::tensorflow::LogSink(::tensorflow::severity == ::tensorflow::FATAL)
#define VLOG(level) \
  if (true) {       \
  } else            \
    ::tensorflow::LogSink(false)
#define OP_REQUIRES_OK(ctx, expr)            \
  do {                                       \
    ::tensorflow::Status _s = (expr);        \
    if (!_s.ok()) return;                    \
  } while (0)
#define RA_CAT2(a, b) a##b
#define RA_CAT(a, b) RA_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilderStub RA_CAT(ra_ref_op_, __LINE__) = ::tensorflow::OpDefBuilderStub(name)
// the registration becomes a factory the driver can call
#define REGISTER_KERNEL_BUILDER(builder, cls)                   \
  ::tensorflow::OpKernel *ra_ref_make_kernel() {                \
    static ::tensorflow::OpKernelConstruction construction;    \
    return new cls(&construction);                              \
  }
