// TEST INFRASTRUCTURE — a stand-in for <RcppArmadillo.h>, written for this repo (not Armadillo code).
//
// Purpose: compile the reference's own gibbs.cpp / sampler.cpp / utils.cpp / main.cpp VERBATIM (from
// /root/reference/src, where they lie) without R, Rcpp or Armadillo, so that the oracle restatement can
// be pinned against the reference's actual control flow. Only the subset of the Armadillo API those
// four files use is provided, with EAGER evaluation (every expression yields a Mat). Element-wise
// operations are the same IEEE operations real Armadillo performs; reductions (sum / accu) are plain
// left-to-right loops — real Armadillo uses a two-accumulator unrolled order for some of them, so
// energies can differ from a true Armadillo build by O(u*eps). This is stated in DESIGN.md.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

#include "Rcpp.h"

namespace arma {

typedef unsigned long long uword;

namespace fill {
struct fill_zeros {};
static const fill_zeros zeros{};
}  // namespace fill

struct SizeMat {
  uword n_rows, n_cols;
};

class Mat;
class Col;
class Row;
class UVec;

// ---------------------------------------------------------------- index vector (arma::uvec)
class UVec {
 public:
  uword n_elem = 0;
  UVec() = default;
  explicit UVec(uword n) : n_elem(n), mem_(n, 0) {}
  UVec(uword n, fill::fill_zeros) : n_elem(n), mem_(n, 0) {}
  UVec(const uword* p, uword n) : n_elem(n), mem_(p, p + n) {}
  uword& operator()(uword i) { return mem_[i]; }
  const uword& operator()(uword i) const { return mem_[i]; }
  uword& operator[](uword i) { return mem_[i]; }
  const uword& operator[](uword i) const { return mem_[i]; }
  UVec head(uword n) const { return UVec(mem_.data(), n); }
  const uword* begin() const { return mem_.data(); }
  const uword* end() const { return mem_.data() + n_elem; }
  const uword* memptr() const { return mem_.data(); }

 private:
  std::vector<uword> mem_;
};
typedef UVec uvec;

// ---------------------------------------------------------------- proxies (declared before Mat)
struct SubCols;   // contiguous block of columns   (tail_cols, col)
struct SubVec;    // contiguous tail of a vector   (tail)
struct Elem1;     // v(uvec)
struct Elem2;     // M.rows(uvec)
struct EachCol;   // M.each_col()
struct EachRow;   // M.each_row()

class Mat {
 public:
  typedef double elem_type;
  uword n_rows = 0, n_cols = 0, n_elem = 0;

  Mat() = default;
  Mat(uword r, uword c) { init(r, c); }
  Mat(uword r, uword c, fill::fill_zeros) { init(r, c); }
  Mat(const double* p, uword r, uword c) { init(r, c); std::copy(p, p + n_elem, mem_.begin()); }

  void init(uword r, uword c) {
    n_rows = r; n_cols = c; n_elem = r * c;
    mem_.assign(n_elem, 0.0);
  }
  double* memptr() { return mem_.data(); }
  const double* memptr() const { return mem_.data(); }
  double* colptr(uword c) { return mem_.data() + c * n_rows; }
  const double* colptr(uword c) const { return mem_.data() + c * n_rows; }

  double& operator()(uword i, uword j) { return mem_[i + j * n_rows]; }
  const double& operator()(uword i, uword j) const { return mem_[i + j * n_rows]; }
  double& operator()(uword i) { return mem_[i]; }
  const double& operator()(uword i) const { return mem_[i]; }
  double& operator[](uword i) { return mem_[i]; }
  const double& operator[](uword i) const { return mem_[i]; }

  inline Elem1 operator()(const UVec& idx);
  inline Elem1 operator()(const UVec& idx) const;
  inline Elem2 rows(const UVec& idx);
  inline Elem2 rows(const UVec& idx) const;
  inline SubCols tail_cols(uword k);
  inline SubCols tail_cols(uword k) const;
  inline SubCols col(uword c);
  inline SubVec tail(uword k);
  inline EachCol each_col();
  inline EachCol each_col() const;
  inline EachRow each_row();
  inline EachRow each_row() const;

  double max() const {
    double m = mem_[0];
    for (uword i = 1; i < n_elem; ++i) m = mem_[i] > m ? mem_[i] : m;
    return m;
  }
  Mat t() const {
    Mat o(n_cols, n_rows);
    for (uword j = 0; j < n_cols; ++j)
      for (uword i = 0; i < n_rows; ++i) o(j, i) = (*this)(i, j);
    return o;
  }
  inline Col diag() const;
  void insert_cols(uword at, uword count) {
    Mat o(n_rows, n_cols + count);
    for (uword j = 0; j < n_cols; ++j) {
      uword dj = j < at ? j : j + count;
      std::copy(colptr(j), colptr(j) + n_rows, o.colptr(dj));
    }
    *this = o;
  }
  template <class F>
  void for_each(F f) {
    for (uword i = 0; i < n_elem; ++i) f(mem_[i]);
  }
  Mat& operator+=(const Mat& b) {
    for (uword i = 0; i < n_elem; ++i) mem_[i] += b[i];
    return *this;
  }
  Mat& operator-=(const Mat& b) {
    for (uword i = 0; i < n_elem; ++i) mem_[i] -= b[i];
    return *this;
  }

 protected:
  std::vector<double> mem_;
};
typedef Mat mat;

class Col : public Mat {
 public:
  Col() = default;
  explicit Col(uword n) : Mat(n, 1) {}
  Col(uword n, fill::fill_zeros) : Mat(n, 1) {}
  Col(const Mat& m) : Mat(m) { n_rows = n_elem; n_cols = 1; }
  Col(const Rcpp::NumericVector& v) : Mat(v.begin(), (uword)v.size(), 1) {}
  Col(const double* p, uword n) : Mat(p, n, 1) {}
  Col(const std::vector<double>& v) : Mat(v.data(), v.size(), 1) {}
};
typedef Col vec;
typedef Col colvec;

class Row : public Mat {
 public:
  Row() = default;
  explicit Row(uword n) : Mat(1, n) {}
  Row(const Mat& m) : Mat(m) { n_cols = n_elem; n_rows = 1; }
  Row(const std::vector<double>& v) : Mat(v.data(), 1, v.size()) {}
};
typedef Row rowvec;

class Cube {
 public:
  uword n_rows = 0, n_cols = 0, n_slices = 0, n_elem = 0;
  Cube() = default;
  Cube(uword r, uword c, uword s, fill::fill_zeros) : n_rows(r), n_cols(c), n_slices(s), n_elem(r * c * s), mem_(r * c * s, 0.0) {}
  Cube(const double* p, uword r, uword c, uword s) : n_rows(r), n_cols(c), n_slices(s), n_elem(r * c * s), mem_(p, p + r * c * s) {}
  struct Slice {
    Cube* q;
    uword s;
    void operator=(const Mat& m) {
      std::copy(m.memptr(), m.memptr() + m.n_elem, q->mem_.data() + s * q->n_rows * q->n_cols);
    }
  };
  Slice slice(uword s) { return Slice{this, s}; }
  const double* memptr() const { return mem_.data(); }

 private:
  std::vector<double> mem_;
};
typedef Cube cube;

// ---------------------------------------------------------------- proxy definitions
struct SubCols {
  Mat* m;
  uword c0, nc;
  operator Mat() const {
    return Mat(m->colptr(c0), m->n_rows, nc);
  }
  void operator=(const Mat& b) { std::copy(b.memptr(), b.memptr() + b.n_elem, m->colptr(c0)); }
  void operator=(const SubCols& b) {
    std::copy(b.m->colptr(b.c0), b.m->colptr(b.c0) + b.m->n_rows * b.nc, m->colptr(c0));
  }
};
struct SubVec {
  Mat* m;
  uword i0, n;
  operator Col() const { return Col(m->memptr() + i0, n); }
  void operator=(const Mat& b) { std::copy(b.memptr(), b.memptr() + b.n_elem, m->memptr() + i0); }
};
struct Elem1 {
  Mat* m;
  const UVec* idx;
  operator Col() const {
    Col o(idx->n_elem);
    for (uword q = 0; q < idx->n_elem; ++q) o[q] = (*m)[(*idx)[q]];
    return o;
  }
  void operator=(const Mat& b) {
    for (uword q = 0; q < idx->n_elem; ++q) (*m)[(*idx)[q]] = b[q];
  }
};
struct Elem2 {
  Mat* m;
  const UVec* idx;
  operator Mat() const {
    Mat o(idx->n_elem, m->n_cols);
    for (uword k = 0; k < m->n_cols; ++k)
      for (uword q = 0; q < idx->n_elem; ++q) o(q, k) = (*m)((*idx)[q], k);
    return o;
  }
  void operator=(const Mat& b) {
    for (uword k = 0; k < m->n_cols; ++k)
      for (uword q = 0; q < idx->n_elem; ++q) (*m)((*idx)[q], k) = b(q, k);
  }
  void operator+=(const Mat& b) {
    for (uword k = 0; k < m->n_cols; ++k)
      for (uword q = 0; q < idx->n_elem; ++q) (*m)((*idx)[q], k) += b(q, k);
  }
  void operator-=(const Mat& b) {
    for (uword k = 0; k < m->n_cols; ++k)
      for (uword q = 0; q < idx->n_elem; ++q) (*m)((*idx)[q], k) -= b(q, k);
  }
};
struct EachCol {
  Mat* m;
  void operator-=(const Mat& v) {
    for (uword k = 0; k < m->n_cols; ++k)
      for (uword i = 0; i < m->n_rows; ++i) (*m)(i, k) -= v[i];
  }
  void operator/=(const Mat& v) {
    for (uword k = 0; k < m->n_cols; ++k)
      for (uword i = 0; i < m->n_rows; ++i) (*m)(i, k) /= v[i];
  }
};
struct EachRow {
  Mat* m;
  void operator-=(const Mat& v) {
    for (uword k = 0; k < m->n_cols; ++k)
      for (uword i = 0; i < m->n_rows; ++i) (*m)(i, k) -= v[k];
  }
  void operator/=(const Mat& v) {
    for (uword k = 0; k < m->n_cols; ++k)
      for (uword i = 0; i < m->n_rows; ++i) (*m)(i, k) /= v[k];
  }
};

inline Elem1 Mat::operator()(const UVec& idx) { return Elem1{this, &idx}; }
inline Elem1 Mat::operator()(const UVec& idx) const { return Elem1{const_cast<Mat*>(this), &idx}; }
inline Elem2 Mat::rows(const UVec& idx) { return Elem2{this, &idx}; }
inline Elem2 Mat::rows(const UVec& idx) const { return Elem2{const_cast<Mat*>(this), &idx}; }
inline SubCols Mat::tail_cols(uword k) { return SubCols{this, n_cols - k, k}; }
inline SubCols Mat::tail_cols(uword k) const { return SubCols{const_cast<Mat*>(this), n_cols - k, k}; }
inline SubCols Mat::col(uword c) { return SubCols{this, c, 1}; }
inline SubVec Mat::tail(uword k) { return SubVec{this, n_elem - k, k}; }
inline EachCol Mat::each_col() { return EachCol{this}; }
inline EachCol Mat::each_col() const { return EachCol{const_cast<Mat*>(this)}; }
inline EachRow Mat::each_row() { return EachRow{this}; }
inline EachRow Mat::each_row() const { return EachRow{const_cast<Mat*>(this)}; }
inline Col Mat::diag() const {
  Col o(n_rows < n_cols ? n_rows : n_cols);
  for (uword i = 0; i < o.n_elem; ++i) o[i] = (*this)(i, i);
  return o;
}

// ---------------------------------------------------------------- element-wise operators
#define ARMA_SHIM_BINOP(OP)                                                        \
  inline Mat operator OP(const Mat& a, const Mat& b) {                             \
    Mat o(a.n_rows, a.n_cols);                                                     \
    for (uword i = 0; i < a.n_elem; ++i) o[i] = a[i] OP b[i];                      \
    return o;                                                                      \
  }                                                                                \
  inline Mat operator OP(const Mat& a, double s) {                                 \
    Mat o(a.n_rows, a.n_cols);                                                     \
    for (uword i = 0; i < a.n_elem; ++i) o[i] = a[i] OP s;                         \
    return o;                                                                      \
  }                                                                                \
  inline Mat operator OP(double s, const Mat& a) {                                 \
    Mat o(a.n_rows, a.n_cols);                                                     \
    for (uword i = 0; i < a.n_elem; ++i) o[i] = s OP a[i];                         \
    return o;                                                                      \
  }
ARMA_SHIM_BINOP(+)
ARMA_SHIM_BINOP(-)
ARMA_SHIM_BINOP(/)
#undef ARMA_SHIM_BINOP

// `*`: scalar scaling or matrix product (only gendata_FAM_helper multiplies matrices).
inline Mat operator*(const Mat& a, double s) {
  Mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; ++i) o[i] = a[i] * s;
  return o;
}
inline Mat operator*(double s, const Mat& a) {
  Mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; ++i) o[i] = s * a[i];
  return o;
}
inline Mat operator*(const Mat& a, const Mat& b) {
  Mat o(a.n_rows, b.n_cols);
  for (uword j = 0; j < b.n_cols; ++j)
    for (uword l = 0; l < a.n_cols; ++l) {
      double blj = b(l, j);
      for (uword i = 0; i < a.n_rows; ++i) o(i, j) += a(i, l) * blj;
    }
  return o;
}
// `%`: Schur product
inline Mat operator%(const Mat& a, const Mat& b) {
  Mat o(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; ++i) o[i] = a[i] * b[i];
  return o;
}
// vector (op) each column of a matrix
inline Mat operator%(const Mat& v, const EachCol& e) {
  Mat o(e.m->n_rows, e.m->n_cols);
  for (uword k = 0; k < o.n_cols; ++k)
    for (uword i = 0; i < o.n_rows; ++i) o(i, k) = v[i] * (*e.m)(i, k);
  return o;
}
inline Mat operator-(const EachCol& e, const Mat& v) {
  Mat o(e.m->n_rows, e.m->n_cols);
  for (uword k = 0; k < o.n_cols; ++k)
    for (uword i = 0; i < o.n_rows; ++i) o(i, k) = (*e.m)(i, k) - v[i];
  return o;
}
inline Mat operator/(const EachCol& e, const Mat& v) {
  Mat o(e.m->n_rows, e.m->n_cols);
  for (uword k = 0; k < o.n_cols; ++k)
    for (uword i = 0; i < o.n_rows; ++i) o(i, k) = (*e.m)(i, k) / v[i];
  return o;
}
inline Mat operator-(const EachRow& e, const Mat& v) {
  Mat o(e.m->n_rows, e.m->n_cols);
  for (uword k = 0; k < o.n_cols; ++k)
    for (uword i = 0; i < o.n_rows; ++i) o(i, k) = (*e.m)(i, k) - v[k];
  return o;
}

#define ARMA_SHIM_UNARY(NAME, EXPR)                          \
  inline Mat NAME(const Mat& a) {                            \
    Mat o(a.n_rows, a.n_cols);                               \
    for (uword i = 0; i < a.n_elem; ++i) {                   \
      const double x = a[i];                                 \
      o[i] = (EXPR);                                         \
    }                                                        \
    return o;                                                \
  }
ARMA_SHIM_UNARY(exp, std::exp(x))
ARMA_SHIM_UNARY(log, std::log(x))
ARMA_SHIM_UNARY(sqrt, std::sqrt(x))
ARMA_SHIM_UNARY(square, x* x)
#undef ARMA_SHIM_UNARY

// ---------------------------------------------------------------- reductions (left-to-right)
inline double accu(const Mat& a) {
  double s = 0;
  for (uword i = 0; i < a.n_elem; ++i) s += a[i];
  return s;
}
inline double sum(const Mat& a) { return accu(a); }  // only ever applied to vectors in the reference
inline Mat sum(const Mat& a, int dim) {
  if (dim == 0) {
    Mat o(1, a.n_cols);
    for (uword j = 0; j < a.n_cols; ++j) {
      double s = 0;
      for (uword i = 0; i < a.n_rows; ++i) s += a(i, j);
      o[j] = s;
    }
    return o;
  }
  Mat o(a.n_rows, 1);  // zeros, then += each column in turn
  for (uword j = 0; j < a.n_cols; ++j)
    for (uword i = 0; i < a.n_rows; ++i) o[i] += a(i, j);
  return o;
}
inline Mat max(const Mat& a, int dim) {
  // only dim == 1 (row-wise) is used
  (void)dim;
  Mat o(a.n_rows, 1);
  for (uword i = 0; i < a.n_rows; ++i) {
    double m = a(i, 0);
    for (uword j = 1; j < a.n_cols; ++j) m = a(i, j) > m ? a(i, j) : m;
    o[i] = m;
  }
  return o;
}
inline Mat mean(const Mat& a, int dim) {
  (void)dim;  // dim == 1
  Mat o = sum(a, 1);
  for (uword i = 0; i < o.n_elem; ++i) o[i] /= (double)a.n_cols;
  return o;
}
inline Mat var(const Mat& a, int norm_type, int dim) {
  (void)dim;  // dim == 1
  Mat mu = mean(a, 1);
  Mat o(a.n_rows, 1);
  for (uword i = 0; i < a.n_rows; ++i) {
    double acc = 0;
    for (uword j = 0; j < a.n_cols; ++j) acc += (a(i, j) - mu[i]) * (a(i, j) - mu[i]);
    o[i] = acc / (norm_type == 1 ? (double)a.n_cols : (double)(a.n_cols - 1));
  }
  return o;
}
inline Mat median(const Mat& a) {
  Mat o(1, a.n_cols);
  for (uword j = 0; j < a.n_cols; ++j) {
    std::vector<double> c(a.colptr(j), a.colptr(j) + a.n_rows);
    std::sort(c.begin(), c.end());
    uword h = c.size() / 2;
    o[j] = (c.size() % 2) ? c[h] : c[h - 1] + (c[h] - c[h - 1]) / 2;
  }
  return o;
}
inline Mat stddev(const Mat& a) {
  Mat o(1, a.n_cols);
  for (uword j = 0; j < a.n_cols; ++j) {
    double mu = 0;
    for (uword i = 0; i < a.n_rows; ++i) mu += a(i, j);
    mu /= (double)a.n_rows;
    double acc = 0;
    for (uword i = 0; i < a.n_rows; ++i) acc += (a(i, j) - mu) * (a(i, j) - mu);
    o[j] = std::sqrt(acc / (double)(a.n_rows - 1));
  }
  return o;
}

// ---------------------------------------------------------------- generators / reshaping
inline Col linspace(double a, double b, uword n) {
  Col o(n);
  if (n == 1) { o[0] = b; return o; }
  double d = (b - a) / (double)(n - 1);
  for (uword i = 0; i < n - 1; ++i) o[i] = a + (double)i * d;
  o[n - 1] = b;
  return o;
}
inline Mat eye(uword r, uword c) {
  Mat o(r, c);
  for (uword i = 0; i < (r < c ? r : c); ++i) o(i, i) = 1.0;
  return o;
}
inline SizeMat size(const Mat& a) { return SizeMat{a.n_rows, a.n_cols}; }
inline Mat reshape(const Mat& a, uword r, uword c) { return Mat(a.memptr(), r, c); }
inline Mat reshape(const Mat& a, SizeMat s) { return Mat(a.memptr(), s.n_rows, s.n_cols); }

}  // namespace arma

// ---------------------------------------------------------------- Rcpp::List / Named (result marshalling)
namespace Rcpp {

class List;
struct Value {
  enum Kind { kInt, kDouble, kMat, kCube, kList } kind = kInt;
  long long i = 0;
  double d = 0;
  std::shared_ptr<arma::Mat> m;
  std::shared_ptr<arma::Cube> q;
  std::shared_ptr<List> l;
  Value() = default;
  Value(int v) : kind(kInt), i(v) {}
  Value(double v) : kind(kDouble), d(v) {}
  Value(const arma::Mat& v) : kind(kMat), m(std::make_shared<arma::Mat>(v)) {}
  Value(const arma::Cube& v) : kind(kCube), q(std::make_shared<arma::Cube>(v)) {}
  inline Value(const List& v);
};
struct NamedValue {
  std::string name;
  Value value;
};
struct Named {
  std::string name;
  explicit Named(const char* n) : name(n) {}
  template <class T>
  NamedValue operator=(const T& v) const { return NamedValue{name, Value(v)}; }
};
class List {
 public:
  std::vector<NamedValue> items;
  template <class... A>
  static List create(const A&... a) {
    List l;
    (l.items.push_back(a), ...);
    return l;
  }
  const Value& at(const std::string& n) const {
    for (const auto& it : items)
      if (it.name == n) return it.value;
    throw Rcpp::exception(("no list element named " + n).c_str());
  }
};
inline Value::Value(const List& v) : kind(kList), l(std::make_shared<List>(v)) {}

}  // namespace Rcpp
