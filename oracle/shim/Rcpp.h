// TEST INFRASTRUCTURE — a stand-in for <Rcpp.h>, written for this repo (not Rcpp code).
// Lets the reference's sources under /root/reference/src be compiled *verbatim* without R/Rcpp so
// their results can pin the oracle restatement. Only the handful of names those files use exist.
// Random variates are routed to caller-installed hooks (rshim::hooks()), consumed in call order,
// which is exactly how the reference consumes R's global RNG stream.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <initializer_list>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rshim {
struct Hooks {
  double (*unif)(void* ctx) = nullptr;               // U(0,1)
  double (*norm)(void* ctx) = nullptr;               // N(0,1)
  double (*gamma)(void* ctx, double shape) = nullptr;  // Gamma(shape, 1)
  void (*print)(void* ctx, const char* text) = nullptr;
  void* ctx = nullptr;
};
Hooks& hooks();  // defined once, in the *_ref_api.cpp translation unit

inline std::string vformat(const char* fmt, va_list ap) {
  char buf[1024];
  vsnprintf(buf, sizeof buf, fmt, ap);
  return buf;
}
// Rcpp::stop / warning use tinyformat, which accepts std::string for %s.
inline const char* fmt_arg(const std::string& s) { return s.c_str(); }
template <class T>
inline T fmt_arg(T v) { return v; }
}  // namespace rshim

#define R_NegInf (-std::numeric_limits<double>::infinity())
#define R_PosInf (std::numeric_limits<double>::infinity())
#define R_FINITE(x) std::isfinite(x)

inline void GetRNGstate() {}
inline void PutRNGstate() {}
inline void R_CheckUserInterrupt() {}
inline double unif_rand() { return rshim::hooks().unif(rshim::hooks().ctx); }
inline double norm_rand() { return rshim::hooks().norm(rshim::hooks().ctx); }
inline double R_pow_di(double x, int n) {
  // R nmath R_pow_di: repeated squaring
  double xn = 1.0;
  if (std::isnan(x)) return x;
  if (n != 0) {
    if (!std::isfinite(x)) return std::pow(x, (double)n);
    bool neg = n < 0;
    if (neg) n = -n;
    for (;;) {
      if (n & 01) xn *= x;
      if (n >>= 1) x *= x; else break;
    }
    if (neg) xn = 1. / xn;
  }
  return xn;
}

inline void Rprintf(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::string s = rshim::vformat(fmt, ap);
  va_end(ap);
  if (rshim::hooks().print) rshim::hooks().print(rshim::hooks().ctx, s.c_str());
}
inline void REprintf(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::string s = rshim::vformat(fmt, ap);
  va_end(ap);
  if (rshim::hooks().print) rshim::hooks().print(rshim::hooks().ctx, s.c_str());
}

namespace R {
inline double runif(double a, double b) { return a + (b - a) * unif_rand(); }       // nmath runif.c
inline double rnorm(double mu, double sigma) { return mu + sigma * norm_rand(); }   // nmath rnorm.c
inline double rgamma(double a, double scale) {                                       // nmath rgamma.c
  return scale * rshim::hooks().gamma(rshim::hooks().ctx, a);
}
}  // namespace R

namespace Rcpp {

struct exception : std::runtime_error {
  using std::runtime_error::runtime_error;
};

template <class... A>
[[noreturn]] inline void stop(const char* fmt, A... args) {
  char buf[1024];
  if constexpr (sizeof...(A) == 0) snprintf(buf, sizeof buf, "%s", fmt);
  else snprintf(buf, sizeof buf, fmt, rshim::fmt_arg(args)...);
  throw Rcpp::exception(buf);
}

class NumericVector {
 public:
  NumericVector() = default;
  explicit NumericVector(int n) : v_(n, 0.0) {}
  double& operator[](int i) { return v_[i]; }
  const double& operator[](int i) const { return v_[i]; }
  int size() const { return (int)v_.size(); }
  int length() const { return (int)v_.size(); }
  const double* begin() const { return v_.data(); }
  const double* end() const { return v_.data() + v_.size(); }
  double* begin() { return v_.data(); }
  double* end() { return v_.data() + v_.size(); }
 private:
  std::vector<double> v_;
};

// getOption(), as<>, checkUserInterrupt: what the repo's own Rcpp glue (htlr_b200/host/rcpp_glue.cpp) uses beyond the
// reference's sources. Options come from a map the test harness fills (htlr_ref_set_option).
struct RObject {
  std::string text;
};
}  // namespace Rcpp
namespace rshim {
std::map<std::string, std::string>& options();   // defined in the *_ref_api.cpp translation unit
}
namespace Rcpp {
class Function {
 public:
  explicit Function(const char* name) : name_(name) {}
  template <class T>
  RObject operator()(const char* key, const T& fallback) const {
    if (name_ != "getOption") throw Rcpp::exception(("shim: unsupported R function " + name_).c_str());
    auto it = rshim::options().find(key);
    if (it != rshim::options().end()) return RObject{it->second};
    std::ostringstream os;
    os << fallback;
    return RObject{os.str()};
  }
 private:
  std::string name_;
};
template <class T>
T as(const RObject& o);
template <>
inline int as<int>(const RObject& o) { return std::stoi(o.text); }
template <>
inline std::string as<std::string>(const RObject& o) { return o.text; }
inline void checkUserInterrupt() {}

inline NumericVector runif(int n) {
  NumericVector o(n);
  for (int i = 0; i < n; ++i) o[i] = unif_rand();
  return o;
}
inline NumericVector rnorm(int n) {
  NumericVector o(n);
  for (int i = 0; i < n; ++i) o[i] = norm_rand();
  return o;
}
inline NumericVector rgamma(int n, double shape) {
  NumericVector o(n);
  for (int i = 0; i < n; ++i) o[i] = R::rgamma(shape, 1.0);
  return o;
}
inline NumericVector rexp(int n) {
  NumericVector o(n);
  for (int i = 0; i < n; ++i) o[i] = -std::log(unif_rand());
  return o;
}

}  // namespace Rcpp
