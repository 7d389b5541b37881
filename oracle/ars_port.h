// TEST INFRASTRUCTURE — CPU oracle, never shipped or measured as product.
//
// Restatement of the reference's Gilks-Wild adaptive rejection sampler for ONE draw
// (the only way gibbs.cpp uses it: `ARS(1, target, x0).Sample()[0]`, gibbs.cpp:344-345,358,367).
// Follows /root/reference/src/ars.cpp arithmetic statement by statement:
//   ctor            ars.cpp:204-242   first tangent at x0; abort if logf(x0) not finite (:217-222)
//   step-out        ars.cpp:138-202   left then right by `stepout` until slope crosses +-tol or logf=-inf
//                                     (left loop always passes h=0 to the insert: quirk 7, ars.cpp:161-167)
//   insert          ars.cpp:22-119    update_hulls
//   draw loop       ars.cpp:272-309   sample_disc (:334-356), sample_elin (:359-417), squeeze test,
//                                     target eval + insert, accept test
//   piece helpers   ars.cpp:424-462   logint_elin, interc
// Layout differs (array of hull records, uniform source object) — arithmetic and branch order do not.
// Validated bit-for-bit against the reference ars.cpp compiled verbatim (oracle/_ref, tests/test_oracle_pinned.py).
#pragma once
#include <cmath>
#include <limits>
#include <stdexcept>
#include <vector>

namespace ars_port {

struct Target {
  virtual void eval(double x, double& logf, double& dlogf) = 0;
  virtual ~Target() = default;
};

// Uniform variates are consumed strictly in the reference's call order.
struct Uniforms {
  virtual double next() = 0;
  virtual ~Uniforms() = default;
};

struct Stats {
  int n_eval = 0, n_unif = 0, n_hulls = 0, n_rej = 0;
};

inline double piece_logint(double f, double df, double t, double lo, double hi, double tol) {
  // ars.cpp:424-450
  double dx = hi - lo;
  double adf = std::fabs(df);
  if (adf <= tol) return f + std::log(dx);
  if (df > tol) return f + df * (hi - t) - std::log(adf) + std::log(1 - std::exp(-adf * dx));
  return f + df * (lo - t) - std::log(adf) + std::log(1 - std::exp(-adf * dx));
}

inline double tangent_cross(double t1, double t2, double f1, double f2, double d1, double d2,
                            double tol) {
  // ars.cpp:453-462
  if (std::fabs(d1 - d2) > tol) return (f2 - f1 - d2 * t2 + d1 * t1) / (d1 - d2);
  return (t1 + t2) / 2.0;
}

class Sampler {
 public:
  struct Hull {
    double t = 0, lw = 0, lo = 0, hi = 0, f = 0, df = 0, sl = 0, sr = 0;
    int left = 0, right = 0;
  };

  Sampler(Target* target, Uniforms* unif, double x0, int max_nhull = 1000, double stepout = 10,
          double tol_d = 1e-5, double tol_dd = 1e-5)
      : tg_(target), un_(unif), cap_(max_nhull), stepout_(stepout), tol_d_(tol_d), tol_dd_(tol_dd),
        h_(max_nhull) {
    const double inf = std::numeric_limits<double>::infinity();
    h_[0].t = x0;
    eval(x0, h_[0].f, h_[0].df);
    if (!std::isfinite(h_[0].f))
      throw std::runtime_error(
          "Error in adaptive rejection sampling:\n"
          "the first tangent point doesn't have positive probability.\n");
    h_[0].lo = -inf;
    h_[0].hi = inf;
    h_[0].left = -1;
    h_[0].right = cap_;
    h_[0].sl = inf;
    h_[0].sr = -inf;
    h_[0].lw = inf;
    nh_ = 1;
  }

  double draw() {
    step_out();
    for (;;) {
      int h = pick_hull();
      double x = pick_point(h_[h].lo, h_[h].hi, h_[h].df);
      double upper = (x - h_[h].t) * h_[h].df + h_[h].f;
      double u = unif();
      double logacc = upper + std::log(u);
      double lower = x >= h_[h].t ? (x - h_[h].t) * h_[h].sr + h_[h].f
                                  : (x - h_[h].t) * h_[h].sl + h_[h].f;
      if (logacc <= lower) return finish(x);
      double f, df;
      eval(x, f, df);
      insert(h, x, f, df);
      if (logacc <= f) return finish(x);
      ++st_.n_rej;
    }
  }

  const Stats& stats() const { return st_; }
  const std::vector<Hull>& hulls() const { return h_; }
  int n_hulls() const { return nh_; }

 private:
  Target* tg_;
  Uniforms* un_;
  const int cap_;
  const double stepout_, tol_d_, tol_dd_;
  std::vector<Hull> h_;
  int nh_ = 0;
  Stats st_;

  void eval(double x, double& f, double& df) {
    tg_->eval(x, f, df);
    ++st_.n_eval;
  }
  double unif() {
    ++st_.n_unif;
    return un_->next();
  }
  double finish(double x) {
    st_.n_hulls = nh_;
    return x;
  }
  void relog(int i) { h_[i].lw = piece_logint(h_[i].f, h_[i].df, h_[i].t, h_[i].lo, h_[i].hi, tol_d_); }

  void insert(int h, double x, double f, double df) {
    const double inf = std::numeric_limits<double>::infinity();
    if (nh_ == cap_) return;
    int lh, rh;
    if (x > h_[h].t) {
      lh = h;
      rh = h_[h].right;
      if ((rh == cap_) & (f == -inf)) {
        if (h_[h].hi != x) {
          h_[h].hi = x;
          relog(h);
        }
        return;
      }
    } else {
      lh = h_[h].left;
      rh = h;
      if ((lh == -1) & (f == -inf)) {
        if (h_[h].lo != x) {
          h_[h].lo = x;
          relog(h);
        }
        return;
      }
    }
    int nh = nh_++;
    Hull& N = h_[nh];
    N.t = x;
    N.f = f;
    N.df = df;
    N.left = lh;
    N.right = rh;
    if (lh == -1) {
      N.lo = h_[h].lo;
      N.sl = inf;
    } else {
      N.lo = tangent_cross(h_[lh].t, N.t, h_[lh].f, N.f, h_[lh].df, N.df, tol_dd_);
      N.sl = (N.f - h_[lh].f) / (N.t - h_[lh].t);
    }
    if (rh == cap_) {
      N.hi = h_[h].hi;
      N.sr = -inf;
    } else {
      N.hi = tangent_cross(N.t, h_[rh].t, N.f, h_[rh].f, N.df, h_[rh].df, tol_dd_);
      N.sr = (N.f - h_[rh].f) / (N.t - h_[rh].t);
    }
    relog(nh);
    if (lh != -1) {
      h_[lh].hi = N.lo;
      h_[lh].right = nh;
      h_[lh].sr = N.sl;
      relog(lh);
    }
    if (rh != cap_) {
      h_[rh].lo = N.hi;
      h_[rh].left = nh;
      h_[rh].sl = N.sr;
      relog(rh);
    }
  }

  void step_out() {
    const double inf = std::numeric_limits<double>::infinity();
    double x, f, df;
    // unbounded support on both sides (the only way gibbs.cpp constructs it)
    x = h_[0].t - stepout_;
    do {
      if (nh_ == cap_)
        throw std::runtime_error(
            "Error in Rejection Sampling: (finite lb)\n"
            "'max_nhull_' is set too small, or your log-PDF is NOT concave.\n");
      eval(x, f, df);
      insert(0, x, f, df);  // always hull 0 on the left side (ars.cpp:161)
      if (f == -inf) break;
      x -= stepout_;
    } while (df < tol_d_);

    int h = 0;
    x = h_[0].t + stepout_;
    do {
      if (nh_ == cap_)
        throw std::runtime_error(
            "Error in Rejection Sampling: (finite ub)\n"
            "'max_nhull' is set too small, or your log-PDF is NOT concave.\n");
      eval(x, f, df);
      insert(h, x, f, df);
      if (!std::isfinite(f)) break;
      x += stepout_;
      h = nh_ - 1;
    } while (df > -tol_d_);
  }

  int pick_hull() {
    // ars.cpp:334-356 — max-shifted cumulative weights, linear scan
    double mx = h_[0].lw;
    for (int i = 1; i < nh_; ++i) mx = std::fmax(h_[i].lw, mx);
    std::vector<double> cw(nh_);
    cw[0] = std::exp(h_[0].lw - mx);
    for (int i = 1; i < nh_; ++i) cw[i] = cw[i - 1] + std::exp(h_[i].lw - mx);
    double u = unif() * cw[nh_ - 1];
    int i = 0;
    while (i < nh_) {
      if (u <= cw[i]) break;
      ++i;
    }
    return i;
  }

  double pick_point(double lo, double hi, double df) {
    // ars.cpp:359-417
    int type = -1;
    bool fault = false;
    if (std::fabs(df) <= tol_d_) {
      if (!(std::isfinite(lo) & std::isfinite(hi))) fault = true; else type = 0;
    }
    if (df > tol_d_) {
      if (!std::isfinite(hi)) fault = true; else type = 1;
    }
    if (df < -tol_d_) {
      if (!std::isfinite(lo)) fault = true; else type = 2;
    }
    if (fault)
      throw std::runtime_error(
          "Error: in C function 'sample_elin':\n"
          "the exp linear function integrates to NAN/INFINITY\n");
    double dx = hi - lo;
    double y = unif();
    if (type == 0) return lo + y * dx;
    if (type == 1) return hi + std::log((1 - y) * std::exp(-df * dx) + y) / df;
    if (type == 2) return lo + std::log(1 - y + y * std::exp(df * dx)) / df;
    throw std::runtime_error("Error: in C function 'sample_elin': unexpected type_lin value\n");
  }
};

// ---- the three targets of sampler.cpp ------------------------------------------------------------

// sampler.cpp:18-30 — NEG prior, x = log sigma^2_j
struct NegTarget : Target {
  double v, K, alpha, log_aw;
  void eval(double x, double& logf, double& dlogf) override {
    logf = -(K / 2.0 - 1.0) * x;
    dlogf = -(K / 2.0 - 1.0);
    double vexi = v / 2.0 / std::exp(x);
    logf -= vexi;
    dlogf += vexi;
    double e = std::exp(x - log_aw);
    logf -= (alpha / 2.0 + 1.0) * std::log(1.0 + 2.0 * e);
    dlogf -= (alpha + 2.0) * e / (1.0 + 2.0 * e);
  }
};

// sampler.cpp:37-49 — generalised horseshoe
struct GhsTarget : Target {
  double v, K, alpha, log_aw;
  void eval(double x, double& logf, double& dlogf) override {
    logf = -(K - 1.0) / 2.0 * x;
    dlogf = -(K - 1.0) / 2.0;
    double vexi = v / 2.0 / std::exp(x);
    logf -= vexi;
    dlogf += vexi;
    double e = std::exp(x - log_aw);
    logf -= (alpha + 1.0) / 2.0 * std::log(1.0 + e);
    dlogf -= (alpha + 1.0) / 2.0 * e / (1.0 + e);
  }
};

// sampler.cpp:57-72 — log w | var_deltas (t prior, eta > 0); O(p) per evaluation
struct LogwTarget : Target {
  const double* vard;  // var_deltas[1..p]
  int p;
  double K, nu, s, eta;
  void eval(double x, double& logf, double& dlogf) override {
    double w = std::exp(x);
    double sdu = (x - s) / eta;
    double wnu = w * nu;
    double a = 0, b = 0;
    for (int j = 0; j < p; ++j) a += wnu / (vard[j] + wnu);
    for (int j = 0; j < p; ++j) b += std::log(vard[j] + wnu);
    dlogf = a;
    logf = b;
    logf *= (-(nu + K) / 2);
    dlogf *= (-(nu + K) / 2);
    logf += p * nu / 2 * x;
    dlogf += p * nu / 2;
    logf += -(sdu * sdu) / 2 - std::log(eta);
    dlogf += -sdu / eta;
  }
};

}  // namespace ars_port
