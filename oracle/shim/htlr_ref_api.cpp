// TEST INFRASTRUCTURE — C entry point around the reference's sampler compiled verbatim:
// /root/reference/src/{main,gibbs,sampler,utils,ars}.cpp against shim/RcppArmadillo.h + shim/Rcpp.h.
// Calls the reference's own `htlr_fit_helper` (main.cpp:7-25) and unpacks the result list
// (gibbs.cpp:128-152). Random variates are served in call order from three sequential streams
// (normals, uniforms, gammas), i.e. exactly how the reference consumes R's RNG.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>

#include "RcppArmadillo.h"

Rcpp::List htlr_fit_helper(int p, int K, int n, arma::mat& X, arma::mat& ymat, arma::uvec& ybase,
                           std::string ptype, double alpha, double s, double eta, int iters_rmc,
                           int iters_h, int thin, int leap_L, int leap_L_h, double leap_step,
                           double hmc_sgmcut, arma::mat& deltas, arma::vec& sigmasbt,
                           bool keep_warmup_hist, int silence, bool legacy);
#ifndef HTLR_GLUE_ONLY   // (-DHTLR_GLUE_ONLY: this file around the repo's own Rcpp glue instead of the reference's main.cpp:
                         // only the htlr_fit_helper entry point exists then - tests/test_rcpp_glue.py)
Rcpp::List std_helper(const arma::mat& A);
arma::vec comp_vardeltas(const arma::mat& deltas);
arma::vec comp_lsl(arma::mat& A);
double log_normcons(arma::mat& A);
Rcpp::List gendata_FAM_helper(int n, arma::mat& muj, const arma::mat& muj_rep, const arma::mat& A,
                              double sd_g, bool stdx);
#endif

namespace {
struct Streams {
  const double *normals = nullptr, *uniforms = nullptr, *gammas = nullptr;
  int64_t n_normals = 0, n_uniforms = 0, n_gammas = 0;
  int64_t c_normals = 0, c_uniforms = 0, c_gammas = 0;
  // keyed fallbacks when no array is given: a tiny LCG-free splitmix stream, enough for timing runs
  uint64_t sm = 0x9E3779B97F4A7C15ull;
  double splitmix() {
    uint64_t z = (sm += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ((double)(z >> 12) + 0.5) * 2.220446049250313e-16;
  }
} g_s;
std::string g_log;
thread_local std::string g_err;

double hook_unif(void*) {
  if (!g_s.uniforms) return g_s.splitmix();
  if (g_s.c_uniforms >= g_s.n_uniforms) throw Rcpp::exception("injected uniforms exhausted");
  return g_s.uniforms[g_s.c_uniforms++];
}
double hook_norm(void*) {
  if (!g_s.normals) {
    double u1 = g_s.splitmix(), u2 = g_s.splitmix();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  if (g_s.c_normals >= g_s.n_normals) throw Rcpp::exception("injected normals exhausted");
  return g_s.normals[g_s.c_normals++];
}
double hook_gamma(void*, double shape) {
  if (!g_s.gammas) {
    // Marsaglia-Tsang on the fallback stream (timing runs only)
    double a = shape < 1.0 ? shape + 1.0 : shape, d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d), g;
    for (;;) {
      double x = hook_norm(nullptr), v = 1.0 + c * x;
      if (v <= 0) continue;
      v = v * v * v;
      double u = g_s.splitmix();
      if (std::log(u) < 0.5 * x * x + d - d * v + d * std::log(v)) { g = d * v; break; }
    }
    if (shape < 1.0) g *= std::pow(g_s.splitmix(), 1.0 / shape);
    return g;
  }
  if (g_s.c_gammas >= g_s.n_gammas) throw Rcpp::exception("injected gammas exhausted");
  return g_s.gammas[g_s.c_gammas++];
}
void hook_print(void*, const char* t) { g_log += t; }

void copy_out(const arma::Mat& m, double* dst) {
  if (dst) std::memcpy(dst, m.memptr(), sizeof(double) * m.n_elem);
}
}  // namespace

rshim::Hooks& rshim::hooks() {
  static Hooks h;
  return h;
}
std::map<std::string, std::string>& rshim::options() {
  static std::map<std::string, std::string> o;
  return o;
}

extern "C" {

struct ref_cfg {
  int32_t p, K, n, ptype;
  double alpha, s, eta;
  int32_t iters_rmc, iters_h, thin, leap_L, leap_L_h;
  double leap_step, hmc_sgmcut;
  int32_t keep_warmup_hist, legacy, silence, pad_;
};

const char* htlr_ref_last_error() { return g_err.c_str(); }
// options(name = value) for the code under test; an empty value removes the option
void htlr_ref_set_option(const char* name, const char* value) {
  if (value && *value) rshim::options()[name] = value; else rshim::options().erase(name);
}
const char* htlr_ref_log() { return g_log.c_str(); }

// Any of the three stream pointers may be null: that stream then comes from an internal generator
// (used for timing runs, where only the law matters).
int htlr_ref_fit(const ref_cfg* c, const double* X, const double* ymat, const int64_t* ybase,
                 const double* deltas, const double* sigmasbt, const double* normals, int64_t n_normals,
                 const double* uniforms, int64_t n_uniforms, const double* gammas, int64_t n_gammas,
                 double* mcdeltas, double* mcsigmasbt, double* mcvardeltas, double* mclogw,
                 double* mcloglike, double* mcuvar, double* mchmcrej, double* ddnloglike,
                 int64_t* consumed /* [3] normals, uniforms, gammas */, double* secs) {
  rshim::Hooks& h = rshim::hooks();
  h.unif = hook_unif; h.norm = hook_norm; h.gamma = hook_gamma; h.print = hook_print;
  g_s = Streams();
  g_s.normals = normals; g_s.n_normals = n_normals;
  g_s.uniforms = uniforms; g_s.n_uniforms = n_uniforms;
  g_s.gammas = gammas; g_s.n_gammas = n_gammas;
  g_log.clear();
  try {
    const int nvar = c->p + 1;
    arma::mat Xm(X, c->n, nvar), ym(ymat, c->n, c->K), dl(deltas, nvar, c->K);
    arma::vec sg(sigmasbt, nvar);
    std::vector<arma::uword> yb(ybase, ybase + c->n);
    arma::uvec ybv(yb.data(), (arma::uword)c->n);
    const char* names[] = {"t", "ghs", "neg", "bogus"};
    auto t0 = std::chrono::steady_clock::now();
    Rcpp::List out = htlr_fit_helper(c->p, c->K, c->n, Xm, ym, ybv, names[c->ptype], c->alpha, c->s,
                                     c->eta, c->iters_rmc, c->iters_h, c->thin, c->leap_L, c->leap_L_h,
                                     c->leap_step, c->hmc_sgmcut, dl, sg, c->keep_warmup_hist != 0,
                                     c->silence, c->legacy != 0);
    if (secs) *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (mcdeltas) {
      const arma::Cube& q = *out.at("mcdeltas").q;
      std::memcpy(mcdeltas, q.memptr(), sizeof(double) * q.n_elem);
    }
    copy_out(*out.at("mcsigmasbt").m, mcsigmasbt);
    copy_out(*out.at("mcvardeltas").m, mcvardeltas);
    copy_out(*out.at("mclogw").m, mclogw);
    copy_out(*out.at("mcloglike").m, mcloglike);
    copy_out(*out.at("mcuvar").m, mcuvar);
    copy_out(*out.at("mchmcrej").m, mchmcrej);
    copy_out(*out.at("mc.param").l->at("DDNloglike").m, ddnloglike);
    if (consumed) {
      consumed[0] = g_s.c_normals; consumed[1] = g_s.c_uniforms; consumed[2] = g_s.c_gammas;
    }
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
  return 0;
}

#ifndef HTLR_GLUE_ONLY
// The reference's exported helpers with restatable known-answer tests (tests/testthat/test-util.r).
void htlr_ref_comp_vardeltas(const double* d, int32_t rows, int32_t cols, double* out) {
  arma::mat m(d, rows, cols);
  copy_out(comp_vardeltas(m), out);
}
void htlr_ref_comp_lsl(const double* a, int32_t rows, int32_t cols, double* out) {
  arma::mat m(a, rows, cols);
  copy_out(comp_lsl(m), out);
}
double htlr_ref_log_normcons(const double* a, int32_t rows, int32_t cols) {
  arma::mat m(a, rows, cols);
  return log_normcons(m);
}
// std_helper (utils.cpp:36-48): X_out n x p, median p, sd p
void htlr_ref_std(const double* a, int32_t rows, int32_t cols, double* X_out, double* med, double* sd) {
  arma::mat m(a, rows, cols);
  Rcpp::List l = std_helper(m);
  copy_out(*l.at("X").m, X_out);
  copy_out(*l.at("median").m, med);
  copy_out(*l.at("sd").m, sd);
}
// gendata_FAM_helper (utils.cpp:72-105) with dense A; normals from the injected stream (n*k then n*p).
int htlr_ref_gendata_fam(int32_t n, int32_t p, int32_t nc, int32_t k, const double* muj,
                         const double* muj_rep, const double* A, double sd_g, int32_t stdx,
                         const double* normals, int64_t n_normals, double* X_out /* n x p */,
                         double* muj_out /* p x nc */) {
  rshim::Hooks& h = rshim::hooks();
  h.unif = hook_unif; h.norm = hook_norm; h.gamma = hook_gamma; h.print = hook_print;
  g_s = Streams();
  g_s.normals = normals; g_s.n_normals = n_normals;
  try {
    arma::mat mj(muj, p, nc), mr(muj_rep, p, n), Am(A, p, k);
    Rcpp::List l = gendata_FAM_helper(n, mj, mr, Am, sd_g, stdx != 0);
    copy_out(*l.at("X").m, X_out);
    copy_out(*l.at("muj").m, muj_out);
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
  return 0;
}
#endif   // HTLR_GLUE_ONLY
}
