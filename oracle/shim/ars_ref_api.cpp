// TEST INFRASTRUCTURE — C entry points around the reference's ARS engine compiled verbatim
// (/root/reference/src/ars.cpp + ars.h against shim/Rcpp.h). The log-density targets are the oracle's
// restatements (ars_port.h) wrapped as `SampleTarget`s, so a mismatch isolates the ENGINE.
#include <cstdint>
#include <string>

#include "ars.h"          // the reference's header
#include "../ars_port.h"  // oracle targets only

namespace {
struct Stream {
  const double* u = nullptr;
  int n = 0, c = 0;
};
Stream g_stream;
double next_unif(void*) {
  if (g_stream.c >= g_stream.n) throw Rcpp::exception("ARS uniforms exhausted");
  return g_stream.u[g_stream.c++];
}
struct Adapter : SampleTarget {
  ars_port::Target* t;
  int evals = 0;
  void eval_logf(const double x, double& logf, double& dlogf) override {
    ++evals;
    t->eval(x, logf, dlogf);
  }
};
thread_local std::string g_err;
}  // namespace

rshim::Hooks& rshim::hooks() {
  static Hooks h;
  return h;
}

extern "C" {

const char* ars_ref_last_error() { return g_err.c_str(); }

// kind: 1 ghs, 2 neg. One draw per feature from x0 = log(vard[j]/K), uniforms u[j*width..].
int ars_ref_batch(int32_t kind, int32_t m, const double* vard, double K, double alpha, double log_aw,
                  const double* u, int32_t width, int32_t max_nhull, double* x_out, int32_t* n_unif,
                  int32_t* n_eval) {
  rshim::hooks().unif = next_unif;
  try {
    for (int j = 0; j < m; ++j) {
      ars_port::NegTarget neg;
      ars_port::GhsTarget ghs;
      Adapter ad;
      if (kind == 2) { neg.v = vard[j]; neg.K = K; neg.alpha = alpha; neg.log_aw = log_aw; ad.t = &neg; }
      else { ghs.v = vard[j]; ghs.K = K; ghs.alpha = alpha; ghs.log_aw = log_aw; ad.t = &ghs; }
      g_stream.u = u + (size_t)j * width;
      g_stream.n = width;
      g_stream.c = 0;
      ARS spl(1, &ad, std::log(vard[j] / K), R_NegInf, R_PosInf, false, max_nhull > 0 ? max_nhull : 1000);
      x_out[j] = spl.Sample()[0];
      if (n_unif) n_unif[j] = g_stream.c;
      if (n_eval) n_eval[j] = ad.evals;
    }
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
  return 0;
}

int ars_ref_logw(int32_t p, const double* vard, double K, double nu, double s, double eta, double x0,
                 const double* u, int32_t width, double* x_out, int32_t* n_unif, int32_t* n_eval) {
  rshim::hooks().unif = next_unif;
  try {
    ars_port::LogwTarget tg;
    tg.vard = vard; tg.p = p; tg.K = K; tg.nu = nu; tg.s = s; tg.eta = eta;
    Adapter ad;
    ad.t = &tg;
    g_stream.u = u; g_stream.n = width; g_stream.c = 0;
    ARS spl(1, &ad, x0);
    *x_out = spl.Sample()[0];
    if (n_unif) *n_unif = g_stream.c;
    if (n_eval) *n_eval = ad.evals;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
  return 0;
}
}
