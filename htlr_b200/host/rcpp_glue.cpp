// Rcpp glue: drop-in replacement for the body of the reference's src/main.cpp (htlr_fit_helper, :7-25).
// R/core.r:198-213 and the generated wrapper `_HTLR_htlr_fit_helper` (src/RcppExports.cpp:62-91) stay as
// they are; only this translation unit replaces main.cpp, and src/Makevars gains
//   PKG_LIBS += -L$(HTLR_CUDA_LIB) -lhtlr_cuda
// It needs R + Rcpp + RcppArmadillo to compile, none of which exist in the build container, so it is not
// part of the in-tree build; everything below the C ABI is built and tested without it.
//
// RNG: by default the chain uses the library's device Philox stream keyed from R's RNG (one unif_rand()
// draw inside the RNGScope, so set.seed() still makes runs reproducible). With
// options(HTLR.cuda.rng = "R") the glue hands R's own generator to the host class (htlr::HostRng): every
// variate is drawn from R in the reference's consumption order (gibbs.cpp:441-460 normals, :90 one uniform,
// utils.cpp:31 gammas), one iteration per device call, so that set.seed() gives the reference's own chain to
// rounding — slow, meant for checks; t prior with eta = 0 and thin = 1 only (cuda_fit.hpp explains why).
// tests/test_rcpp_glue.py compiles this file against the repo's Rcpp/Armadillo stand-ins (test infrastructure) and
// compares it, R-order mode included, with the reference's own htlr_fit_helper on the same streams.
#include "RcppArmadillo.h"

#include "cuda_fit.hpp"

// [[Rcpp::depends(RcppArmadillo)]]

// [[Rcpp::export]]
Rcpp::List htlr_fit_helper(int p, int K, int n, arma::mat& X, arma::mat& ymat, arma::uvec& ybase, std::string ptype,
                           double alpha, double s, double eta, int iters_rmc, int iters_h, int thin, int leap_L,
                           int leap_L_h, double leap_step, double hmc_sgmcut, arma::mat& deltas, arma::vec& sigmasbt,
                           bool keep_warmup_hist, int silence, bool legacy) {
  std::vector<int64_t> yb(ybase.begin(), ybase.end());  // arma::uword -> int64_t
  htlr::DeviceOptions opt;
  opt.device = Rcpp::as<int>(Rcpp::Function("getOption")("HTLR.cuda.device", 0));
  const bool r_order = Rcpp::as<std::string>(Rcpp::Function("getOption")("HTLR.cuda.rng", "device")) == "R";
  htlr::HostRng from_r{[] { return R::rnorm(0.0, 1.0); }, [] { return R::runif(0.0, 1.0); },
                       [](double shape) { return R::rgamma(shape, 1.0); }};
  if (r_order) opt.host_rng = &from_r;
  else opt.seed = (uint64_t)(unif_rand() * 9007199254740992.0);  // 53 bits from R's stream (RNGScope is open)
  htlr::FitOutput o;
  try {
    // silence = 1 below: progress is printed here with Rprintf, not by the host class with printf
    htlr::CudaFit fit(p, K, n, X.memptr(), ymat.memptr(), yb.data(), ptype, alpha, s, eta, iters_rmc, iters_h, thin,
                      leap_L, leap_L_h, leap_step, hmc_sgmcut, deltas.memptr(), sigmasbt.memptr(), keep_warmup_hist,
                      1, legacy, opt);
    int printed = 0;
    fit.StartSampling([&](int done, const htlr_iter_stats* st, int chunk) {
      if (silence == 0)
        for (int i = 0; i < chunk; ++i, ++printed)  // gibbs.cpp:119-121
          Rprintf("Iter%4d: deviance=%5.3f, logw=%6.2f, nuvar=%3.0f, hmcrej=%4.2f\n", printed - iters_h,
                  -st[i].loglike / n, st[i].logw, st[i].uvar, st[i].rej);
      // gibbs.cpp:124 polls R_CheckUserInterrupt() every 256 iterations; here every pipelined unit of 8. The
      // throwing form unwinds through ~CudaFit, which waits for the iterations already in flight and frees the
      // device state (a longjmp out of R_CheckUserInterrupt would skip that).
      Rcpp::checkUserInterrupt();
      return false;
    });
    o = fit.Output();
  } catch (const htlr::CudaFitError& e) {
    Rcpp::stop(e.what());
  }
  const int nvar = p + 1, H = o.hist_len;
  arma::rowvec ddn(o.DDNloglike);
  arma::cube mcdeltas(o.mcdeltas.data(), nvar, K, H);
  arma::mat mcsigmasbt(o.mcsigmasbt.data(), nvar, H), mcvardeltas(o.mcvardeltas.data(), nvar, H);
  arma::vec mclogw(o.mclogw), mcloglike(o.mcloglike), mcuvar(o.mcuvar), mchmcrej(o.mchmcrej);
  auto mc_param = Rcpp::List::create(Rcpp::Named("iter.rmc") = iters_rmc, Rcpp::Named("iter.warm") = iters_h,
                                     Rcpp::Named("thin") = thin, Rcpp::Named("leap") = leap_L,
                                     Rcpp::Named("leap.warm") = leap_L_h, Rcpp::Named("leap.step") = leap_step,
                                     Rcpp::Named("sgmsq.cut") = o.sgmsq_cut, Rcpp::Named("DDNloglike") = ddn);
  return Rcpp::List::create(Rcpp::Named("p") = p, Rcpp::Named("n") = n, Rcpp::Named("K") = K,
                            Rcpp::Named("mc.param") = mc_param, Rcpp::Named("mcdeltas") = mcdeltas,
                            Rcpp::Named("mclogw") = mclogw, Rcpp::Named("mcsigmasbt") = mcsigmasbt,
                            Rcpp::Named("mcvardeltas") = mcvardeltas, Rcpp::Named("mcloglike") = mcloglike,
                            Rcpp::Named("mcuvar") = mcuvar, Rcpp::Named("mchmcrej") = mchmcrej);
}
