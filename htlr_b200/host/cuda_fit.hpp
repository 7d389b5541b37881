// Host-side C++ mirror of the reference's `class Fit` for the CUDA path.
//
// The reference constructs `Fit(p, K, n, X, ymat, ybase, ptype, alpha, s, eta, iters_rmc, iters_h, thin,
// leap_L, leap_L_h, leap_step, hmc_sgmcut, deltas, sigmasbt, keep_warmup_hist, silence, legacy)`, calls
// `StartSampling()` and returns `OutputR()` (/root/reference/src/main.cpp:7-25, src/gibbs.h:91-100).
// `htlr::CudaFit` has the same three steps with the same argument order and meaning, but owns no
// arithmetic: everything numeric happens behind the C ABI of include/htlr_cuda.h. It depends on nothing
// but the C++ standard library, so the same header serves the Rcpp glue (rcpp_glue.cpp, needs R) and the
// stand-alone driver (fit_host_main.cpp, built and tested here).
//
// Error behaviour mirrors the reference: every failure becomes a C++ exception (`htlr::CudaFitError`)
// with the reference's message text where the reference has one (`Rcpp::stop`, ars.cpp:217-222 etc.);
// the Rcpp glue lets BEGIN_RCPP/END_RCPP turn it into an R error.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/htlr_cuda.h"

namespace htlr {

struct CudaFitError : std::runtime_error {
  int code;
  CudaFitError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// Fit::OutputR (gibbs.cpp:128-152): same names (dots become underscores), same shapes, column-major.
struct FitOutput {
  int p = 0, n = 0, K = 0;
  // mc.param
  int iter_rmc = 0, iter_warm = 0, thin = 0, leap = 0, leap_warm = 0;
  double leap_step = 0, sgmsq_cut = 0;
  std::vector<double> DDNloglike;   // 1 x (p+1)
  int hist_len = 0;                 // H
  std::vector<double> mcdeltas;     // (p+1) x K x H
  std::vector<double> mclogw;       // H x 1
  std::vector<double> mcsigmasbt;   // (p+1) x H
  std::vector<double> mcvardeltas;  // (p+1) x H
  std::vector<double> mcloglike, mcuvar, mchmcrej;  // H x 1
};

// A host random-number stream in the REFERENCE's consumption order (the Rcpp glue points these at R's own
// generator: R::rnorm, R::runif, R::rgamma): per thin-step u*K standard normals for the momenta, feature-major
// (GenMomt, gibbs.cpp:441-460), one uniform for the Metropolis test (gibbs.cpp:90), then p gamma variates of shape
// (alpha + K)/2 in feature order (spl_sgm_ig, utils.cpp:29-33). With it a fit consumes the host's stream exactly as the
// reference's C++ sampler does, so set.seed() gives the reference's chain to rounding. u is only known once the
// previous iteration's variances exist, so this mode runs one iteration per device call (slow; meant for checks). The
// ARS priors and the logw update draw a data-dependent number of uniforms: t prior with eta = 0 and thin = 1 only.
struct HostRng {
  std::function<double()> norm, unif;
  std::function<double(double)> gamma;   // shape -> Gamma(shape, 1)
};

// Extras the CUDA path needs beyond the 22 reference arguments.
struct DeviceOptions {
  int device = 0;
  uint64_t seed = 0;            // device Philox key (ignored when `inject` is given)
  int algo = HTLR_ALGO_AUTO;
  bool use_graph = true;
  void* nccl_comm = nullptr;    // row-sharded X: communicator from htlr_cuda_nccl_init
  int ars_inject_width = 0;
  const HostRng* host_rng = nullptr;   // draw every variate from the host's stream, in the reference's order
};

class CudaFit {
 public:
  // progress(iterations done, stats of the chunk just finished, chunk length) -> true to abort.
  // The Rcpp glue prints the reference's progress line and polls R_CheckUserInterrupt from it.
  using Progress = std::function<bool(int, const htlr_iter_stats*, int)>;

  CudaFit(int p, int K, int n, const double* X, const double* ymat, const int64_t* ybase, const std::string& ptype,
          double alpha, double s, double eta, int iters_rmc, int iters_h, int thin, int leap_L, int leap_L_h,
          double leap_step, double hmc_sgmcut, const double* deltas, const double* sigmasbt, bool keep_warmup_hist,
          int silence, bool legacy, const DeviceOptions& opt = DeviceOptions(), const htlr_inject* inject = nullptr)
      : inject_(inject), host_rng_(opt.host_rng) {
    cfg_ = htlr_cfg{};
    cfg_.struct_size = sizeof(htlr_cfg);
    cfg_.abi_version = HTLR_CUDA_ABI_VERSION;
    cfg_.p = p; cfg_.K = K; cfg_.n = n;
    // Fit::UpdateSigmas (gibbs.cpp:299-309) stops with "Unsupported prior type" at the first sigma update;
    // here the same text is raised before any device work.
    if (ptype == "t") cfg_.ptype = HTLR_PRIOR_T;
    else if (ptype == "ghs") cfg_.ptype = HTLR_PRIOR_GHS;
    else if (ptype == "neg") cfg_.ptype = HTLR_PRIOR_NEG;
    else throw CudaFitError(HTLR_E_ARG, "Unsupported prior type " + ptype);
    cfg_.alpha = alpha; cfg_.s = s; cfg_.eta = eta;
    cfg_.iters_rmc = iters_rmc; cfg_.iters_h = iters_h; cfg_.thin = thin;
    cfg_.leap_L = leap_L; cfg_.leap_L_h = leap_L_h; cfg_.leap_step = leap_step;
    cfg_.hmc_sgmcut = hmc_sgmcut;
    cfg_.keep_warmup_hist = keep_warmup_hist; cfg_.silence = silence; cfg_.legacy = legacy;
    cfg_.device = opt.device;
    cfg_.rng_mode = (inject || host_rng_) ? HTLR_RNG_INJECT : HTLR_RNG_DEVICE;
    if (host_rng_ && (cfg_.ptype != HTLR_PRIOR_T || eta > 1e-10 || thin != 1))
      throw CudaFitError(HTLR_E_ARG, "the host-order RNG mode supports the t prior with eta = 0 and thin = 1 only "
                                     "(ARS and thinning consume a data-dependent number of variates)");
    cfg_.algo = opt.algo; cfg_.seed = opt.seed; cfg_.use_graph = opt.use_graph;
    cfg_.x_on_device = 0; cfg_.nccl_comm = opt.nccl_comm; cfg_.ars_inject_width = opt.ars_inject_width;
    const int rc = htlr_cuda_create(&cfg_, X, n, ymat, ybase, deltas, sigmasbt, &h_);
    if (rc) throw CudaFitError(rc, htlr_cuda_last_error(nullptr));
  }
  CudaFit(const CudaFit&) = delete;
  CudaFit& operator=(const CudaFit&) = delete;
  ~CudaFit() { htlr_cuda_destroy(h_); }

  // Fit::StartSampling (gibbs.cpp:52-126). Device RNG: pipelined — units of 8 iterations, up to 64 in flight,
  // each unit's trace slices copied into the output arrays while the following units run (htlr_cuda_collect);
  // progress() is called per unit, far inside the reference's interrupt-poll period of 256 iterations
  // (gibbs.cpp:124). Injected variates are positional, so they run as one blocking call.
  void StartSampling(const Progress& progress = nullptr) {
    const int total = cfg_.iters_h + cfg_.iters_rmc;
    if (host_rng_) {
      // one iteration per call: the restricted set of iteration i is a function of the variances iteration i-1 left
      const size_t nvar = (size_t)cfg_.p + 1, K = (size_t)cfg_.K;
      const double cut = cfg_.hmc_sgmcut > 0 ? cfg_.hmc_sgmcut * cfg_.hmc_sgmcut : cfg_.hmc_sgmcut;   // gibbs.cpp:14
      const double shape = (cfg_.alpha + cfg_.K) / 2;
      std::vector<double> sig(nvar), normals, gammas((size_t)cfg_.p);
      htlr_iter_stats st{};
      for (int done = 0; done < total; ++done) {
        check(htlr_cuda_get_state(h_, nullptr, nullptr, sig.data(), nullptr, nullptr, nullptr));
        size_t u = 0;
        for (size_t j = 0; j < nvar; ++j) u += sig[j] > cut;            // WhichUpdate, gibbs.cpp:156-170
        normals.resize(u * K);
        for (double& z : normals) z = host_rng_->norm();                  // GenMomt: j-major, k inner
        double unif = host_rng_->unif();                                  // the accept test
        for (double& gq : gammas) gq = host_rng_->gamma(shape);           // UpdateSigmasT, features 1..p
        htlr_inject inj{};
        inj.normals = normals.data(); inj.n_normals = (int64_t)normals.size();
        inj.uniforms = &unif; inj.n_uniforms = 1;
        inj.gammas = gammas.data(); inj.n_gammas = (int64_t)gammas.size();
        check(htlr_cuda_run(h_, 1, &inj, &st));
        if (cfg_.silence == 0)
          std::printf("Iter%4d: deviance=%5.3f, logw=%6.2f, nuvar=%3.0f, hmcrej=%4.2f\n", done - cfg_.iters_h,
                      -st.loglike / cfg_.n, st.logw, st.uvar, st.rej);
        if (progress && progress(done + 1, &st, 1)) throw CudaFitError(HTLR_E_STATE, "interrupted");
      }
      return;
    }
    const int unit = inject_ ? total : 8, depth = 64;
    std::vector<htlr_iter_stats> st(unit > 0 ? unit : 1);
    if (!inject_) { size_output(); collected_ = true; }
    int enq = 0;
    for (int done = 0; done < total;) {
      const int upto = total - done < unit ? total : done + unit;
      if (inject_) {
        check(htlr_cuda_run(h_, total, inject_, st.data()));
      } else {
        while (enq < total && enq - done < depth) {
          const int nq = total - enq < unit ? total - enq : unit;
          check(htlr_cuda_run_async(h_, nq));
          enq += nq;
        }
        check(htlr_cuda_collect(h_, upto, done, st.data(), out_.mcdeltas.data(), out_.mcsigmasbt.data(),
                                out_.mcvardeltas.data(), out_.mclogw.data(), out_.mcloglike.data(),
                                out_.mcuvar.data(), out_.mchmcrej.data()));
      }
      const int chunk = upto - done;
      if (cfg_.silence == 0) {
        for (int i = 0; i < chunk; ++i)   // gibbs.cpp:119-121
          std::printf("Iter%4d: deviance=%5.3f, logw=%6.2f, nuvar=%3.0f, hmcrej=%4.2f\n", done + i - cfg_.iters_h,
                      -st[i].loglike / cfg_.n, st[i].logw, st[i].uvar, st[i].rej);
      }
      done = upto;
      if (progress && progress(done, st.data(), chunk)) throw CudaFitError(HTLR_E_STATE, "interrupted");
    }
  }

  // Fit::OutputR (gibbs.cpp:128-152).
  FitOutput Output() {
    size_output();
    FitOutput& o = out_;
    if (collected_) {   // the trace arrived during StartSampling
      check(htlr_cuda_fetch(h_, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, o.DDNloglike.data()));
    } else {
      check(htlr_cuda_fetch(h_, o.mcdeltas.data(), o.mcsigmasbt.data(), o.mcvardeltas.data(), o.mclogw.data(),
                            o.mcloglike.data(), o.mcuvar.data(), o.mchmcrej.data(), o.DDNloglike.data()));
    }
    FitOutput r = std::move(out_);   // a later call downloads everything again
    out_ = FitOutput();
    collected_ = false;
    return r;
  }

  htlr_handle* handle() { return h_; }

 private:
  void size_output() {
    FitOutput& o = out_;
    if (o.hist_len) return;
    const size_t nvar = (size_t)cfg_.p + 1, K = (size_t)cfg_.K, H = (size_t)htlr_cuda_hist_len(h_);
    o.p = cfg_.p; o.n = cfg_.n; o.K = cfg_.K;
    o.iter_rmc = cfg_.iters_rmc; o.iter_warm = cfg_.iters_h; o.thin = cfg_.thin;
    o.leap = cfg_.leap_L; o.leap_warm = cfg_.leap_L_h; o.leap_step = cfg_.leap_step;
    o.sgmsq_cut = cfg_.hmc_sgmcut > 0 ? cfg_.hmc_sgmcut * cfg_.hmc_sgmcut : cfg_.hmc_sgmcut;  // gibbs.cpp:14
    o.hist_len = (int)H;
    o.DDNloglike.resize(nvar);
    o.mcdeltas.resize(nvar * K * H);
    o.mcsigmasbt.resize(nvar * H);
    o.mcvardeltas.resize(nvar * H);
    o.mclogw.resize(H); o.mcloglike.resize(H); o.mcuvar.resize(H); o.mchmcrej.resize(H);
  }
  void check(int rc) {
    if (rc) throw CudaFitError(rc, htlr_cuda_last_error(h_));
  }
  htlr_cfg cfg_;
  htlr_handle* h_ = nullptr;
  const htlr_inject* inject_;
  const HostRng* host_rng_ = nullptr;
  FitOutput out_;
  bool collected_ = false;
};

// The reference's free function (main.cpp:7-25), same 22 arguments in the same order.
inline FitOutput htlr_fit_helper(int p, int K, int n, const double* X, const double* ymat, const int64_t* ybase,
                                 const std::string& ptype, double alpha, double s, double eta, int iters_rmc,
                                 int iters_h, int thin, int leap_L, int leap_L_h, double leap_step, double hmc_sgmcut,
                                 const double* deltas, const double* sigmasbt, bool keep_warmup_hist, int silence,
                                 bool legacy, const DeviceOptions& opt = DeviceOptions(),
                                 const htlr_inject* inject = nullptr, const CudaFit::Progress& progress = nullptr) {
  CudaFit f(p, K, n, X, ymat, ybase, ptype, alpha, s, eta, iters_rmc, iters_h, thin, leap_L, leap_L_h, leap_step,
            hmc_sgmcut, deltas, sigmasbt, keep_warmup_hist, silence, legacy, opt, inject);
  f.StartSampling(progress);
  return f.Output();
}

}  // namespace htlr
