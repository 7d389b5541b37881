// TEST INFRASTRUCTURE — CPU oracle for the HTLR restricted-Gibbs / HMC sampler.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// It is never the thing shipped or measured as product.
//
// A plain-C++ restatement of the reference's `Fit` class, statement by statement:
//   ctor                 /root/reference/src/gibbs.cpp:4-50
//   StartSampling        gibbs.cpp:52-126        (order of one MCMC iteration; see `iterate`)
//   WhichUpdate          gibbs.cpp:156-170
//   UpdatePredProb       gibbs.cpp:176-191  + find_normlv / log_sum_exp(mat)  utils.cpp:10-26
//   DetachFixlv          gibbs.cpp:197-229
//   UpdateDNlogLike      gibbs.cpp:236-250
//   UpdateLogLike        gibbs.cpp:254-261
//   UpdateDNlogPrior     gibbs.cpp:267-272
//   UpdateDNlogPost      gibbs.cpp:279-283
//   UpdateMomtAndDeltas  gibbs.cpp:291-297
//   UpdateSigmas*        gibbs.cpp:299-388  + spl_sgm_ig utils.cpp:29-33, targets sampler.cpp:18-72
//   Traject              gibbs.cpp:390-419
//   UpdateVarDeltas      gibbs.cpp:425-429
//   CompNegEnergy        gibbs.cpp:432-437
//   GenMomt / MoveMomt / UpdateStepSizes / Cache / Restore / IsFault / Initialize   gibbs.cpp:441-530
// Loop nests are j-outer / k / i-inner with sequential sums exactly as gibbs.cpp:179-188, 203-227, 239-249.
// Arrays are column-major like Armadillo's.
//
// Known, documented deviations (DESIGN.md "oracle"):
//  * Armadillo's arma::sum / arma::accu over u- or p-length vectors (gibbs.cpp:434-435, sampler.cpp:63-64)
//    use an unrolled two-accumulator order in real Armadillo; here they are sequential. Affects the
//    energy by O(u*eps).
//  * Random variates come from a pluggable source (keyed Philox or injected arrays) instead of R's
//    Mersenne-Twister stream; consumption ORDER per thin-step is the reference's
//    (u*K normals j-major, 1 uniform, then p gammas or per-feature ARS uniforms).
//
// Pinning: checked against the reference's own sources compiled verbatim (oracle/_ref: ars.cpp against
// an Rcpp shim; gibbs.cpp/sampler.cpp/utils.cpp/main.cpp against an Armadillo-subset shim) by
// tests/test_ref_pin.py, and against fixtures generated from that build (tests/golden/).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "ars_port.h"
#include "rng_ref.h"

namespace {

struct Cfg {
  int32_t p, K, n;
  int32_t ptype;  // 0 "t", 1 "ghs", 2 "neg"
  double alpha, s, eta;
  int32_t iters_rmc, iters_h, thin, leap_L, leap_L_h;
  double leap_step, hmc_sgmcut;
  int32_t keep_warmup_hist, legacy;
  int32_t ars_max_nhull;  // 1000 in the reference (ars.h:61)
};

// Variates in reference consumption order. KEYED: pure functions of (seed, step, j, k).
// INJECT: sequential cursors over caller-provided arrays (normals, accept uniforms, gammas) and a
// fixed-width per-(step, feature) block of ARS uniforms; with no ARS block array the ARS uniforms are
// taken from the same sequential uniform stream as the accept test — the reference's exact order.
struct Rand : ars_port::Uniforms {
  int mode = 0;  // 0 keyed, 1 inject
  oracle_rng::Keyed keyed;
  const double *normals = nullptr, *uniforms = nullptr, *gammas = nullptr, *ars_u = nullptr;
  int64_t n_normals = 0, n_uniforms = 0, n_gammas = 0, ars_width = 0, n_ars_blocks = 0;
  int64_t c_normals = 0, c_uniforms = 0, c_gammas = 0;
  // current ARS stream
  uint64_t ars_step = 0;
  uint32_t ars_j = 0, ars_draw = 0;
  uint32_t ars_purpose = oracle_rng::kArs;
  int64_t ars_block = 0;

  double norm(uint64_t step, int j, int k) {
    if (mode == 0) return keyed.norm(oracle_rng::kMomentum, step, (uint32_t)j, (uint32_t)k);
    if (c_normals >= n_normals) throw std::runtime_error("injected normals exhausted");
    return normals[c_normals++];
  }
  double accept_unif(uint64_t step) {
    if (mode == 0) return keyed.unif(oracle_rng::kAccept, step, 0, 0);
    if (c_uniforms >= n_uniforms) throw std::runtime_error("injected uniforms exhausted");
    return uniforms[c_uniforms++];
  }
  double gamma(uint64_t step, int j, double shape) {
    if (mode == 0) return keyed.gamma(step, (uint32_t)j, shape);
    if (c_gammas >= n_gammas) throw std::runtime_error("injected gammas exhausted");
    return gammas[c_gammas++];
  }
  void ars_begin(uint64_t step, int j, uint32_t purpose, int64_t block) {
    ars_step = step;
    ars_j = (uint32_t)j;
    ars_draw = 0;
    ars_purpose = purpose;
    ars_block = block;
  }
  double next() override {
    uint32_t d = ars_draw++;
    if (mode == 0) return keyed.unif(ars_purpose, ars_step, ars_j, d);
    if (!ars_u) {  // single sequential uniform stream shared with the accept test (reference order)
      if (c_uniforms >= n_uniforms) throw std::runtime_error("injected uniforms exhausted");
      return uniforms[c_uniforms++];
    }
    if (ars_block >= n_ars_blocks || (int64_t)d >= ars_width)
      throw std::runtime_error("injected ARS uniforms exhausted");
    return ars_u[ars_block * ars_width + d];
  }
};

struct Fit {
  Cfg c;
  int C, nvar;
  double sgmsq_cut;
  std::vector<double> X, ymat;  // n x nvar, n x K (column-major)
  std::vector<int64_t> ybase;
  std::vector<double> DDNloglike;

  // traces
  int hist_len;
  std::vector<double> mc_deltas, mc_sigmasbt, mc_var_deltas, mc_logw, mc_loglike, mc_uvar, mc_hmcrej;

  // state
  double logw;
  int nuvar = 0, nfvar = 0;
  std::vector<int> ids_update, ids_fix;
  std::vector<double> lv, lv_old, lv_fix, norm_lv, pred_prob, pred_prob_old, DNloglike, deltas,
      deltas_old, momt, DNlogprior, DNlogprior_old, DNlogpost, sumsq_deltas, sum_deltas, var_deltas,
      var_deltas_old, step_sizes, sigmasbt;
  double loglike = 0, loglike_old = 0;

  Rand rnd;
  int i_mc = 0;         // next outer iteration
  uint64_t step = 0;    // global thin-step counter
  // per outer iteration stats (length iters_h + iters_rmc)
  std::vector<double> it_loglike, it_logw, it_uvar, it_rej;
  // ARS effort counters
  int64_t ars_evals = 0, ars_unifs = 0, ars_draws = 0;
  int ars_max_hulls = 0;

  double& Xe(int i, int j) { return X[(size_t)j * c.n + i]; }

  Fit(const Cfg& cfg, const double* Xp, const double* ymatp, const int64_t* ybasep,
      const double* deltasp, const double* sigmasbtp)
      : c(cfg), C(cfg.K + 1), nvar(cfg.p + 1) {
    const int n = c.n, K = c.K;
    if (c.ptype < 0 || c.ptype > 2) throw std::runtime_error("Unsupported prior type");
    X.assign(Xp, Xp + (size_t)n * nvar);
    ymat.assign(ymatp, ymatp + (size_t)n * K);
    ybase.assign(ybasep, ybasep + n);
    for (int i = 0; i < n; ++i)
      if (ybase[i] < 0 || ybase[i] >= C) throw std::runtime_error("ybase out of range [0, C)");
    sgmsq_cut = c.hmc_sgmcut > 0 ? c.hmc_sgmcut * c.hmc_sgmcut : c.hmc_sgmcut;  // gibbs.cpp:14
    DDNloglike.assign(nvar, 0.0);                                                // gibbs.cpp:15
    for (int j = 0; j < nvar; ++j) {
      double a = 0;
      for (int i = 0; i < n; ++i) a += Xe(i, j) * Xe(i, j);
      DDNloglike[j] = a / 4;
    }
    logw = c.s;
    sigmasbt.assign(sigmasbtp, sigmasbtp + nvar);
    ids_update.assign(nvar, 0);
    ids_fix.assign(nvar, 0);
    hist_len = c.keep_warmup_hist ? (c.iters_rmc + c.iters_h + 1) : (c.iters_rmc + 1);
    mc_logw.assign(hist_len, 0.0);
    mc_logw[0] = logw;
    mc_sigmasbt.assign((size_t)nvar * hist_len, 0.0);
    std::copy(sigmasbt.begin(), sigmasbt.end(), mc_sigmasbt.begin());
    deltas.assign(deltasp, deltasp + (size_t)nvar * K);
    mc_deltas.assign((size_t)nvar * K * hist_len, 0.0);
    std::copy(deltas.begin(), deltas.end(), mc_deltas.begin());
    mc_var_deltas.assign((size_t)nvar * hist_len, 0.0);
    mc_loglike.assign(hist_len, 0.0);
    mc_uvar.assign(hist_len, 0.0);
    mc_hmcrej.assign(hist_len, 0.0);
    lv.assign((size_t)n * C, 0.0);
    lv_fix.assign((size_t)n * C, 0.0);
    norm_lv.assign((size_t)n * C, 0.0);
    pred_prob.assign((size_t)n * C, 0.0);
    DNloglike.assign((size_t)nvar * K, 0.0);
    momt.assign((size_t)nvar * K, 0.0);
    DNlogprior.assign((size_t)nvar * K, 0.0);
    DNlogpost.assign((size_t)nvar * K, 0.0);
    sumsq_deltas.assign(nvar, 0.0);
    sum_deltas.assign(nvar, 0.0);
    var_deltas.assign(nvar, 0.0);
    step_sizes.assign(nvar, 0.0);
    int tot = c.iters_h + c.iters_rmc;
    it_loglike.assign(tot, 0.0);
    it_logw.assign(tot, 0.0);
    it_uvar.assign(tot, 0.0);
    it_rej.assign(tot, 0.0);
  }

  // gibbs.cpp:156-170
  void WhichUpdate(bool init = false) {
    nuvar = 0;
    nfvar = 0;
    double cut = init ? -1 : sgmsq_cut;
    for (int j = 0; j < nvar; ++j) {
      if (sigmasbt[j] > cut) ids_update[nuvar++] = j; else ids_fix[nfvar++] = j;
    }
  }

  // gibbs.cpp:176-191, utils.cpp:10-26
  void UpdatePredProb() {
    const int n = c.n, K = c.K;
    std::copy(lv_fix.begin() + n, lv_fix.end(), lv.begin() + n);
    for (int q = 0; q < nuvar; ++q) {
      int j = ids_update[q];
      for (int k = 0; k < K; ++k) {
        double d = deltas[(size_t)k * nvar + j];
        double* lvk = &lv[(size_t)(k + 1) * n];
        const double* xj = &X[(size_t)j * n];
        for (int i = 0; i < n; ++i) lvk[i] += xj[i] * d;
      }
    }
    for (int i = 0; i < n; ++i) {
      double m = lv[i];
      for (int cc = 1; cc < C; ++cc) m = std::max(m, lv[(size_t)cc * n + i]);
      double sm = 0;
      for (int cc = 0; cc < C; ++cc) sm += std::exp(lv[(size_t)cc * n + i] - m);
      double lse = std::log(sm) + m;
      for (int cc = 0; cc < C; ++cc) {
        double nl = lv[(size_t)cc * n + i] - lse;
        norm_lv[(size_t)cc * n + i] = nl;
        pred_prob[(size_t)cc * n + i] = std::exp(nl);
      }
    }
  }

  // gibbs.cpp:197-229
  void DetachFixlv() {
    const int n = c.n, K = c.K;
    if (nuvar <= nvar / 2) {
      std::copy(lv.begin() + n, lv.end(), lv_fix.begin() + n);
      for (int q = 0; q < nuvar; ++q) {
        int j = ids_update[q];
        for (int k = 0; k < K; ++k) {
          double d = deltas[(size_t)k * nvar + j];
          double* f = &lv_fix[(size_t)(k + 1) * n];
          const double* xj = &X[(size_t)j * n];
          for (int i = 0; i < n; ++i) f[i] -= xj[i] * d;
        }
      }
    } else {
      std::fill(lv_fix.begin() + n, lv_fix.end(), 0.0);
      for (int q = 0; q < nfvar; ++q) {
        int j = ids_fix[q];
        for (int k = 0; k < K; ++k) {
          double d = deltas[(size_t)k * nvar + j];
          double* f = &lv_fix[(size_t)(k + 1) * n];
          const double* xj = &X[(size_t)j * n];
          for (int i = 0; i < n; ++i) f[i] += xj[i] * d;
        }
      }
    }
  }

  // gibbs.cpp:236-250
  void UpdateDNlogLike() {
    const int n = c.n, K = c.K;
    std::vector<double> tmp((size_t)n * K);
    for (int k = 0; k < K; ++k)
      for (int i = 0; i < n; ++i)
        tmp[(size_t)k * n + i] = pred_prob[(size_t)(k + 1) * n + i] - ymat[(size_t)k * n + i];
    for (int q = 0; q < nuvar; ++q) {
      int j = ids_update[q];
      const double* xj = &X[(size_t)j * n];
      for (int k = 0; k < K; ++k) {
        double a = 0;
        const double* t = &tmp[(size_t)k * n];
        for (int i = 0; i < n; ++i) a += xj[i] * t[i];
        DNloglike[(size_t)k * nvar + j] = a;
      }
    }
  }

  // gibbs.cpp:254-261
  void UpdateLogLike() {
    loglike = 0;
    for (int i = 0; i < c.n; ++i) loglike += norm_lv[(size_t)ybase[i] * c.n + i];
  }

  // gibbs.cpp:267-272
  void UpdateDNlogPrior() {
    const int K = c.K;
    for (int q = 0; q < nuvar; ++q) {
      int j = ids_update[q];
      double sd = deltas[j];
      for (int k = 1; k < K; ++k) sd += deltas[(size_t)k * nvar + j];
      sum_deltas[j] = sd;
    }
    for (int k = 0; k < K; ++k)
      for (int q = 0; q < nuvar; ++q) {
        int j = ids_update[q];
        DNlogprior[(size_t)k * nvar + j] = deltas[(size_t)k * nvar + j] - sum_deltas[j] / C;
      }
  }

  // gibbs.cpp:279-283
  void UpdateDNlogPost() {
    for (int k = 0; k < c.K; ++k)
      for (int q = 0; q < nuvar; ++q) {
        int j = ids_update[q];
        size_t e = (size_t)k * nvar + j;
        DNlogpost[e] = DNloglike[e] + DNlogprior[e] / sigmasbt[j];
      }
  }

  // gibbs.cpp:291-297
  void UpdateMomtAndDeltas() {
    for (int k = 0; k < c.K; ++k)
      for (int q = 0; q < nuvar; ++q) {
        int j = ids_update[q];
        size_t e = (size_t)k * nvar + j;
        momt[e] -= step_sizes[j] / 2 * DNlogpost[e];
      }
    for (int k = 0; k < c.K; ++k)
      for (int q = 0; q < nuvar; ++q) {
        int j = ids_update[q];
        size_t e = (size_t)k * nvar + j;
        deltas[e] += step_sizes[j] * momt[e];
      }
  }

  // gibbs.cpp:467-471
  void MoveMomt() {
    for (int k = 0; k < c.K; ++k)
      for (int q = 0; q < nuvar; ++q) {
        int j = ids_update[q];
        size_t e = (size_t)k * nvar + j;
        momt[e] -= step_sizes[j] / 2 * DNlogpost[e];
      }
  }

  // gibbs.cpp:477-481
  void UpdateStepSizes() {
    for (int q = 0; q < nuvar; ++q) {
      int j = ids_update[q];
      step_sizes[j] = c.leap_step / std::sqrt(DDNloglike[j] + c.K / sigmasbt[j] / C);
    }
  }

  // gibbs.cpp:425-429
  void UpdateVarDeltas() {
    const int K = c.K;
    for (int q = 0; q < nuvar; ++q) {
      int j = ids_update[q];
      double ss = deltas[j] * deltas[j];
      for (int k = 1; k < K; ++k) ss += deltas[(size_t)k * nvar + j] * deltas[(size_t)k * nvar + j];
      sumsq_deltas[j] = ss;
      var_deltas[j] = ss - (sum_deltas[j] * sum_deltas[j]) / C;
    }
  }

  // gibbs.cpp:432-437
  double CompNegEnergy() {
    double logprior = 0;
    for (int q = 0; q < nuvar; ++q) {
      int j = ids_update[q];
      logprior += var_deltas[j] / sigmasbt[j];
    }
    double pm = 0;
    for (int k = 0; k < c.K; ++k)
      for (int q = 0; q < nuvar; ++q) {
        double m = momt[(size_t)k * nvar + ids_update[q]];
        pm += m * m;
      }
    return loglike - logprior / 2 - pm / 2;
  }

  // gibbs.cpp:441-460
  void GenMomt() {
    for (int q = 0; q < nuvar; ++q) {
      int j = ids_update[q];
      for (int k = 0; k < c.K; ++k) momt[(size_t)k * nvar + j] = rnd.norm(step, j, k);
    }
  }

  // gibbs.cpp:483-501
  void CacheOldValues() {
    lv_old = lv;
    pred_prob_old = pred_prob;
    deltas_old = deltas;
    DNlogprior_old = DNlogprior;
    var_deltas_old = var_deltas;
    loglike_old = loglike;
  }
  void RestoreOldValues() {
    lv = lv_old;
    pred_prob = pred_prob_old;
    deltas = deltas_old;
    DNlogprior = DNlogprior_old;
    var_deltas = var_deltas_old;
    loglike = loglike_old;
  }

  // gibbs.cpp:503-516
  bool IsFault(double cri = 20) {
    for (int q = 0; q < nuvar; ++q)
      for (int k = 0; k < c.K; ++k)
        if (std::fabs(deltas[(size_t)k * nvar + ids_update[q]]) > cri) return true;
    return false;
  }

  // gibbs.cpp:390-419; the trajectory body is exposed separately for the deterministic-piece tests
  int PickL(int imc) {
    if (imc < c.iters_h / 2.0) {
      logw = -10;
      return c.leap_L_h;
    } else if (imc < c.iters_h) {
      logw = c.s;
      return c.leap_L_h;
    }
    logw = c.s;
    return c.leap_L;
  }
  void Leapfrog(int L) {
    for (int t = 0; t < L; ++t) {
      UpdateMomtAndDeltas();
      UpdatePredProb();
      UpdateDNlogPrior();
      UpdateDNlogLike();
      UpdateDNlogPost();
      MoveMomt();
    }
  }

  // gibbs.cpp:311-331, utils.cpp:29-33
  void UpdateSigmasT() {
    double alpha_post = (c.alpha + c.K) / 2;
    double w = std::exp(logw);
    for (int j = 1; j < nvar; ++j) {
      double g = rnd.gamma(step, j, alpha_post);
      if (c.legacy)
        sigmasbt[j] = 1.0 / g * (c.alpha * w + var_deltas[j]) / 2.0;  // gibbs.cpp:320
      else
        sigmasbt[j] = (1.0 / g * (c.alpha * w + var_deltas[j]) / 2.0);  // utils.cpp:32
    }
    UpdateLogw();
  }

  // gibbs.cpp:333-348
  void UpdateLogw() {
    if (c.eta > 1E-10) {
      if (c.eta < 0.01) {
        logw = c.s;
      } else {
        ars_port::LogwTarget tg;
        tg.vard = &var_deltas[1];
        tg.p = c.p;
        tg.K = c.K;
        tg.nu = c.alpha;
        tg.s = c.s;
        tg.eta = c.eta;
        rnd.ars_begin(step, 0, oracle_rng::kLogwArs, (int64_t)step * (c.p + 1) + c.p);
        ars_port::Sampler spl(&tg, &rnd, logw, c.ars_max_nhull);
        logw = spl.draw();
        note_ars(spl);
      }
    }
  }

  void note_ars(const ars_port::Sampler& s) {
    ars_evals += s.stats().n_eval;
    ars_unifs += s.stats().n_unif;
    ars_draws += 1;
    ars_max_hulls = std::max(ars_max_hulls, s.stats().n_hulls);
  }

  // gibbs.cpp:351-388 (legacy and non-legacy branches are arithmetically identical)
  void UpdateSigmasSgm() {
    double log_aw = logw + std::log(c.alpha);
    ars_port::NegTarget neg;
    ars_port::GhsTarget ghs;
    neg.K = ghs.K = c.K;
    neg.alpha = ghs.alpha = c.alpha;
    neg.log_aw = ghs.log_aw = log_aw;
    for (int j = 1; j < nvar; ++j) {
      ars_port::Target* tg;
      if (c.ptype == 2) { neg.v = var_deltas[j]; tg = &neg; } else { ghs.v = var_deltas[j]; tg = &ghs; }
      rnd.ars_begin(step, j, oracle_rng::kArs, (int64_t)step * (c.p + 1) + (j - 1));
      ars_port::Sampler spl(tg, &rnd, std::log(var_deltas[j] / c.K), c.ars_max_nhull);
      sigmasbt[j] = std::exp(spl.draw());
      note_ars(spl);
    }
  }

  void UpdateSigmas() {
    if (c.ptype == 0) UpdateSigmasT(); else UpdateSigmasSgm();
  }

  // gibbs.cpp:519-530
  void Initialize() {
    WhichUpdate(true);
    UpdatePredProb();
    UpdateLogLike();
    mc_loglike[0] = loglike;
    UpdateDNlogPrior();
    UpdateVarDeltas();
    std::copy(var_deltas.begin(), var_deltas.end(), mc_var_deltas.begin());
  }

  // one outer iteration, gibbs.cpp:57-125
  void iterate() {
    const int K = c.K;
    double no_uvar = 0, rej = 0;
    for (int i_thin = 0; i_thin < c.thin; ++i_thin) {
      WhichUpdate();
      no_uvar += nuvar;
      GenMomt();
      UpdateStepSizes();
      DetachFixlv();
      CacheOldValues();
      double nenergy_old = CompNegEnergy();
      UpdateDNlogPrior();
      UpdateDNlogLike();
      UpdateDNlogPost();
      int L = PickL(i_mc);
      Leapfrog(L);
      UpdateLogLike();
      UpdateVarDeltas();
      double nenergy = CompNegEnergy();
      double u = rnd.accept_unif(step);
      if (std::log(u) > (nenergy - nenergy_old) || IsFault()) {
        RestoreOldValues();
        rej++;
      }
      UpdateSigmas();
      ++step;
    }
    no_uvar /= c.thin;
    rej /= c.thin;
    int i_rmc = c.keep_warmup_hist ? (i_mc + 1) : (i_mc - c.iters_h + 1);
    if (i_rmc > 0) {
      std::copy(deltas.begin(), deltas.end(), mc_deltas.begin() + (size_t)i_rmc * nvar * K);
      std::copy(sigmasbt.begin(), sigmasbt.end(), mc_sigmasbt.begin() + (size_t)i_rmc * nvar);
      std::copy(var_deltas.begin(), var_deltas.end(), mc_var_deltas.begin() + (size_t)i_rmc * nvar);
      mc_logw[i_rmc] = logw;
      mc_loglike[i_rmc] = loglike;
      mc_uvar[i_rmc] = no_uvar;
      mc_hmcrej[i_rmc] = rej;
    }
    it_loglike[i_mc] = loglike;
    it_logw[i_mc] = logw;
    it_uvar[i_mc] = no_uvar;
    it_rej[i_mc] = rej;
    ++i_mc;
  }
};

thread_local std::string g_err;

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  } catch (...) {
    g_err = "unknown error";
    return 2;
  }
}

struct ArrUniforms : ars_port::Uniforms {
  const double* u;
  int n, c = 0;
  double next() override {
    if (c >= n) throw std::runtime_error("ARS uniforms exhausted");
    return u[c++];
  }
};

}  // namespace

extern "C" {

struct oracle_cfg {
  int32_t p, K, n, ptype;
  double alpha, s, eta;
  int32_t iters_rmc, iters_h, thin, leap_L, leap_L_h;
  double leap_step, hmc_sgmcut;
  int32_t keep_warmup_hist, legacy, ars_max_nhull, pad_;
};

const char* oracle_last_error() { return g_err.c_str(); }

int oracle_create(const oracle_cfg* oc, const double* X, const double* ymat, const int64_t* ybase,
                  const double* deltas, const double* sigmasbt, void** out) {
  return guarded([&] {
    Cfg c{oc->p, oc->K, oc->n, oc->ptype, oc->alpha, oc->s, oc->eta, oc->iters_rmc, oc->iters_h,
          oc->thin, oc->leap_L, oc->leap_L_h, oc->leap_step, oc->hmc_sgmcut, oc->keep_warmup_hist,
          oc->legacy, oc->ars_max_nhull > 0 ? oc->ars_max_nhull : 1000};
    auto* f = new Fit(c, X, ymat, ybase, deltas, sigmasbt);
    *out = f;
  });
}

void oracle_destroy(void* h) { delete static_cast<Fit*>(h); }

void oracle_set_seed(void* h, uint64_t seed) {
  auto* f = static_cast<Fit*>(h);
  f->rnd.mode = 0;
  f->rnd.keyed.seed = seed;
}

// Pointers must stay valid while the oracle runs. Cursors restart at 0.
void oracle_set_inject(void* h, const double* normals, int64_t n_normals, const double* uniforms,
                       int64_t n_uniforms, const double* gammas, int64_t n_gammas,
                       const double* ars_u, int64_t ars_width, int64_t n_ars_blocks) {
  auto* f = static_cast<Fit*>(h);
  Rand& r = f->rnd;
  r.mode = 1;
  r.normals = normals; r.n_normals = n_normals; r.c_normals = 0;
  r.uniforms = uniforms; r.n_uniforms = n_uniforms; r.c_uniforms = 0;
  r.gammas = gammas; r.n_gammas = n_gammas; r.c_gammas = 0;
  r.ars_u = ars_u; r.ars_width = ars_width; r.n_ars_blocks = n_ars_blocks;
}

int oracle_initialize(void* h) {
  return guarded([&] { static_cast<Fit*>(h)->Initialize(); });
}

// Runs the next n outer iterations (Initialize must have been called). Returns seconds in *secs.
int oracle_run(void* h, int n_iters, double* secs) {
  return guarded([&] {
    auto* f = static_cast<Fit*>(h);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n_iters; ++i) {
      if (f->i_mc >= f->c.iters_h + f->c.iters_rmc) throw std::runtime_error("chain already finished");
      f->iterate();
    }
    if (secs) *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  });
}

int oracle_hist_len(void* h) { return static_cast<Fit*>(h)->hist_len; }

void oracle_fetch(void* h, double* mcdeltas, double* mcsigmasbt, double* mcvardeltas, double* mclogw,
                  double* mcloglike, double* mcuvar, double* mchmcrej, double* ddnloglike) {
  auto* f = static_cast<Fit*>(h);
  auto cp = [](const std::vector<double>& v, double* d) { if (d) std::copy(v.begin(), v.end(), d); };
  cp(f->mc_deltas, mcdeltas);
  cp(f->mc_sigmasbt, mcsigmasbt);
  cp(f->mc_var_deltas, mcvardeltas);
  cp(f->mc_logw, mclogw);
  cp(f->mc_loglike, mcloglike);
  cp(f->mc_uvar, mcuvar);
  cp(f->mc_hmcrej, mchmcrej);
  cp(f->DDNloglike, ddnloglike);
}

void oracle_iter_stats(void* h, double* loglike, double* logw, double* uvar, double* rej) {
  auto* f = static_cast<Fit*>(h);
  auto cp = [](const std::vector<double>& v, double* d) { if (d) std::copy(v.begin(), v.end(), d); };
  cp(f->it_loglike, loglike);
  cp(f->it_logw, logw);
  cp(f->it_uvar, uvar);
  cp(f->it_rej, rej);
}

// Current state snapshot (any pointer may be null). lv/pred are n x C; the rest nvar x K or nvar.
void oracle_state(void* h, double* lv, double* pred, double* deltas, double* sigmasbt, double* var_deltas,
                  double* momt, double* dnloglike, double* dnlogpost, double* loglike, double* logw,
                  int32_t* nuvar, int32_t* iup) {
  auto* f = static_cast<Fit*>(h);
  auto cp = [](const std::vector<double>& v, double* d) { if (d) std::copy(v.begin(), v.end(), d); };
  cp(f->lv, lv);
  cp(f->pred_prob, pred);
  cp(f->deltas, deltas);
  cp(f->sigmasbt, sigmasbt);
  cp(f->var_deltas, var_deltas);
  cp(f->momt, momt);
  cp(f->DNloglike, dnloglike);
  cp(f->DNlogpost, dnlogpost);
  if (loglike) *loglike = f->loglike;
  if (logw) *logw = f->logw;
  if (nuvar) *nuvar = f->nuvar;
  if (iup) std::copy(f->ids_update.begin(), f->ids_update.begin() + f->nuvar, iup);
}

void oracle_ars_effort(void* h, int64_t* evals, int64_t* unifs, int64_t* draws, int32_t* max_hulls) {
  auto* f = static_cast<Fit*>(h);
  *evals = f->ars_evals;
  *unifs = f->ars_unifs;
  *draws = f->ars_draws;
  *max_hulls = f->ars_max_hulls;
}

// ---- deterministic pieces (north_star level-1 parity gate) -------------------------------------------

// lv = X[:,iup] deltas[iup] (lv_fix = 0), softmax, loglike, gradient DNloglike over iup.
// Mirrors UpdatePredProb + UpdateLogLike + UpdateDNlogLike with the given restricted set.
int oracle_eval(void* h, const int32_t* iup, int32_t u, const double* deltas, double* lv, double* pred,
                double* loglike, double* grad /* u x K, row q = iup[q] */) {
  return guarded([&] {
    auto* f = static_cast<Fit*>(h);
    Fit g = *f;  // work on a copy; the chain state is untouched
    g.nuvar = u;
    std::copy(iup, iup + u, g.ids_update.begin());
    std::copy(deltas, deltas + (size_t)g.nvar * g.c.K, g.deltas.begin());
    std::fill(g.lv_fix.begin(), g.lv_fix.end(), 0.0);
    g.UpdatePredProb();
    g.UpdateLogLike();
    g.UpdateDNlogLike();
    if (lv) std::copy(g.lv.begin(), g.lv.end(), lv);
    if (pred) std::copy(g.pred_prob.begin(), g.pred_prob.end(), pred);
    if (loglike) *loglike = g.loglike;
    if (grad)
      for (int k = 0; k < g.c.K; ++k)
        for (int q = 0; q < u; ++q) grad[(size_t)k * u + q] = g.DNloglike[(size_t)k * g.nvar + iup[q]];
  });
}

// One HMC proposal from the CURRENT chain state with caller-supplied momenta (nvar x K, rows outside
// the restricted set ignored) and L leapfrog steps; no accept/reject, chain state untouched.
// Outputs the proposal: deltas, momt (nvar x K), lv (n x C), loglike, var_deltas (nvar), and the two
// energies of gibbs.cpp:76,87.
int oracle_trajectory(void* h, const double* momt_in, int32_t L, double logw_for_run, double* deltas_out,
                      double* momt_out, double* lv_out, double* loglike_out, double* var_deltas_out,
                      double* nenergy_old, double* nenergy_new, int32_t* nuvar_out) {
  return guarded([&] {
    auto* f = static_cast<Fit*>(h);
    Fit g = *f;
    (void)logw_for_run;
    g.WhichUpdate();
    for (int q = 0; q < g.nuvar; ++q) {
      int j = g.ids_update[q];
      for (int k = 0; k < g.c.K; ++k) g.momt[(size_t)k * g.nvar + j] = momt_in[(size_t)k * g.nvar + j];
    }
    g.UpdateStepSizes();
    g.DetachFixlv();
    g.CacheOldValues();
    double e0 = g.CompNegEnergy();
    g.UpdateDNlogPrior();
    g.UpdateDNlogLike();
    g.UpdateDNlogPost();
    g.Leapfrog(L);
    g.UpdateLogLike();
    g.UpdateVarDeltas();
    double e1 = g.CompNegEnergy();
    auto cp = [](const std::vector<double>& v, double* d) { if (d) std::copy(v.begin(), v.end(), d); };
    cp(g.deltas, deltas_out);
    cp(g.momt, momt_out);
    cp(g.lv, lv_out);
    cp(g.var_deltas, var_deltas_out);
    if (loglike_out) *loglike_out = g.loglike;
    if (nenergy_old) *nenergy_old = e0;
    if (nenergy_new) *nenergy_new = e1;
    if (nuvar_out) *nuvar_out = g.nuvar;
  });
}

// One ARS draw per feature with per-feature uniform blocks u[j*width .. (j+1)*width).
// kind: 1 ghs, 2 neg. Outputs x (log sigma^2), number of uniforms consumed, target evaluations, hulls.
int oracle_ars_batch(int32_t kind, int32_t m, const double* vard, double K, double alpha, double log_aw,
                     const double* u, int32_t width, int32_t max_nhull, double* x_out, int32_t* n_unif,
                     int32_t* n_eval, int32_t* n_hulls) {
  return guarded([&] {
    for (int j = 0; j < m; ++j) {
      ars_port::NegTarget neg;
      ars_port::GhsTarget ghs;
      ars_port::Target* tg;
      if (kind == 2) { neg.v = vard[j]; neg.K = K; neg.alpha = alpha; neg.log_aw = log_aw; tg = &neg; }
      else { ghs.v = vard[j]; ghs.K = K; ghs.alpha = alpha; ghs.log_aw = log_aw; tg = &ghs; }
      ArrUniforms au;
      au.u = u + (size_t)j * width;
      au.n = width;
      ars_port::Sampler spl(tg, &au, std::log(vard[j] / K), max_nhull > 0 ? max_nhull : 1000);
      x_out[j] = spl.draw();
      if (n_unif) n_unif[j] = spl.stats().n_unif;
      if (n_eval) n_eval[j] = spl.stats().n_eval;
      if (n_hulls) n_hulls[j] = spl.stats().n_hulls;
    }
  });
}

// logw ARS draw (sampler.cpp:57-72 target) from start point x0.
int oracle_ars_logw(int32_t p, const double* vard, double K, double nu, double s, double eta, double x0,
                    const double* u, int32_t width, double* x_out, int32_t* n_unif, int32_t* n_eval) {
  return guarded([&] {
    ars_port::LogwTarget tg;
    tg.vard = vard; tg.p = p; tg.K = K; tg.nu = nu; tg.s = s; tg.eta = eta;
    ArrUniforms au;
    au.u = u; au.n = width;
    ars_port::Sampler spl(&tg, &au, x0, 1000);
    *x_out = spl.draw();
    if (n_unif) *n_unif = spl.stats().n_unif;
    if (n_eval) *n_eval = spl.stats().n_eval;
  });
}

// Keyed variates, for cross-checking the device RNG twin.
void oracle_rng_dump(uint64_t seed, uint32_t purpose, uint64_t step, int32_t n_a, int32_t n_b, int32_t kind,
                     double shape, double* out) {
  oracle_rng::Keyed kq;
  kq.seed = seed;
  for (int a = 0; a < n_a; ++a)
    for (int b = 0; b < n_b; ++b) {
      double v;
      if (kind == 0) v = kq.unif(purpose, step, a, b);
      else if (kind == 1) v = kq.norm(purpose, step, a, b);
      else v = kq.gamma(step, a, shape);
      out[(size_t)a * n_b + b] = v;
    }
}

// Formula restatements used by the reference's own unit tests (tests/testthat/test-util.r:20-48):
// comp_vardeltas utils.cpp:51-56, comp_lsl utils.cpp:59-63, log_normcons utils.cpp:66-69.
void oracle_comp_vardeltas(const double* deltas, int32_t rows, int32_t cols, double* out) {
  for (int j = 0; j < rows; ++j) {
    double s = 0, ss = 0;
    for (int k = 0; k < cols; ++k) {
      double d = deltas[(size_t)k * rows + j];
      s += d;
      ss += d * d;
    }
    out[j] = ss - (s * s) / (cols + 1);
  }
}
void oracle_comp_lsl(const double* A, int32_t rows, int32_t cols, double* out) {
  for (int i = 0; i < rows; ++i) {
    double m = 0.0;  // the inserted zero column
    for (int k = 0; k < cols; ++k) m = std::max(m, A[(size_t)k * rows + i]);
    double sm = std::exp(0.0 - m);
    for (int k = 0; k < cols; ++k) sm += std::exp(A[(size_t)k * rows + i] - m);
    out[i] = std::log(sm) + m;
  }
}

}  // extern "C"
