// Per-feature variance ("sigmasbt") updates and the logw update. Compiled with --fmad=false so the
// scalar arithmetic of the samplers rounds like the CPU reference's (no fused multiply-adds).
//
// Reference functions replaced:
//   k_sigma (t prior)    Fit::UpdateSigmasT + spl_sgm_ig            src/gibbs.cpp:311-331, src/utils.cpp:29-33
//   k_sigma (ghs / neg)  Fit::UpdateSigmasSgm/Ghs/Neg + ARS engine  src/gibbs.cpp:351-388, src/sampler.cpp:18-49,
//                                                                   src/ars.cpp (see ars.cuh)
//   k_logw               Fit::UpdateLogw + SamplerLogw              src/gibbs.cpp:333-348, src/sampler.cpp:57-72
#include "ars.cuh"
#include "common.cuh"
#include "launch.h"
#include "rng.cuh"

namespace htlr {

// t prior: one thread per feature. sigma^2_j = (alpha*e^logw + var_deltas_j)/2 / Gamma((alpha+K)/2, 1).
__global__ void __launch_bounds__(256) k_sigma_t(Geom g, Bufs b) {
  Ctl* ctl = b.ctl;
  const int p = ctl->p;
  const double shape = (ctl->alpha + ctl->K) / 2;
  const double w = exp(ctl->logw);
  const KeyedRng rng{ctl->seed};
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < p; q += gridDim.x * blockDim.x) {
    const int j = q + 1;
    double gm;
    if (ctl->rng_mode == 1) {
      const long long at = ctl->cur_gam + q;
      if (at >= ctl->n_gam) { set_err(ctl, kErrInjectGammas, 0); gm = 1.0; }
      else gm = b.inj_gam[at];
    } else {
      gm = rng.gamma(ctl->step, (uint32_t)j, shape);
    }
    b.sigmasbt[j] = (1.0 / gm * (ctl->alpha * w + b.var_deltas[j]) / 2.0);
  }
}

// t prior, eta = 0 (the htlr() default): the three per-feature kernels that end a thin-step in ONE launch —
//   commit  keep the proposal (gather through the inverse map posof) or RestoreOldValues   gibbs.cpp:89-95, 493-501
//   sigma   UpdateSigmasT                                                                    gibbs.cpp:311-331
//   record  trace slice + per-iteration scalars + cursor bookkeeping                         gibbs.cpp:100-114
// Thread <-> feature, so a feature's committed coefficients / variance are still in registers when its sigma is
// drawn and its trace entries are written.
__global__ void __launch_bounds__(256) k_tail_t(Geom g, Bufs b) {
  Ctl* ctl = b.ctl;
  timeline_mark(ctl, 8);
  const int nvar = g.nvar, K = g.K, p = ctl->p;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  const bool accept = ctl->accept != 0;
  const double shape = (ctl->alpha + ctl->K) / 2;
  const double w = exp(ctl->logw);
  const KeyedRng rng{ctl->seed};
  const bool last_thin = (ctl->i_thin == ctl->thin - 1);
  const int i_rmc = ctl->keep_warmup_hist ? (ctl->i_mc + 1) : (ctl->i_mc - ctl->iters_h + 1);
  const bool rec = last_thin && i_rmc > 0;
  double* md = b.mcdeltas + (int64_t)(rec ? i_rmc : 0) * nvar * K;
  double* ms = b.mcsigmasbt + (int64_t)(rec ? i_rmc : 0) * nvar;
  double* mv = b.mcvardeltas + (int64_t)(rec ? i_rmc : 0) * nvar;
  for (int j = tid; j < nvar; j += nt) {
    const int r = b.posof[j];
    double vd = b.var_deltas[j];
    // the gamma draw depends on nothing loaded here: it runs while posof / var_deltas are on their way (the
    // branch on r below would otherwise stall the thread for an L2 round trip first)
    double gm = 1.0;
    if (j >= 1) {
      if (ctl->rng_mode == 1) {
        const long long at = ctl->cur_gam + (j - 1);
        if (at >= ctl->n_gam) { set_err(ctl, kErrInjectGammas, 0); gm = 1.0; }
        else gm = b.inj_gam[at];
      } else {
        gm = rng.gamma(ctl->step, (uint32_t)j, shape);
      }
    }
    if (accept && r >= 0) {
      vd = b.vdc[r];
      b.var_deltas[j] = vd;
      for (int k = 0; k < K; ++k) {
        const double d = b.dc[(int64_t)r * K + k];
        b.deltas[(int64_t)k * nvar + j] = d;
        if (rec) md[(int64_t)k * nvar + j] = d;
      }
    } else if (rec) {
      for (int k = 0; k < K; ++k) md[(int64_t)k * nvar + j] = b.deltas[(int64_t)k * nvar + j];
    }
    double sg = b.sigmasbt[j];
    if (j >= 1) {
      sg = (1.0 / gm * (ctl->alpha * w + vd) / 2.0);
      b.sigmasbt[j] = sg;
    }
    if (rec) { ms[j] = sg; mv[j] = vd; }
  }
  if (!accept) {
    const int64_t tot = (int64_t)K * g.n_s;
    for (int64_t e = tid; e < tot; e += nt) { b.lv[e] = b.lv_old[e]; b.R[e] = b.R_old[e]; }
  }
  if (last_block(&ctl->ticket[3], gridDim.x)) {
    if (threadIdx.x == 0) {
      if (accept) ctl->loglike = ctl->loglike_new;
      // advance inject cursors (reference consumption: u*K normals, 1 uniform, p gammas per thin-step)
      ctl->cur_norm += (long long)ctl->u * K;
      ctl->cur_unif += 1;
      ctl->cur_gam += p;
      ctl->step += 1;
      ctl->step_in_call += 1;
      if (last_thin) {
        const double uv = ctl->uvar_acc / ctl->thin, rj = ctl->rej_acc / ctl->thin;
        if (i_rmc > 0) {
          b.mclogw[i_rmc] = ctl->logw;
          b.mcloglike[i_rmc] = ctl->loglike;
          b.mcuvar[i_rmc] = uv;
          b.mchmcrej[i_rmc] = rj;
        }
        Red4 st;
        st.a = ctl->loglike; st.b = ctl->logw; st.c = uv; st.d = rj;
        b.stats[ctl->i_mc] = st;
        ctl->uvar_acc = 0;
        ctl->rej_acc = 0;
        ctl->i_thin = 0;
        ctl->i_mc += 1;
      } else {
        ctl->i_thin += 1;
      }
      timeline_mark_end(ctl, 9);
    }
  }
}

// ghs / neg: ARS on x = log sigma^2_j from x0 = log(var_deltas_j / K), kArsGroup lanes per feature (two features per
// warp: see ars.cuh).
constexpr int kArsGroup = 16;
// 3 blocks per SM (80 registers): measured 6-7 % faster than 2 blocks at 108 registers; 4 blocks add nothing. The
// engine is bound by the FP64 pipe doing group-uniform scalar work in every lane.
__global__ void __launch_bounds__(256, 3) k_sigma_ars(Geom g, Bufs b) {
  Ctl* ctl = b.ctl;
  const int p = ctl->p;
  const int lane = threadIdx.x & (kArsGroup - 1);
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) / kArsGroup, ngrp = (gridDim.x * blockDim.x) / kArsGroup;
  const double log_aw = ctl->logw + log(ctl->alpha);
  for (int q = grp; q < p; q += ngrp) {
    const int j = q + 1;
    SgmTarget tg{ctl->ptype, b.var_deltas[j], (double)ctl->K, ctl->alpha, log_aw};
    const double x0 = log(tg.v / tg.K);
    double x = 0;
    int status;
    {
      ArsUniforms un{false, nullptr, 0, 0, ArsKeyedUniforms{ctl->seed, ctl->step, kRngArs, (uint32_t)j, 0u}};
      if (ctl->rng_mode == 1) {
        un.inject = true;
        const long long blk = ctl->step_in_call * (long long)(p + 1) + q;
        un.u = b.inj_ars + blk * ctl->ars_width;
        un.n = blk < ctl->n_ars_blocks ? ctl->ars_width : 0;
      }
      ArsWarp<SgmTarget, ArsUniforms, kArsGroup> ars(tg, un);
      status = ars.sample(x0, x);
      if (ars.full && lane == 0) atomicAdd(&ctl->ars_full, 1ull);
    }
    if (status != kArsOk) {
      if (lane == 0) set_err(ctl, status == kArsNoUniforms ? kErrInjectArs : kErrArs, status);
      x = NAN;
    }
    if (lane == 0) b.sigmasbt[j] = exp(x);
  }
}

// logw | var_deltas (t prior, eta > 0): block-collective target, every warp runs the engine redundantly.
struct LogwEval {
  const double* vd;  // var_deltas[1..p]
  int p;
  double K, nu, s, eta;
  double* sm;
  __device__ void operator()(double x, double& logf, double& dlogf) const {
    const double w = exp(x);
    const double sdu = (x - s) / eta;
    const double wnu = w * nu;
    double a = 0, c = 0;
    for (int j = threadIdx.x; j < p; j += blockDim.x) {
      const double t = vd[j] + wnu;
      a += wnu / t;
      c += log(t);
    }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    dlogf = a;
    logf = c;
    logf *= (-(nu + K) / 2);
    dlogf *= (-(nu + K) / 2);
    logf += p * nu / 2 * x;
    dlogf += p * nu / 2;
    logf += -(sdu * sdu) / 2 - log(eta);
    dlogf += -sdu / eta;
  }
};

__global__ void __launch_bounds__(1024) k_logw(Geom g, Bufs b) {
  Ctl* ctl = b.ctl;
  __shared__ double sm[32];
  if (!(ctl->eta > 1e-10)) return;
  if (ctl->eta < 0.01) {
    if (threadIdx.x == 0) ctl->logw = ctl->s;
    return;
  }
  LogwEval ev{b.var_deltas + 1, ctl->p, (double)ctl->K, ctl->alpha, ctl->s, ctl->eta, sm};
  double x = 0;
  int status;
  const double x0 = ctl->logw;
  if (ctl->rng_mode == 1) {
    const long long blk = ctl->step_in_call * (long long)(ctl->p + 1) + ctl->p;
    ArsArrayUniforms un{b.inj_ars + blk * ctl->ars_width, blk < ctl->n_ars_blocks ? ctl->ars_width : 0, 0};
    ArsWarp<LogwEval, ArsArrayUniforms> ars(ev, un);
    status = ars.sample(x0, x);
    if (ars.full && threadIdx.x == 0) atomicAdd(&ctl->ars_full, 1ull);
  } else {
    ArsKeyedUniforms un{ctl->seed, ctl->step, kRngLogwArs, 0u, 0u};
    ArsWarp<LogwEval, ArsKeyedUniforms> ars(ev, un);
    status = ars.sample(x0, x);
    if (ars.full && threadIdx.x == 0) atomicAdd(&ctl->ars_full, 1ull);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (status != kArsOk) { set_err(ctl, status == kArsNoUniforms ? kErrInjectArs : kErrArs, status); x = NAN; }
    ctl->logw = x;
  }
}

// Standalone batch (test hook / htlr_cuda_ars).
__global__ void __launch_bounds__(256) k_ars_batch(int kind, int m, const double* vard, double K, double alpha,
                                                    double log_aw, const double* uniforms, int width,
                                                    unsigned long long seed, unsigned long long step, double* x_out,
                                                    int* n_unif, int* n_eval, int* status_out) {
  const int lane = threadIdx.x & (kArsGroup - 1);
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) / kArsGroup, ngrp = (gridDim.x * blockDim.x) / kArsGroup;
  for (int q = grp; q < m; q += ngrp) {
    SgmTarget tg{kind, vard[q], K, alpha, log_aw};
    const double x0 = log(tg.v / tg.K);
    double x = NAN;
    int status, nu, ne;
    if (uniforms) {
      ArsArrayUniforms un{uniforms + (long long)q * width, width, 0};
      ArsWarp<SgmTarget, ArsArrayUniforms, kArsGroup> ars(tg, un);
      status = ars.sample(x0, x);
      nu = ars.n_unif; ne = ars.n_eval;
      if (status == kArsOk && ars.full) status = kArsOkTableFull;
    } else {
      ArsKeyedUniforms un{seed, step, kRngArs, (uint32_t)(q + 1), 0u};
      ArsWarp<SgmTarget, ArsKeyedUniforms, kArsGroup> ars(tg, un);
      status = ars.sample(x0, x);
      nu = ars.n_unif; ne = ars.n_eval;
      if (status == kArsOk && ars.full) status = kArsOkTableFull;
    }
    if (lane == 0) {
      x_out[q] = (status == kArsOk || status == kArsOkTableFull) ? x : NAN;
      if (n_unif) n_unif[q] = nu;
      if (n_eval) n_eval[q] = ne;
      if (status_out) status_out[q] = status;
    }
  }
}

__global__ void k_rng_dump(unsigned long long seed, unsigned int purpose, unsigned long long step, int n_a, int n_b,
                           int kind, double shape, double* out) {
  const KeyedRng rng{seed};
  const long long tot = (long long)n_a * n_b;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const uint32_t a = (uint32_t)(e / n_b), bb = (uint32_t)(e % n_b);
    double v;
    if (kind == 0) v = rng.unif(purpose, step, a, bb);
    else if (kind == 1) v = rng.norm(purpose, step, a, bb);
    else v = rng.gamma(step, a, shape);
    out[e] = v;
  }
}

void launch_sigma(const Geom& g, const Bufs& b, cudaStream_t st) {
  // ptype is fixed for the life of the handle; the host picks the kernel
  k_sigma_t<<<g.e_grid, 256, 0, st>>>(g, b);
}
void launch_tail_t(const Geom& g, const Bufs& b, cudaStream_t st) { k_tail_t<<<g.e_grid, 256, 0, st>>>(g, b); }
void launch_sigma_ars(const Geom& g, const Bufs& b, cudaStream_t st) {
  const int p = g.nvar - 1;
  const int per_block = 256 / kArsGroup;   // features in flight per block
  int blocks = (p + per_block - 1) / per_block;
  if (blocks > g.sm_count * 8) blocks = g.sm_count * 8;
  if (blocks < 1) blocks = 1;
  k_sigma_ars<<<blocks, 256, 0, st>>>(g, b);
}
void launch_logw(const Geom& g, const Bufs& b, cudaStream_t st) { k_logw<<<1, 1024, 0, st>>>(g, b); }

void launch_ars_standalone(int kind, int m, const double* vard, double K, double alpha, double log_aw,
                           const double* uniforms, int width, unsigned long long seed, unsigned long long step,
                           double* x_out, int* n_unif, int* n_eval, int* status, cudaStream_t st) {
  const int per_block = 256 / kArsGroup;
  int blocks = (m + per_block - 1) / per_block;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_ars_batch<<<blocks, 256, 0, st>>>(kind, m, vard, K, alpha, log_aw, uniforms, width, seed, step, x_out, n_unif,
                                      n_eval, status);
}
void launch_rng_dump(unsigned long long seed, unsigned int purpose, unsigned long long step, int n_a, int n_b,
                     int kind, double shape, double* out, cudaStream_t st) {
  k_rng_dump<<<256, 256, 0, st>>>(seed, purpose, step, n_a, n_b, kind, shape, out);
}

}  // namespace htlr
