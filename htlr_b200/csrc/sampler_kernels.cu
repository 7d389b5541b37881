// HMC / restricted-Gibbs kernels of the two-sweep path (the reference's own structure: one X sweep for
// the linear predictor and one for the gradient per leapfrog step). All FP64.
//
// Reference functions replaced (paths relative to the reference root):
//   k_colsumsq     DDNloglike_ = colsum(X^2)/4                         src/gibbs.cpp:15
//   k_select       Fit::WhichUpdate                                    src/gibbs.cpp:156-170
//   k_begin        GenMomt, UpdateStepSizes, CacheOldValues(deltas),   src/gibbs.cpp:441-460, 477-481, 483-491,
//                  CompNegEnergy (old)                                 432-437
//   k_lv_sweep     triple loops of UpdatePredProb / DetachFixlv        src/gibbs.cpp:179-188, 203-227
//   k_softmax      find_normlv + exp, UpdateLogLike, residual of       src/utils.cpp:10-26, gibbs.cpp:189-190,
//                  UpdateDNlogLike (pred - ymat)                       254-261, 238
//   k_grad_sweep   triple loop of UpdateDNlogLike                      src/gibbs.cpp:239-249
//   k_post         UpdateDNlogPrior, UpdateDNlogPost, MoveMomt,        src/gibbs.cpp:267-283, 467-471, 291-297
//                  UpdateMomtAndDeltas
//   k_energy       UpdateVarDeltas, CompNegEnergy (new), IsFault,      src/gibbs.cpp:425-437, 503-516, 89-95
//                  Metropolis test
//   k_commit       keep proposal / RestoreOldValues                    src/gibbs.cpp:493-501
//   k_record       trace recording                                     src/gibbs.cpp:100-114
//
// Layout: X column-major n x nvar with even leading dimension (each feature one contiguous column);
// n-length arrays [k][i] with stride n_s; compact per-feature arrays indexed by position r in iup,
// [r*K + k].
#include <algorithm>
#include <cstdio>

#include "common.cuh"
#include "launch.h"
#include "rng.cuh"

namespace htlr {

#define DISPATCH_K(KV, ...)                          \
  switch (KV) {                                      \
    case 1: { constexpr int KT = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int KT = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int KT = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int KT = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int KT = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int KT = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int KT = 7; __VA_ARGS__; } break; \
    case 8: { constexpr int KT = 8; __VA_ARGS__; } break; \
    default: { constexpr int KT = 0; __VA_ARGS__; } break; \
  }

template <int KT>
struct KDim {
  static constexpr int kMax = KT > 0 ? KT : kMaxK;
};

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// ------------------------------------------------------------------------------------------------
__global__ void k_colsumsq(Geom g, const double* __restrict__ X, double* __restrict__ ddn) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int j = w; j < g.nvar; j += nw) {
    const double* x = X + (int64_t)j * g.ld;
    double a = 0;
    for (int i = lane; i < g.n; i += 32) a += x[i] * x[i];
    a = warp_sum(a);
    if (lane == 0) ddn[j] = a / 4;
  }
}

// ------------------------------------------------------------------------------------------------
// Ordered compaction of the restricted set. A block of 32 warps owns a contiguous segment of features, warp w the
// contiguous sub-segment [j0, j1) which it walks 32 features at a time (coalesced), so ascending order falls out of
// ballots: pass 1 counts, an exclusive scan over the 32 warp totals gives each warp its offset inside the block,
// pass 2 writes. With several blocks (one per 4096 features: a single SM reading and writing 300+ KB was 12 us
// of every iteration at p = 20000) the block totals are chained: block b publishes (epoch, count) and waits for
// the words of blocks 0..b-1 - blocks are dispatched in index order, so the ones it waits for are running or done.
constexpr int kSelMaxBlocks = 32;
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(1024) k_select(Geom g, Bufs b) {
  Ctl* ctl = b.ctl;
  timeline_mark(ctl, 0);
  __shared__ int s_w[32];
  __shared__ int s_total, s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int nvar = g.nvar, G = gridDim.x, blk = blockIdx.x;
  const double cut = ctl->init_all ? -1.0 : ctl->sgmsq_cut;
  const unsigned want = ctl->sel_epoch + 1u;                 // this launch's epoch (bumped by the last block to finish)
  const int segB = ((nvar + G - 1) / G + 1023) & ~1023;     // features per block, a multiple of 32 * 32
  const int seg = segB / 32;                                 // features per warp, a multiple of 32
  const int j0 = min(nvar, blk * segB + w * seg), j1 = min(nvar, j0 + seg);
  int cnt = 0;
#pragma unroll 16
  for (int j = j0 + lane; j - lane < j1; j += 32) {   // 16 loads in flight: two or three L2 round trips per segment
    const bool f = j < j1 && b.sigmasbt[j] > cut;
    cnt += __popc(__ballot_sync(0xffffffffu, f));
  }
  if (lane == 0) s_w[w] = cnt;
  __syncthreads();
  if (w == 0) {
    const int v = s_w[lane];
    int iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += t;
    }
    s_w[lane] = iv - v;  // exclusive warp offsets
    const int total = __shfl_sync(0xffffffffu, iv, 31);
    if (lane == 0) {
      s_total = total;
      if (G > 1) {
        const unsigned long long word = ((unsigned long long)want << 32) | (unsigned)total;
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&ctl->sel_flag[blk]), "l"(word) : "memory");
      }
    }
    // selected features in the blocks before this one
    int before = 0;
    if (lane < blk) {
      unsigned long long f;
      do { f = ld_acquire_u64(&ctl->sel_flag[lane]); } while ((unsigned)(f >> 32) != want);
      before = (int)(unsigned)f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) s_prefix = before;
  }
  __syncthreads();
  int pu = s_prefix + s_w[w];   // next position in iup for this warp
  int pf = j0 - pu;             // next position in ifix
#pragma unroll 4
  for (int j = j0 + lane; j - lane < j1; j += 32) {
    const bool in = j < j1;
    const bool f = in && b.sigmasbt[j] > cut;
    const unsigned mu = __ballot_sync(0xffffffffu, f), mf = __ballot_sync(0xffffffffu, in && !f);
    const unsigned below = (1u << lane) - 1u;
    if (f) { const int r = pu + __popc(mu & below); b.iup[r] = j; b.posof[j] = r; }
    else if (in) { b.ifix[pf + __popc(mf & below)] = j; b.posof[j] = -1; }
    pu += __popc(mu);
    pf += __popc(mf);
  }
  if (blk == G - 1 && tid == 0) {
    const int u = s_prefix + s_total;
    ctl->u = u;
    ctl->nfix = nvar - u;
    ctl->ticket[4] = 0;  // grid-barrier counter of the fused trajectory kernel
    ctl->traj_done = 0;
    if (!ctl->init_all) {
      ctl->uvar_acc += u;
      // Traject's choice of logw (gibbs.cpp:394-408); L itself is chosen by the host
      ctl->logw = (ctl->i_mc < ctl->iters_h / 2.0) ? -10.0 : ctl->s;
    }
  }
  if (G > 1 && last_block(&ctl->ticket[5], G)) {
    if (tid == 0) ctl->sel_epoch = want;   // every block has read its epoch and the words it waited for
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_begin(Geom g, Bufs b) {
  Ctl* ctl = b.ctl;
  timeline_mark(ctl, 1);
  __shared__ double sm[32];
  const int u = ctl->u, K = g.K, nvar = g.nvar;
  const KeyedRng rng{ctl->seed};
  double pa = 0, pb = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < u; r += gridDim.x * blockDim.x) {
    const int j = b.iup[r];
    const double sig = b.sigmasbt[j];
    b.sigc[r] = sig;
    b.stepc[r] = ctl->leap_step / sqrt(b.ddn[j] + (double)K / sig / (double)ctl->C);
    pa += b.var_deltas[j] / sig;
    for (int k = 0; k < K; ++k) {
      b.dc[(int64_t)r * K + k] = b.deltas[(int64_t)k * nvar + j];
      double m;
      if (b.user_momt) {
        m = b.user_momt[(int64_t)k * nvar + j];
      } else if (ctl->rng_mode == 1) {
        long long at = ctl->cur_norm + (long long)r * K + k;
        if (at >= ctl->n_norm) { set_err(ctl, kErrInjectNormals, 0); m = 0; }
        else m = b.inj_norm[at];
      } else {
        m = rng.norm(kRngMomentum, ctl->step, (uint32_t)j, (uint32_t)k);
      }
      b.momt[(int64_t)r * K + k] = m;
      pb += m * m;
    }
  }
  // CacheOldValues (gibbs.cpp:483-491) for the row-sized state: the linear predictor and the residual of the current
  // point. The fused trajectory kernels carry both from iteration to iteration (no detach sweep, no softmax
  // before the first gradient), so a rejected proposal must find them again.
  {
    const int64_t tot = (int64_t)K * g.n_s;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
      b.lv_old[e] = b.lv[e];
      b.R_old[e] = b.R[e];
    }
  }
  pa = block_sum(pa, sm);
  pb = block_sum(pb, sm);
  // The old energy (CompNegEnergy at the current point, gibbs.cpp:432-437) is only read by the Metropolis test:
  // the block partials are left for k_energy's last block to add (same order as a reduction here would use), which
  // saves this kernel a fence, a ticket and a second round of loads on every iteration.
  if (threadIdx.x == 0) {
    b.red_begin[blockIdx.x].a = pa;
    b.red_begin[blockIdx.x].b = pb;
    if (blockIdx.x == 0) { b.red_begin[0].c = (double)gridDim.x; timeline_mark_end(ctl, 11); }
  }
}

// ------------------------------------------------------------------------------------------------
// Which columns a sweep covers and how many partial-sum slots it fills. Shared by k_lv_sweep and
// k_softmax so both agree without a host round trip.
struct ColSet {
  const int* idx;
  int cnt;
  const double* cg;  // global coefficient matrix (nvar x K col-major) or null for compact dc
  int slots;
};
__device__ __forceinline__ ColSet resolve_cols(const Geom& g, const Bufs& b, int mode, const int* idx_in,
                                               const int* cnt_in, const double* coef_global) {
  const Ctl* ctl = b.ctl;
  ColSet c;
  if (mode == 0) {
    if (ctl->u <= g.nvar / 2) { c.idx = b.iup; c.cnt = ctl->u; c.cg = nullptr; }
    else { c.idx = b.ifix; c.cnt = ctl->nfix; c.cg = b.deltas; }
  } else if (mode == 1) {
    c.idx = b.iup; c.cnt = ctl->u; c.cg = nullptr;
  } else {
    c.idx = idx_in; c.cnt = *cnt_in; c.cg = coef_global;
  }
  int want = (c.cnt + 7) / 8;
  c.slots = min(g.lv_cs, max(1, want));
  return c;
}

// partial[cs][k][i] = sum over this block's share of the column set of X[i, j] * coef[j, k].
// Thread = two consecutive rows (128-bit loads, 4 KB contiguous per block per column). The block's column
// indices and coefficients are staged in shared memory in chunks (no dependent global loads in the loop, and
// the streaming X loads bypass L1 allocation so they cannot evict anything useful); kU columns per pass keep
// kU * 16 bytes per thread in flight.
constexpr int kLvChunk = 512;   // columns staged per chunk
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
template <int KT>
__global__ void __launch_bounds__(256) k_lv_sweep(Geom g, Bufs b, int mode, const int* idx_in,
                                                   const int* cnt_in, const double* coef_global) {
  constexpr int KM = KDim<KT>::kMax;
  const int K = KT > 0 ? KT : g.K;
  extern __shared__ __align__(16) unsigned char lv_smem[];
  double* s_coef = reinterpret_cast<double*>(lv_smem);                    // [kLvChunk][K]
  int64_t* s_off = reinterpret_cast<int64_t*>(s_coef + (size_t)kLvChunk * min(K, KM));  // [kLvChunk] column offsets (doubles)
  const ColSet cs = resolve_cols(g, b, mode, idx_in, cnt_in, coef_global);
  if (mode == 0) timeline_mark(b.ctl, 2);
  if (blockIdx.x == 0 && threadIdx.x == 0) b.ctl->cols_swept += (unsigned long long)cs.cnt;
  const int nvar = g.nvar;
  constexpr int kU = 8;
  // The grid is ONE full wave of resident blocks (launch_lv_sweep); a block walks the (row tile, column split) items
  // with the grid's stride, so every block gets the same number of items (the number of splits is chosen for that):
  // a second, nearly empty wave of a few blocks cannot keep HBM busy and cost 30 % at 200000 rows.
  const int items = g.lv_rt * cs.slots;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int rt = item / cs.slots, s = item - rt * cs.slots;
    const int c0 = (int)((int64_t)s * cs.cnt / cs.slots), c1 = (int)((int64_t)(s + 1) * cs.cnt / cs.slots);
    const int i = rt * 512 + 2 * threadIdx.x;
    const bool active = i < g.n;
    const bool tail = (i + 1 >= g.n);
    const double* Xi = b.X + (active ? i : 0);
    // more than kMaxK classes (generic instantiation only): kMaxK at a time, the columns are read once per class chunk
    for (int kc0 = 0; kc0 < K; kc0 += KM) {
    const int KC = min(KM, K - kc0);
    double a0[KM], a1[KM];
#pragma unroll
    for (int k = 0; k < KM; ++k) { a0[k] = 0; a1[k] = 0; }
    for (int ch0 = c0; ch0 < c1; ch0 += kLvChunk) {
      const int cn = min(kLvChunk, c1 - ch0);
      __syncthreads();   // the previous chunk is consumed
      for (int e = threadIdx.x; e < cn; e += blockDim.x) {
        const int j = __ldg(cs.idx + ch0 + e);
        s_off[e] = (int64_t)j * g.ld;
        for (int k = 0; k < KC; ++k)
          s_coef[e * KC + k] = cs.cg ? __ldg(cs.cg + (int64_t)(kc0 + k) * nvar + j)
                                     : __ldg(b.dc + (int64_t)(ch0 + e) * K + kc0 + k);
      }
      __syncthreads();
      if (!active) continue;
      int r = 0;
      for (; r + kU <= cn; r += kU) {
        double2 x[kU];
#pragma unroll
        for (int q = 0; q < kU; ++q) x[q] = ld_stream2(Xi + s_off[r + q]);
#pragma unroll
        for (int q = 0; q < kU; ++q) {
#pragma unroll
          for (int k = 0; k < KM; ++k) {
            if (k < KC) {
              const double c = s_coef[(r + q) * KC + k];
              a0[k] = fma(x[q].x, c, a0[k]);
              a1[k] = fma(x[q].y, c, a1[k]);
            }
          }
        }
      }
      for (; r < cn; ++r) {
        const double2 x = ld_stream2(Xi + s_off[r]);
#pragma unroll
        for (int k = 0; k < KM; ++k) {
          if (k < KC) {
            const double c = s_coef[r * KC + k];
            a0[k] = fma(x.x, c, a0[k]);
            a1[k] = fma(x.y, c, a1[k]);
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < KM; ++k) {
        if (k < KC) {
          double* o = b.lvpart + ((int64_t)s * K + kc0 + k) * g.n_s + i;
          if (tail) o[0] = a0[k];
          else *reinterpret_cast<double2*>(o) = make_double2(a0[k], a1[k]);
        }
      }
    }
    }   // class chunk
  }
}

// ------------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256) k_softmax(Geom g, Bufs b, int what, const int* cnt_in, double* lv_out,
                                                  double* R_out, double* pred_out, double* loglike_out) {
  constexpr int KM = KDim<KT>::kMax;
  const int K = KT > 0 ? KT : g.K;
  Ctl* ctl = b.ctl;
  __shared__ double sm[32];
  if (what == 0) timeline_mark(ctl, 3);
  const ColSet cs = resolve_cols(g, b, what == 0 ? 0 : (what == 1 ? 1 : 2), nullptr, cnt_in, nullptr);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ll = 0;
  if (i < g.n) {
    double l[KM];
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      if (k < K) {
        // slots summed in order, 16 loads in flight at a time (one L2 round trip per 16 slots, not per slot)
        double sacc = 0;
        for (int s0 = 0; s0 < cs.slots; s0 += 16) {
          double pv[16];
#pragma unroll
          for (int q = 0; q < 16; ++q)
            pv[q] = (s0 + q < cs.slots) ? __ldcg(&b.lvpart[((int64_t)(s0 + q) * K + k) * g.n_s + i]) : 0.0;
#pragma unroll
          for (int q = 0; q < 16; ++q) sacc += pv[q];
        }
        const int64_t e = (int64_t)k * g.n_s + i;
        if (what == 0) {
          const double cur = b.lv[e];
          b.lv_old[e] = cur;
          b.lv_fix[e] = (ctl->u <= g.nvar / 2) ? cur - sacc : sacc;
          l[k] = cur;
        } else if (what == 1) {
          l[k] = b.lv_fix[e] + sacc;
          b.lv[e] = l[k];
        } else {
          l[k] = sacc;
          lv_out[e] = sacc;
        }
      }
    }
    // find_normlv (utils.cpp:10-26): max-shifted log-sum-exp over the C classes, class 0 has lv = 0
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < KM; ++k) if (k < K) m = fmax(m, l[k]);
    double se = exp(0.0 - m);
#pragma unroll
    for (int k = 0; k < KM; ++k) if (k < K) se += exp(l[k] - m);
    const double lse = log(se) + m;
    const int yb = b.ybase[i];
    if (yb == 0) ll = 0.0 - lse;
    if (pred_out) pred_out[i] = exp(0.0 - lse);
    double* R = what == 2 ? R_out : b.R;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      if (k < K) {
        const double nl = l[k] - lse;
        const double pr = exp(nl);
        if (yb == k + 1) ll = nl;
        if (pred_out) pred_out[(int64_t)(k + 1) * g.n + i] = pr;
        R[(int64_t)k * g.n_s + i] = pr - b.ymat[(int64_t)k * g.n_s + i];
      }
    }
  }
  if (what == 0) return;
  ll = block_sum(ll, sm);
  if (threadIdx.x == 0) b.red[blockIdx.x].c = ll;
  if (last_block(&ctl->ticket[1], gridDim.x)) {
    double A = 0;
    for (unsigned q = threadIdx.x; q < gridDim.x; q += blockDim.x) A += b.red[q].c;
    A = block_sum(A, sm);
    if (threadIdx.x == 0) {
      if (what == 1) ctl->loglike_new = A; else *loglike_out = A;
    }
  }
}

// More than kMaxK classes: the same kernel with the row's linear predictors in their global array instead of
// registers (three passes over the classes: sum the partials and find the maximum, sum the exponentials, residuals).
__global__ void __launch_bounds__(256) k_softmax_wide(Geom g, Bufs b, int what, const int* cnt_in, double* lv_out,
                                                       double* R_out, double* pred_out, double* loglike_out) {
  const int K = g.K;
  Ctl* ctl = b.ctl;
  __shared__ double sm[32];
  const ColSet cs = resolve_cols(g, b, what == 0 ? 0 : (what == 1 ? 1 : 2), nullptr, cnt_in, nullptr);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ll = 0;
  if (i < g.n) {
    double* lvd = what == 2 ? lv_out : b.lv;   // where the row's predictors live
    double m = 0.0;
    for (int k = 0; k < K; ++k) {
      double sacc = 0;
      for (int s0 = 0; s0 < cs.slots; ++s0) sacc += __ldcg(&b.lvpart[((int64_t)s0 * K + k) * g.n_s + i]);
      const int64_t e = (int64_t)k * g.n_s + i;
      double l;
      if (what == 0) {
        l = b.lv[e];
        b.lv_old[e] = l;
        b.lv_fix[e] = (ctl->u <= g.nvar / 2) ? l - sacc : sacc;
      } else if (what == 1) {
        l = b.lv_fix[e] + sacc;
        lvd[e] = l;
      } else {
        l = sacc;
        lvd[e] = l;
      }
      m = fmax(m, l);
    }
    double se = exp(0.0 - m);
    for (int k = 0; k < K; ++k) se += exp(lvd[(int64_t)k * g.n_s + i] - m);
    const double lsl = log(se) + m;
    const int yb = b.ybase[i];
    if (yb == 0) ll = 0.0 - lsl;
    if (pred_out) pred_out[i] = exp(0.0 - lsl);
    double* R = what == 2 ? R_out : b.R;
    for (int k = 0; k < K; ++k) {
      const double nl = lvd[(int64_t)k * g.n_s + i] - lsl;
      const double pr = exp(nl);
      if (yb == k + 1) ll = nl;
      if (pred_out) pred_out[(int64_t)(k + 1) * g.n + i] = pr;
      R[(int64_t)k * g.n_s + i] = pr - b.ymat[(int64_t)k * g.n_s + i];
    }
  }
  if (what == 0) return;
  ll = block_sum(ll, sm);
  if (threadIdx.x == 0) b.red[blockIdx.x].c = ll;
  if (last_block(&ctl->ticket[1], gridDim.x)) {
    double A = 0;
    for (unsigned q = threadIdx.x; q < gridDim.x; q += blockDim.x) A += b.red[q].c;
    A = block_sum(A, sm);
    if (threadIdx.x == 0) {
      if (what == 1) ctl->loglike_new = A; else *loglike_out = A;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// g[r, k] = sum_i X[i, idx[r]] * R[i, k].
// A block = 8 warps = 8 consecutive row sub-tiles of PP*64 rows; a lane keeps the residual of ITS rows (PP row pairs
// x K classes) in registers for all the columns it sees - the first version staged the residual tile in shared
// memory and read K * 16 bytes of it per 16 bytes of X, which at K = 4 is 70 % of an SM's shared-memory bandwidth at
// HBM speed and held the sweep at 4.4 TB/s (ncu: profiles/r02_ncu_tall_sweeps.txt). All warps of the block take the
// same CU columns at a time. The columns come through a THREAD-PRIVATE ring in shared memory filled with cp.async
// (16 bytes per copy, kGradDepth passes of CU * PP copies per lane in flight: 192 KB per SM): with the loads held in
// registers, two passes (128 KB per SM at 254 registers) were the most that fit and the sweep stayed at 4.9 TB/s,
// 43 % of the stall samples waiting for the next pass's loads; a ring slot belongs to one thread, so the ring itself
// needs no barrier. The warps reduce their CU*K dot products
// over the lanes with a transposing butterfly (V-1 + 5-log2(V) shuffles instead of 5V), then over the 8 warps
// through shared memory (double-buffered: one block barrier per CU columns). Fixed order everywhere: deterministic.
// The grid is one full wave; a block walks (row tile, column group) items with the grid's stride (see k_lv_sweep).
constexpr int kGradDepth = 3;   // passes in flight per thread
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
template <int KT>
struct GradShape {   // PP row pairs per lane, CU columns per pass: PP*KT <= 16 (residual: 64 registers), PP*CU <= 16
  static constexpr int PP = KT == 0 ? 1 : (KT <= 2 ? 8 : (KT <= 4 ? 4 : 2));
  static constexpr int CU = KT == 0 ? 2 : (KT > 4 ? 4 : 16 / PP);   // CU * K <= 32 dot products per pass
};
template <int KT>
constexpr size_t grad_smem_bytes() {
  return (size_t)kGradDepth * GradShape<KT>::CU * GradShape<KT>::PP * 256 * sizeof(double2);
}
constexpr int grad_rows_per_block(int K) { return 8 * 64 * (K > 8 ? 1 : (K <= 2 ? 8 : (K <= 4 ? 4 : 2))); }

template <int KT>
__global__ void __launch_bounds__(256) k_grad_sweep(Geom g, Bufs b, const int* idx, const int* cnt_in,
                                                     const double* __restrict__ R, double* __restrict__ gout) {
  constexpr int KM = KDim<KT>::kMax;
  constexpr int PP = GradShape<KT>::PP, CU = GradShape<KT>::CU;
  constexpr int V = CU * KM;                                         // dot products per pass
  constexpr int VP = V <= 2 ? 2 : (V <= 4 ? 4 : (V <= 8 ? 8 : (V <= 16 ? 16 : 32)));
  constexpr int LOG = VP == 2 ? 1 : (VP == 4 ? 2 : (VP == 8 ? 3 : (VP == 16 ? 4 : 5)));
  static_assert(V <= 32, "at most 32 dot products per pass");
  const int K = KT > 0 ? KT : g.K;
  __shared__ double s_part[2][8][VP];
  extern __shared__ __align__(16) double2 ring[];   // [kGradDepth][CU][PP][256 threads]
  const int cnt = *cnt_in;
  if (blockIdx.x == 0 && threadIdx.x == 0) b.ctl->cols_swept += (unsigned long long)cnt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int vmine = lane >> (5 - LOG);                     // the value this lane holds after the transposing reduction
  const bool vwriter = (lane & ((32 >> LOG) - 1)) == 0;
  const int items = g.g_rt * g.g_gx;
  int buf = 0;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int rt = item / g.g_gx, cg = item - rt * g.g_gx;          // row tile (g_tr rows), column group
    const int i0 = rt * g.g_tr + warp * (PP * 64);                   // this warp's sub-tile
    // more than kMaxK classes (generic instantiation only): kMaxK at a time, the columns are read once per class chunk
    for (int kc0 = 0; kc0 < K; kc0 += KM) {
    const int KC = min(KM, K - kc0);
    // residual of this lane's rows (rows >= n hold 0 in R up to the stride n_s; beyond it: nothing to load)
    double rr[PP][2][KM];
    bool have[PP], odd_last[PP];
#pragma unroll
    for (int pp = 0; pp < PP; ++pp) {
      const int row = i0 + 2 * (lane + 32 * pp);
      have[pp] = row < g.n;
      odd_last[pp] = row + 1 >= g.n;
#pragma unroll
      for (int k = 0; k < KM; ++k) {
        double2 t = make_double2(0.0, 0.0);
        if (k < KC && have[pp]) t = *reinterpret_cast<const double2*>(R + (int64_t)(kc0 + k) * g.n_s + row);
        rr[pp][0][k] = t.x;
        rr[pp][1][k] = odd_last[pp] ? 0.0 : t.y;
      }
    }
    const double* xt = b.X + i0 + 2 * lane;
    double* out = gout + (int64_t)rt * g.nvar * K;
    // pass i+D is requested as soon as pass i has been consumed; one commit group per pass, empty ones included, so
    // that "all but the newest D-1 groups are complete" always means "the pass to consume has landed"
    auto fetch = [&](int slot, int r0) {
      if (r0 < cnt) {
#pragma unroll
        for (int c = 0; c < CU; ++c) {
          const int64_t off = (int64_t)__ldg(idx + min(r0 + c, cnt - 1)) * g.ld;
#pragma unroll
          for (int pp = 0; pp < PP; ++pp)
            if (have[pp]) cp_async16(&ring[((slot * CU + c) * PP + pp) * 256 + threadIdx.x], xt + off + 64 * pp);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int rstep = g.g_gx * CU;
#pragma unroll
    for (int d = 0; d < kGradDepth; ++d) fetch(d, cg * CU + d * rstep);
    int slot = 0;
    for (int r0 = cg * CU; r0 < cnt; r0 += rstep) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kGradDepth - 1) : "memory");
      double2 x[CU][PP];
#pragma unroll
      for (int c = 0; c < CU; ++c)
#pragma unroll
        for (int pp = 0; pp < PP; ++pp)
          x[c][pp] = have[pp] ? ring[((slot * CU + c) * PP + pp) * 256 + threadIdx.x] : make_double2(0.0, 0.0);
      fetch(slot, r0 + kGradDepth * rstep);   // the slot is free again (its values sit in registers)
      if (++slot == kGradDepth) slot = 0;
      double v[VP];
#pragma unroll
      for (int c = 0; c < CU; ++c) {
#pragma unroll
        for (int k = 0; k < KM; ++k) {
          double a0 = 0, a1 = 0;
          if (k < KC) {
#pragma unroll
            for (int pp = 0; pp < PP; ++pp) {
              a0 = fma(x[c][pp].x, rr[pp][0][k], a0);
              a1 = fma(odd_last[pp] ? 0.0 : x[c][pp].y, rr[pp][1][k], a1);   // a pad row may hold anything
            }
          }
          v[c * KM + k] = a0 + a1;
        }
      }
#pragma unroll
      for (int q = V; q < VP; ++q) v[q] = 0;
      // transposing butterfly: the value set is halved at each of the first LOG levels, plain xor levels after
      {
        int half = VP / 2;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          if (half >= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int q = 0; q < VP / 2; ++q) {
              if (q < half) {
                const double send = upper ? v[q] : v[q + half];
                const double keep = upper ? v[q + half] : v[q];
                v[q] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            half >>= 1;
          } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
          }
        }
      }
      if (vwriter) s_part[buf][warp][vmine] = v[0];
      __syncthreads();
      if (warp == 0 && lane < V) {
        double a = s_part[buf][0][lane];
#pragma unroll
        for (int ww = 1; ww < 8; ++ww) a += s_part[buf][ww][lane];
        const int c = lane / KM, k = lane - c * KM;
        if (k < KC && r0 + c < cnt) out[(int64_t)(r0 + c) * K + kc0 + k] = a;
      }
      buf ^= 1;   // the next pass writes the other buffer: one barrier per pass is enough
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // (only empty groups are left)
    __syncthreads();
    }   // class chunk
  }
}

// Sum the row-tile partials (needed on its own only before an all-reduce).
__global__ void k_gsum(Geom g, Bufs b, const int* cnt_in, double* gout) {
  const int cnt = *cnt_in, K = g.K;
  const int64_t total = (int64_t)cnt * K, all = (int64_t)g.nvar * K;
  // the whole nvar*K buffer goes through the all-reduce (its length must be known when the graph is captured, the
  // restricted-set size is not): entries past the restricted set are zeroed, so that nothing stale is ever re-summed
  // in place (they used to grow by a factor of `world` per call)
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < all; e += (int64_t)gridDim.x * blockDim.x) {
    double a = 0;
    if (e < total)
      for (int rt = 0; rt < g.g_rt; ++rt) a += b.gpart[(int64_t)rt * g.nvar * K + e];
    gout[e] = a;
  }
  // the local log-likelihood rides along in the trailing slot (row-sharded all-reduce)
  if (blockIdx.x == 0 && threadIdx.x == 0) gout[(int64_t)g.nvar * K] = b.ctl->loglike_new;
}

// ------------------------------------------------------------------------------------------------
// After gradient sweep s (0..L): DNlogprior, DNlogpost, then MoveMomt of step s (s >= 1) and
// UpdateMomtAndDeltas of step s+1 (s < L) — both use the same DNlogpost, two separate half kicks as in
// the reference. Products are kept un-fused so every element matches the CPU arithmetic.
__global__ void __launch_bounds__(256) k_post(Geom g, Bufs b, int s, int L, int g_is_summed) {
  const Ctl* ctl = b.ctl;
  const int u = ctl->u, K = g.K;
  const double Cd = (double)ctl->C;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < u; r += gridDim.x * blockDim.x) {
    double* d = b.dc + (int64_t)r * K;
    double* mo = b.momt + (int64_t)r * K;
    const double sig = b.sigc[r], st = b.stepc[r], hs = st / 2;
    double sd = d[0];
    for (int k = 1; k < K; ++k) sd += d[k];
    const double sdc = sd / Cd;
    for (int k = 0; k < K; ++k) {
      double gl;
      if (g_is_summed || g.g_rt == 1) {
        gl = b.g[(int64_t)r * K + k];
      } else {
        gl = 0;
        for (int rt = 0; rt < g.g_rt; ++rt) gl += b.gpart[(int64_t)rt * g.nvar * K + (int64_t)r * K + k];
      }
      const double post = gl + (d[k] - sdc) / sig;
      double m = mo[k];
      const double kick = __dmul_rn(hs, post);
      if (s >= 1) m = __dsub_rn(m, kick);
      if (s < L) {
        m = __dsub_rn(m, kick);
        d[k] = __dadd_rn(d[k], __dmul_rn(st, m));
      }
      mo[k] = m;
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_energy(Geom g, Bufs b, int g_has_loglike) {
  Ctl* ctl = b.ctl;
  timeline_mark(ctl, 7);
  __shared__ double sm[32];
  const int u = ctl->u, K = g.K;
  const double Cd = (double)ctl->C;
  double pa = 0, pb = 0, fault = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < u; r += gridDim.x * blockDim.x) {
    const double* d = b.dc + (int64_t)r * K;
    const double* mo = b.momt + (int64_t)r * K;
    double sd = d[0], ss = __dmul_rn(d[0], d[0]);
    if (fabs(d[0]) > 20) fault = 1;
    pb += mo[0] * mo[0];
    for (int k = 1; k < K; ++k) {
      sd += d[k];
      ss = __dadd_rn(ss, __dmul_rn(d[k], d[k]));
      if (fabs(d[k]) > 20) fault = 1;
      pb += mo[k] * mo[k];
    }
    const double vd = ss - __dmul_rn(sd, sd) / Cd;
    b.vdc[r] = vd;
    pa += vd / b.sigc[r];
  }
  pa = block_sum(pa, sm);
  pb = block_sum(pb, sm);
  fault = block_sum(fault, sm);
  if (threadIdx.x == 0) { b.red[blockIdx.x].a = pa; b.red[blockIdx.x].b = pb; b.red[blockIdx.x].d = fault; }
  if (last_block(&ctl->ticket[2], gridDim.x)) {
    double A = 0, B = 0, F = 0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) { A += b.red[i].a; B += b.red[i].b; F += b.red[i].d; }
    A = block_sum(A, sm);
    B = block_sum(B, sm);
    F = block_sum(F, sm);
    // k_begin's partials: the energy at the starting point
    double A0 = 0, B0 = 0;
    const unsigned nb0 = (unsigned)b.red_begin[0].c;
    for (unsigned i = threadIdx.x; i < nb0; i += blockDim.x) { A0 += b.red_begin[i].a; B0 += b.red_begin[i].b; }
    A0 = block_sum(A0, sm);
    B0 = block_sum(B0, sm);
    if (threadIdx.x == 0) {
      ctl->nenergy_old = ctl->loglike - A0 / 2 - B0 / 2;
      ctl->loglike_old = ctl->loglike;
      if (g_has_loglike) ctl->loglike_new = b.g[(int64_t)g.nvar * K];
      const double ne = ctl->loglike_new - A / 2 - B / 2;
      ctl->nenergy_new = ne;
      double uu;
      if (ctl->rng_mode == 1) {
        if (ctl->cur_unif >= ctl->n_unif) { set_err(ctl, kErrInjectUniforms, 0); uu = 0.5; }
        else uu = b.inj_unif[ctl->cur_unif];
      } else {
        const KeyedRng rng{ctl->seed};
        uu = rng.unif(kRngAccept, ctl->step, 0, 0);
      }
      // gibbs.cpp:90 — NaN energies fall through to "accept", exactly as in the reference
      const bool reject = (log(uu) > (ne - ctl->nenergy_old)) || (F > 0);
      ctl->accept = reject ? 0 : 1;
      if (reject) ctl->rej_acc += 1;
      timeline_mark_end(ctl, 10);
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_commit(Geom g, Bufs b) {
  Ctl* ctl = b.ctl;
  const int u = ctl->u, K = g.K, nvar = g.nvar;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (ctl->accept) {
    for (int r = tid; r < u; r += nt) {
      const int j = b.iup[r];
      for (int k = 0; k < K; ++k) b.deltas[(int64_t)k * nvar + j] = b.dc[(int64_t)r * K + k];
      b.var_deltas[j] = b.vdc[r];
    }
    if (tid == 0) ctl->loglike = ctl->loglike_new;
  } else {
    const int64_t tot = (int64_t)K * g.n_s;
    for (int64_t e = tid; e < tot; e += nt) { b.lv[e] = b.lv_old[e]; b.R[e] = b.R_old[e]; }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_record(Geom g, Bufs b) {
  Ctl* ctl = b.ctl;
  const int nvar = g.nvar, K = g.K;
  const bool last_thin = (ctl->i_thin == ctl->thin - 1);
  const int i_rmc = ctl->keep_warmup_hist ? (ctl->i_mc + 1) : (ctl->i_mc - ctl->iters_h + 1);
  if (last_thin && i_rmc > 0) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    double* md = b.mcdeltas + (int64_t)i_rmc * nvar * K;
    for (int64_t e = tid; e < (int64_t)nvar * K; e += nt) md[e] = b.deltas[e];
    double* ms = b.mcsigmasbt + (int64_t)i_rmc * nvar;
    double* mv = b.mcvardeltas + (int64_t)i_rmc * nvar;
    for (int j = tid; j < nvar; j += nt) { ms[j] = b.sigmasbt[j]; mv[j] = b.var_deltas[j]; }
  }
  if (last_block(&ctl->ticket[3], gridDim.x)) {
    if (threadIdx.x == 0) {
      // advance inject cursors (reference consumption: u*K normals, 1 uniform, p gammas per thin-step)
      ctl->cur_norm += (long long)ctl->u * K;
      ctl->cur_unif += 1;
      if (ctl->ptype == 0) ctl->cur_gam += ctl->p;
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
    }
  }
}

// UpdateDNlogPrior + UpdateVarDeltas over every feature (Initialize, gibbs.cpp:527-528)
__global__ void k_vardeltas_all(Geom g, Bufs b) {
  const int nvar = g.nvar, K = g.K;
  const double Cd = (double)(K + 1);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nvar; j += gridDim.x * blockDim.x) {
    double sd = b.deltas[j], ss = __dmul_rn(sd, sd);
    for (int k = 1; k < K; ++k) {
      const double d = b.deltas[(int64_t)k * nvar + j];
      sd += d;
      ss = __dadd_rn(ss, __dmul_rn(d, d));
    }
    b.var_deltas[j] = ss - __dmul_rn(sd, sd) / Cd;
  }
}

// Proposal in full (nvar-sized) layout for the trajectory test hook.
__global__ void k_scatter_proposal(Geom g, Bufs b, double* deltas_out, double* momt_out, double* vd_out) {
  const Ctl* ctl = b.ctl;
  const int u = ctl->u, K = g.K, nvar = g.nvar;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < u; r += gridDim.x * blockDim.x) {
    const int j = b.iup[r];
    for (int k = 0; k < K; ++k) {
      deltas_out[(int64_t)k * nvar + j] = b.dc[(int64_t)r * K + k];
      momt_out[(int64_t)k * nvar + j] = b.momt[(int64_t)r * K + k];
    }
    vd_out[j] = b.vdc[r];
  }
}

// ------------------------------------------------------------------------------------------------
// Posterior summaries straight from the device-resident trace (R/mccoef.r:119-156: summary(method = mean) over
// the used slices, htlr_sdb = sqrt(comp_vardeltas(mean deltas)/C) without the intercept, nzero_idx =
// which(sdb > cut * max(sdb))). Slices are summed in ascending order, one thread per coefficient.
__global__ void k_trace_mean(Geom g, Bufs b, int first, int last, int thin, double* __restrict__ mean_out) {
  const int64_t tot = (int64_t)g.nvar * g.K;
  int cnt = 0;
  for (int sidx = first; sidx <= last; sidx += thin) ++cnt;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
    double a = 0;
    for (int sidx = first; sidx <= last; sidx += thin) a += b.mcdeltas[(int64_t)sidx * tot + e];
    mean_out[e] = a / (double)cnt;
  }
}
// sdb[j-1] for features j = 1..p and per-block maxima (red[].a)
__global__ void __launch_bounds__(256) k_trace_sdb(Geom g, Bufs b, const double* __restrict__ mean, double* __restrict__ sdb) {
  __shared__ double sm[32];
  const int nvar = g.nvar, K = g.K;
  const double Cd = (double)(K + 1);
  double mx = 0;
  for (int j = 1 + blockIdx.x * blockDim.x + threadIdx.x; j < nvar; j += gridDim.x * blockDim.x) {
    double sd = mean[j], ss = __dmul_rn(sd, sd);
    for (int k = 1; k < K; ++k) {
      const double d = mean[(int64_t)k * nvar + j];
      sd += d;
      ss = __dadd_rn(ss, __dmul_rn(d, d));
    }
    const double v = sqrt((ss - __dmul_rn(sd, sd) / Cd) / Cd);   // comp_vardeltas (utils.cpp:51-56), comp_sdb (R/util.r:91-105)
    sdb[j - 1] = v;
    mx = fmax(mx, v);
  }
  // block maximum (values are >= 0 or NaN; fmax ignores NaN like R's max(..., na.rm) would not - documented)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) sm[w] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, sm[i]);
    b.red[blockIdx.x].a = t;
  }
}
__global__ void k_trace_nzero(Geom g, Bufs b, const double* __restrict__ sdb, double cut, int nblocks, int* __restrict__ flags) {
  double mx = 0;
  for (int i = 0; i < nblocks; ++i) mx = fmax(mx, b.red[i].a);
  const double thr = cut * mx;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < g.nvar - 1; j += gridDim.x * blockDim.x) flags[j] = sdb[j] > thr;
}

// ================================================================================================
void launch_trace_summary(const Geom& g, const Bufs& b, int first, int last, int thin, double cut, double* mean_out,
                          double* sdb_out, int* flags_out, cudaStream_t st) {
  const int64_t tot = (int64_t)g.nvar * g.K;
  const int gm = (int)std::min<int64_t>((tot + 255) / 256, 8 * (int64_t)g.sm_count);
  k_trace_mean<<<gm, 256, 0, st>>>(g, b, first, last, thin, mean_out);
  k_trace_sdb<<<g.e_grid, 256, 0, st>>>(g, b, mean_out, sdb_out);
  k_trace_nzero<<<g.e_grid, 256, 0, st>>>(g, b, sdb_out, cut, g.e_grid, flags_out);
}
void launch_colsumsq(const Geom& g, const double* X, double* ddn, cudaStream_t st) {
  k_colsumsq<<<g.sm_count * 4, 256, 0, st>>>(g, X, ddn);
}
void launch_select(const Geom& g, const Bufs& b, cudaStream_t st) {
  // measured (tools/timeline.py): p = 20000, 5 blocks 8.2 us against 13 us with one; p = 5000, 2 blocks 6.9 us against
  // 5.2 us with one (the chained hand-over costs ~2 us) - so one block up to 8192 features
  const int blocks = g.nvar <= 8192 ? 1 : std::min(kSelMaxBlocks, (g.nvar + 4095) / 4096);
  k_select<<<blocks, 1024, 0, st>>>(g, b);
}
void launch_begin(const Geom& g, const Bufs& b, cudaStream_t st) { k_begin<<<g.e_grid, 256, 0, st>>>(g, b); }

static size_t lv_smem_bytes(int K) {   // <= 72 KB (kMaxK classes at a time)
  return (size_t)kLvChunk * (std::min(K, kMaxK) * sizeof(double) + sizeof(int64_t));
}

void launch_lv_sweep(const Geom& g, const Bufs& b, int mode, const int* idx, const int* cnt,
                     const double* coef_global, cudaStream_t st) {
  const size_t smem = lv_smem_bytes(g.K);
  DISPATCH_K(g.K, {
    static PerDeviceOnce configured;
    if (configured.needed()) {
      cudaFuncSetAttribute(k_lv_sweep<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
      configured.mark();
    }
    k_lv_sweep<KT><<<g.lv_grid, 256, smem, st>>>(g, b, mode, idx, cnt, coef_global);
  });
}

// Resident blocks per SM of the two sweep kernels (the grids are sized to exactly one wave of them).
int lv_sweep_blocks_per_sm(int K) {
  int nb = 0;
  DISPATCH_K(K, {
    cudaFuncSetAttribute(k_lv_sweep<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_lv_sweep<KT>, 256, lv_smem_bytes(K));
  });
  return nb > 0 ? nb : 1;
}
int grad_sweep_blocks_per_sm(int K) {
  int nb = 0;
  DISPATCH_K(K, {
    cudaFuncSetAttribute(k_grad_sweep<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grad_smem_bytes<KT>());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_grad_sweep<KT>, 256, grad_smem_bytes<KT>());
  });
  return nb > 0 ? nb : 1;
}
int grad_sweep_tile_rows(int K) { return grad_rows_per_block(K); }
void launch_softmax(const Geom& g, const Bufs& b, int what, const int* cnt, double* lv_out, double* R_out,
                    double* pred_out, double* loglike_out, cudaStream_t st) {
  const int grid = (g.n + 255) / 256;
  if (g.K > kMaxK) {
    k_softmax_wide<<<grid, 256, 0, st>>>(g, b, what, cnt, lv_out, R_out, pred_out, loglike_out);
    return;
  }
  DISPATCH_K(g.K, k_softmax<KT><<<grid, 256, 0, st>>>(g, b, what, cnt, lv_out, R_out, pred_out, loglike_out));
}
void launch_grad_sweep(const Geom& g, const Bufs& b, const int* idx, const int* cnt, const double* R,
                       double* gout, cudaStream_t st) {
  DISPATCH_K(g.K, {
    static PerDeviceOnce configured;
    if (configured.needed()) {
      cudaFuncSetAttribute(k_grad_sweep<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grad_smem_bytes<KT>());
      configured.mark();
    }
    k_grad_sweep<KT><<<g.g_grid, 256, grad_smem_bytes<KT>(), st>>>(g, b, idx, cnt, R, gout);
  });
}
void launch_gsum(const Geom& g, const Bufs& b, const int* cnt, double* gout, cudaStream_t st) {
  k_gsum<<<g.e_grid, 256, 0, st>>>(g, b, cnt, gout);
}
void launch_post(const Geom& g, const Bufs& b, int s, int L, int g_is_summed, cudaStream_t st) {
  k_post<<<g.e_grid, 256, 0, st>>>(g, b, s, L, g_is_summed);
}
void launch_energy(const Geom& g, const Bufs& b, int g_has_loglike, cudaStream_t st) {
  k_energy<<<g.e_grid, 256, 0, st>>>(g, b, g_has_loglike);
}
void launch_commit(const Geom& g, const Bufs& b, cudaStream_t st) { k_commit<<<g.e_grid, 256, 0, st>>>(g, b); }
void launch_record(const Geom& g, const Bufs& b, cudaStream_t st) { k_record<<<g.e_grid, 256, 0, st>>>(g, b); }
void launch_vardeltas_all(const Geom& g, const Bufs& b, cudaStream_t st) {
  k_vardeltas_all<<<g.e_grid, 256, 0, st>>>(g, b);
}
void launch_scatter_proposal(const Geom& g, const Bufs& b, double* deltas_out, double* momt_out, double* vd_out,
                             cudaStream_t st) {
  k_scatter_proposal<<<g.e_grid, 256, 0, st>>>(g, b, deltas_out, momt_out, vd_out);
}

}  // namespace htlr
