// Fused leapfrog trajectory: ONE persistent cooperative kernel runs all L+1 gradient evaluations of an
// HMC proposal, reading each restricted-set column of X once per leapfrog step (the reference reads it
// twice: UpdatePredProb src/gibbs.cpp:179-188 and UpdateDNlogLike src/gibbs.cpp:239-249).
//
// Why one sweep suffices: MoveMomt of step t (gibbs.cpp:467-471) and UpdateMomtAndDeltas of step t+1
// (gibbs.cpp:291-297) use the same DNlogpost, and both the gradient of column j and column j's
// contribution to the next linear predictor need only column j plus the n x K residual. So while a
// column tile sits in shared memory the block computes g_j = X_j' r, applies the two half kicks as two
// separate subtractions (bit-compatible with the reference), drifts delta_j, and accumulates X_j delta_j
// into its private lv partial. Algorithmic X traffic per iteration: 8 n (L+1) u bytes (+ detach).
//
// Structure (grid = one 512-thread block per SM, cooperative launch):
//   * each block owns a contiguous range of restricted-set columns, split in tiles of T columns;
//   * thread 0 streams the tiles into a shared-memory ring with cp.async.bulk (1-D TMA copies, one per
//     column, completion on an mbarrier); the stream runs ahead across leapfrog steps, and if the
//     block's whole range fits in the ring it is loaded once and stays resident for the trajectory;
//   * gradient phase: lane <-> (column of the tile, row subgroup), 128-bit conflict-free transposed
//     reads (column stride chosen so 8 consecutive lanes hit 8 distinct 16-byte bank groups), the
//     residual rows a thread needs are held in registers for the whole sweep; no shuffles across rows
//     except log2(32/T) steps, then a fixed-order sum over the 16 warps;
//   * post phase: T threads apply prior gradient, kicks and drift (un-fused products);
//   * lv phase: thread <-> two consecutive rows, loop over the tile's columns;
//   * after the sweep the per-block lv partials are summed in block order (deterministic), rows are
//     soft-maxed and the new residual published — two grid-wide syncs per leapfrog step.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>

#include "common.cuh"
#include "launch.h"

namespace cg = cooperative_groups;

namespace htlr {

namespace {

constexpr int kFW = 16;          // warps per block
constexpr int kFT = kFW * 32;    // threads per block
constexpr int kNlpMax = 2;       // row pairs per thread in the lv phase (n <= 2048)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

template <int KT>
struct NrpMax {
  static constexpr int v = KT <= 2 ? 8 : 4;
};

}  // namespace

template <int KT, int TT>
__global__ void __launch_bounds__(kFT, 1) k_traj_fused(Geom g, Bufs b, FusedGeom fg, int L) {
  constexpr int RS = 32 / TT;            // row subgroups per warp in the gradient phase
  constexpr int NRPM = NrpMax<KT>::v;    // row pairs a thread covers in the gradient phase (max)
  cg::grid_group grid = cg::this_grid();
  Ctl* ctl = b.ctl;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = fg.S;
  double* ring = reinterpret_cast<double*>(smem_raw);                      // [NS][TT][S]
  double* gw = ring + (size_t)fg.NS * TT * S;                               // [kFW][TT*KT]
  double* dts = gw + kFW * TT * KT;                                         // [TT*KT]
  double* redsm = dts + TT * KT;                                            // [32]
  double* dl = redsm + 32;                                                  // [rpc*KT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dl + fg.rpc * KT);           // [NS]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int n = g.n, n_s = g.n_s, npairs = (n + 1) >> 1;
  const int u = ctl->u;
  const double Cd = (double)ctl->C;

  // ---- column ownership: whole tiles, contiguous, only as many blocks as there are tiles
  const int tiles_total = (u + TT - 1) / TT;
  const int gact = min(G, tiles_total);
  int t0 = 0, t1 = 0;
  if (cta < gact) {
    t0 = (int)((int64_t)cta * tiles_total / gact);
    t1 = (int)((int64_t)(cta + 1) * tiles_total / gact);
  }
  const int ntile = t1 - t0;
  const int c0 = t0 * TT, c1 = min(u, t1 * TT);
  const bool resident = ntile <= fg.NS;
  const int64_t total_loads = resident ? ntile : (int64_t)ntile * (L + 1);
  const uint32_t col_bytes = (uint32_t)(((n + 1) >> 1) << 4);  // n rounded up to even, in bytes

  if (tid == 0) {
    for (int sidx = 0; sidx < fg.NS; ++sidx) mbar_init(&bars[sidx], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t q) {  // thread 0 only
    const int ti = (int)(q % ntile), stage = (int)(q % fg.NS);
    const int r0 = c0 + ti * TT;
    const int ncol = min(TT, c1 - r0);
    mbar_expect_tx(&bars[stage], col_bytes * (uint32_t)ncol);
    double* dst = ring + (size_t)stage * TT * S;
    for (int jj = 0; jj < ncol; ++jj) {
      const int j = b.iup[r0 + jj];
      bulk_g2s(dst + (size_t)jj * S, b.X + (int64_t)j * g.ld, col_bytes, &bars[stage]);
    }
  };
  if (tid == 0) {
    const int64_t pre = total_loads < fg.NS ? total_loads : fg.NS;
    for (int64_t q = 0; q < pre; ++q) issue(q);
    if (cta == 0) ctl->cols_swept += (unsigned long long)u * (unsigned long long)(L + 1);
  }

  // gradient-phase coordinates
  const int gj = lane % TT, grs = lane / TT;
  double Rg[NRPM][2][KT];
  int64_t q = 0;  // running tile counter of this block (ring position)

  for (int s = 0; s <= L; ++s) {
    // ---- residual rows of this thread into registers (published by the previous softmax phase)
#pragma unroll
    for (int m = 0; m < NRPM; ++m) {
      const int rp = grs + RS * (w + kFW * m);
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        double r0 = 0, r1 = 0;
        if (m < fg.nrp && rp < npairs) {
          const double* Rk = b.R + (int64_t)k * n_s + 2 * rp;
          r0 = __ldcg(Rk);
          r1 = (2 * rp + 1 < n) ? __ldcg(Rk + 1) : 0.0;
        }
        Rg[m][0][k] = r0;
        Rg[m][1][k] = r1;
      }
    }
    double la[kNlpMax][2][KT];
#pragma unroll
    for (int pp = 0; pp < kNlpMax; ++pp)
#pragma unroll
      for (int k = 0; k < KT; ++k) { la[pp][0][k] = 0; la[pp][1][k] = 0; }

    for (int ti = 0; ti < ntile; ++ti, ++q) {
      const int64_t qq = resident ? ti : q;
      const int stage = (int)(qq % fg.NS);
      if (!resident || s == 0) mbar_wait(&bars[stage], (uint32_t)((qq / fg.NS) & 1));
      const double* tile = ring + (size_t)stage * TT * S;
      const int r0 = c0 + ti * TT;
      const int ncol = min(TT, c1 - r0);

      // ---- gradient phase
      double acc[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) acc[k] = 0;
      if (gj < ncol) {
        const double* xc = tile + (size_t)gj * S;
#pragma unroll
        for (int m = 0; m < NRPM; ++m) {
          const int rp = grs + RS * (w + kFW * m);
          if (m < fg.nrp && rp < npairs) {
            double2 x = *reinterpret_cast<const double2*>(xc + 2 * rp);
            if (2 * rp + 1 >= n) x.y = 0;
#pragma unroll
            for (int k = 0; k < KT; ++k) acc[k] = fma(x.x, Rg[m][0][k], fma(x.y, Rg[m][1][k], acc[k]));
          }
        }
      }
#pragma unroll
      for (int o = TT; o < 32; o <<= 1)
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      if (lane < TT)
#pragma unroll
        for (int k = 0; k < KT; ++k) gw[(w * TT + lane) * KT + k] = acc[k];
      __syncthreads();

      // the ring slot consumed by the previous tile is free now: refill it (stream runs ahead across steps)
      if (tid == 0 && !resident && q >= 1) {
        const int64_t nxt = q - 1 + fg.NS;
        if (nxt < total_loads) issue(nxt);
      }

      // ---- post phase: prior gradient, MoveMomt of step s, UpdateMomtAndDeltas of step s+1
      if (tid < ncol) {
        const int r = r0 + tid;
        double* d = b.dc + (int64_t)r * KT;
        double* mo = b.momt + (int64_t)r * KT;
        const double sig = b.sigc[r], stp = b.stepc[r], hs = stp / 2;
        double dv[KT];
        double sd = 0;
#pragma unroll
        for (int k = 0; k < KT; ++k) { dv[k] = d[k]; sd = (k == 0) ? dv[0] : sd + dv[k]; }
        const double sdc = sd / Cd;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          double gl = 0;
#pragma unroll
          for (int ww = 0; ww < kFW; ++ww) gl += gw[(ww * TT + tid) * KT + k];
          const double post = gl + (dv[k] - sdc) / sig;
          double m = mo[k];
          const double kick = __dmul_rn(hs, post);
          if (s >= 1) m = __dsub_rn(m, kick);
          if (s < L) {
            m = __dsub_rn(m, kick);
            dv[k] = __dadd_rn(dv[k], __dmul_rn(stp, m));
            d[k] = dv[k];
          }
          mo[k] = m;
          dts[tid * KT + k] = dv[k];
        }
      }
      __syncthreads();

      // ---- lv phase: this block's partial of X[:, iup] * delta
      if (s < L) {
#pragma unroll
        for (int pp = 0; pp < kNlpMax; ++pp) {
          const int rp = tid + kFT * pp;
          if (pp < fg.nlp && rp < npairs) {
            for (int jj = 0; jj < ncol; ++jj) {
              const double2 x = *reinterpret_cast<const double2*>(tile + (size_t)jj * S + 2 * rp);
#pragma unroll
              for (int k = 0; k < KT; ++k) {
                const double c = dts[jj * KT + k];
                la[pp][0][k] = fma(x.x, c, la[pp][0][k]);
                la[pp][1][k] = fma(x.y, c, la[pp][1][k]);
              }
            }
          }
        }
      }
    }
    if (s == L) break;

    // ---- publish this block's lv partial
    if (cta < gact) {
#pragma unroll
      for (int pp = 0; pp < kNlpMax; ++pp) {
        const int rp = tid + kFT * pp;
        if (pp < fg.nlp && rp < npairs) {
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            double* o = b.lvpart + ((int64_t)cta * KT + k) * n_s + 2 * rp;
            *reinterpret_cast<double2*>(o) = make_double2(la[pp][0][k], la[pp][1][k]);
          }
        }
      }
    }
    grid.sync();

    // ---- reduce the partials in block order, softmax, residual, log-likelihood (rows split over all blocks)
    {
      const int row0 = cta * fg.rpc, nrows = max(0, min(n, row0 + fg.rpc) - row0);
      for (int item = w; item < nrows * KT; item += kFW) {
        const int ii = item / KT, k = item - ii * KT;
        const double* src = b.lvpart + (int64_t)k * n_s + row0 + ii;
        double a = 0;
        for (int sl = lane; sl < gact; sl += 32) a += __ldcg(src + (int64_t)sl * KT * n_s);
        a = warp_sum(a);
        if (lane == 0) dl[ii * KT + k] = b.lv_fix[(int64_t)k * n_s + row0 + ii] + a;
      }
      __syncthreads();
      double ll = 0;
      if (tid < nrows) {
        const int i = row0 + tid;
        double l[KT];
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k) { l[k] = dl[tid * KT + k]; m = fmax(m, l[k]); }
        double se = exp(0.0 - m);
#pragma unroll
        for (int k = 0; k < KT; ++k) se += exp(l[k] - m);
        const double lse = log(se) + m;
        const int yb = b.ybase[i];
        if (yb == 0) ll = 0.0 - lse;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const double nl = l[k] - lse;
          if (yb == k + 1) ll = nl;
          b.lv[(int64_t)k * n_s + i] = l[k];
          b.R[(int64_t)k * n_s + i] = exp(nl) - b.ymat[(int64_t)k * n_s + i];
        }
      }
      ll = block_sum(ll, redsm);
      if (tid == 0) b.red[cta].c = ll;
    }
    grid.sync();
  }

  if (cta == 0 && tid == 0) {
    double A = 0;
    for (int c = 0; c < G; ++c) A += __ldcg(&b.red[c].c);
    ctl->loglike_new = A;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
namespace {

template <int KT, int TT>
cudaError_t launch_one(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st) {
  static bool configured = false;
  auto kern = k_traj_fused<KT, TT>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fg.smem_bytes_max);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  Geom gg = g;
  Bufs bb = b;
  FusedGeom ff = fg;
  int LL = L;
  void* args[] = {&gg, &bb, &ff, &LL};
  return cudaLaunchCooperativeKernel((void*)kern, dim3(fg.grid), dim3(kFT), args, fg.smem_bytes, st);
}

template <int KT>
cudaError_t launch_k(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st) {
  switch (fg.T) {
    case 1: return launch_one<KT, 1>(g, b, fg, L, st);
    case 2: return launch_one<KT, 2>(g, b, fg, L, st);
    case 4: return launch_one<KT, 4>(g, b, fg, L, st);
    case 8: return launch_one<KT, 8>(g, b, fg, L, st);
    case 16: return launch_one<KT, 16>(g, b, fg, L, st);
    case 32: return launch_one<KT, 32>(g, b, fg, L, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

// Picks the tile shape for (n, K); returns false when the fused path does not apply (a column must fit
// in shared memory with room for a pipeline, K <= 4).
bool fused_plan(const Geom& g, int smem_optin_bytes, FusedGeom* out) {
  const int n = g.n, K = g.K;
  if (K < 1 || K > 4) return false;
  const int npairs = (n + 1) / 2;
  if (npairs > kFT * kNlpMax) return false;
  const int nrpm = K <= 2 ? 8 : 4;
  const int rpc = (n + g.sm_count - 1) / g.sm_count;
  for (int T = 32; T >= 1; T >>= 1) {
    const int RS = 32 / T;
    // half-stride h = S/2 (in 16-byte chunks) must spread 8 consecutive lanes over 8 distinct bank groups
    int h = npairs;
    if (T >= 8) { if (h % 2 == 0) h += 1; }
    else if (T == 4) { while (h % 4 != 2) ++h; }
    else if (T == 2) { while (h % 8 != 4) ++h; }
    const int S = 2 * h;
    const size_t tile_bytes = (size_t)T * S * sizeof(double);
    const size_t other = (size_t)(kFW * T * K + T * K + 32 + rpc * K) * sizeof(double) + 8 * sizeof(uint64_t) + 256;
    if (tile_bytes > (T == 1 ? 100 * 1024 : 48 * 1024)) continue;
    const int nrp = (npairs + RS * kFW - 1) / (RS * kFW);
    if (nrp > nrpm) continue;
    const size_t budget = (size_t)smem_optin_bytes - other - 1024;
    int NS = (int)std::min<size_t>(8, budget / tile_bytes);
    if (NS < 2) continue;
    out->T = T; out->S = S; out->NS = NS; out->nrp = nrp; out->nlp = (npairs + kFT - 1) / kFT; out->rpc = rpc;
    out->grid = g.sm_count;
    out->smem_bytes = (size_t)NS * tile_bytes + other;
    out->smem_bytes_max = (size_t)smem_optin_bytes;
    return true;
  }
  return false;
}

cudaError_t launch_traj_fused(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st) {
  switch (g.K) {
    case 1: return launch_k<1>(g, b, fg, L, st);
    case 2: return launch_k<2>(g, b, fg, L, st);
    case 3: return launch_k<3>(g, b, fg, L, st);
    case 4: return launch_k<4>(g, b, fg, L, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace htlr
