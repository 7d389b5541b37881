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
// Structure: grid = one block per SM (cooperative launch), 18 warps with fixed roles that hand tiles to
// each other through mbarriers only (no block-wide barrier in the steady state):
//   warp 0      producer — streams this block's column tiles (T columns x n rows) into a shared-memory
//               ring with cp.async.bulk (1-D TMA copies, one lane per column, completion counted on the
//               stage's `full` mbarrier). It runs ahead across leapfrog steps, so the ring is full again
//               when the block comes back from the grid-wide reduction; if the block's whole column range
//               fits in the ring it is loaded once and stays resident for the whole trajectory.
//   warp 1      post — for each tile sums the 16 per-warp gradient partials in fixed order, adds the prior
//               gradient, applies MoveMomt / UpdateMomtAndDeltas (un-fused products) and publishes the
//               tile's new coefficients.
//   warps 2-17  compute — gradient phase: lane <-> (column of the tile, row subgroup), 128-bit
//               conflict-free transposed reads (column stride chosen so 8 consecutive lanes hit 8 distinct
//               16-byte bank groups), the residual rows a thread needs live in registers for the sweep;
//               lv phase (one tile behind, so the post warp is off the critical path): thread <-> two
//               consecutive rows, loop over the tile's columns.
// After each sweep the per-block lv partials are summed in block order (deterministic), rows are
// soft-maxed and the new residual is published: two grid-wide barriers per leapfrog step.
#include <algorithm>
#include <cstdio>

#include "common.cuh"
#include "launch.h"

namespace htlr {

namespace {

constexpr int kCW = 16;                 // compute warps
constexpr int kCT = kCW * 32;           // compute threads
constexpr int kFT = kCT + 64;           // + producer warp + post warp
constexpr int kNlpMax = 2;              // row pairs per compute thread in the lv phase (n <= 2048)
constexpr int kPartSlots = 5;           // lv partials one lane sums in the row phase: grids of up to 160 blocks

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier over all kFT threads of every block (blocks are co-resident: cooperative launch).
// `counter` only ever grows during one launch; it is zeroed by k_begin before every trajectory.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& gen, unsigned nblocks) {
  named_bar(1, kFT);  // orders every thread's prior writes before thread 0's release below
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    const unsigned target = (gen + 1) * nblocks;
    while (ld_acquire(counter) < target) { __nanosleep(32); }
  }
  named_bar(1, kFT);
  ++gen;
}

template <int KT>
struct NrpMax {
  static constexpr int v = 4;  // row pairs per compute thread in the gradient phase (registers: 2*4*K doubles)
};

}  // namespace

template <int KT, int TT>
__global__ void __launch_bounds__(kFT, 1) k_traj_fused(Geom g, Bufs b, FusedGeom fg, int L) {
  constexpr int RS = 32 / TT;            // row subgroups per warp in the gradient phase
  constexpr int NRPM = NrpMax<KT>::v;    // row pairs a thread covers in the gradient phase (max)
  Ctl* ctl = b.ctl;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = fg.S, NS = fg.NS;
  double* ring = reinterpret_cast<double*>(smem_raw);                      // [NS][TT][S]
  double* gw = ring + (size_t)NS * TT * S;                                  // [2][kCW][TT*KT]
  double* dts = gw + 2 * kCW * TT * KT;                                     // [2][TT*KT]
  double* redsm = dts + 2 * TT * KT;                                        // [kCW]
  double* dl = redsm + kCW;                                                 // [rpc*KT]
  double* cst = dl + fg.rpc * KT;                     // per-column state of this block: [ncmax][2*KT+2] = d, m, sig, step
  uint64_t* bars = reinterpret_cast<uint64_t*>(cst + (size_t)fg.ncmax * (2 * KT + 2));
  uint64_t* full = bars;            // [NS]  producer -> compute
  uint64_t* empty = bars + NS;      // [NS]  compute  -> producer
  uint64_t* gdone = empty + NS;     // [2]   compute  -> post
  uint64_t* pdone = gdone + 2;      // [2]   post     -> compute
  int* cidx = reinterpret_cast<int*>(pdone + 2);      // [ncmax] X column of each owned restricted-set position

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned G = gridDim.x;
  const int cta = blockIdx.x;
  const int n = g.n, n_s = g.n_s, npairs = (n + 1) >> 1;
  const int u = ctl->u;
  unsigned* gcounter = &ctl->ticket[4];
  unsigned gen = 0;

  // ---- column ownership: whole tiles, contiguous, only as many blocks as there are tiles
  const int tiles_total = (u + TT - 1) / TT;
  const int gact = min((int)G, tiles_total);
  int t0 = 0, t1 = 0;
  if (cta < gact) {
    t0 = (int)((int64_t)cta * tiles_total / gact);
    t1 = (int)((int64_t)(cta + 1) * tiles_total / gact);
  }
  const int ntile = t1 - t0;
  const int c0 = t0 * TT, c1 = min(u, t1 * TT);
  const bool resident = ntile <= NS;
  const int64_t total_loads = resident ? ntile : (int64_t)ntile * (L + 1);
  const uint32_t col_bytes = (uint32_t)(((n + 1) >> 1) << 4);  // n rounded up to even, in bytes

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kCW); }
    for (int i = 0; i < 2; ++i) { mbar_init(&gdone[i], kCW); mbar_init(&pdone[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (cta == 0) ctl->cols_swept += (unsigned long long)u * (unsigned long long)(L + 1);
  }
  // this block's column indices and per-column state live in shared memory for the whole trajectory
  constexpr int CS = 2 * KT + 2;
  for (int e = tid; e < c1 - c0; e += kFT) {
    const int r = c0 + e;
    cidx[e] = b.iup[r];
    double* st = cst + (size_t)e * CS;
#pragma unroll
    for (int k = 0; k < KT; ++k) { st[k] = b.dc[(int64_t)r * KT + k]; st[KT + k] = b.momt[(int64_t)r * KT + k]; }
    st[2 * KT] = b.sigc[r];
    st[2 * KT + 1] = b.stepc[r];
  }
  __syncthreads();

  if (warp == 0) {
    // ================================ producer ================================
    int64_t issued = 0;
    int stage = 0, ti = 0;
    uint32_t eph = 1;  // parity to wait for on empty[stage]; flips each time the ring wraps (first lap: no wait)
    for (int s = 0; s <= L; ++s) {
      int64_t limit;
      if (resident) limit = ntile;
      else {
        limit = (int64_t)(s + 1) * ntile + (s < L ? min(NS, ntile) : 0);
        if (limit > total_loads) limit = total_loads;
      }
      for (; issued < limit; ++issued) {
        if (issued >= NS) mbar_wait(&empty[stage], eph);
        const int r0 = c0 + ti * TT;
        const int ncol = min(TT, c1 - r0);
        if (lane == 0) mbar_expect_tx(&full[stage], col_bytes * (uint32_t)ncol);
        __syncwarp();
        if (lane < ncol) {
          const int j = cidx[r0 - c0 + lane];
          bulk_g2s(ring + ((size_t)stage * TT + lane) * S, b.X + (int64_t)j * g.ld, col_bytes, &full[stage]);
        }
        if (++ti == ntile) ti = 0;
        if (++stage == NS) { stage = 0; eph ^= 1; }
      }
      if (s < L) { grid_barrier(gcounter, gen, G); grid_barrier(gcounter, gen, G); }
    }
  } else if (warp == 1) {
    // ================================ post ================================
    // lane <-> (column jj, class k, quarter qt of the compute warps): 4 partial sums each, combined in a
    // fixed order with two shuffles, so the 16-warp gradient sum costs ~6 dependent adds, not 16.
    const double Cd = (double)ctl->C;
    constexpr int PER = TT * KT;                 // (column, class) pairs per tile
    constexpr int ROUNDS = (PER * 4 + 31) / 32;  // 8 pairs per round
    int64_t tc = 0;  // tiles processed by this block so far (parity of gdone / pdone)
    for (int s = 0; s <= L; ++s) {
      for (int ti = 0; ti < ntile; ++ti, ++tc) {
        const int bsel = (int)(tc & 1);
        const uint32_t par = (uint32_t)((tc >> 1) & 1);
        const int e0 = ti * TT;                 // first owned column of the tile (block-local)
        const int ncol = min(TT, (c1 - c0) - e0);
        mbar_wait(&gdone[bsel], par);
        const double* gwb = gw + (size_t)bsel * kCW * PER;
        double* dtb = dts + bsel * PER;
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
          const int pr = rd * 8 + (lane >> 2), qt = lane & 3;   // pair index, quarter
          const int jj = pr / KT, k = pr - jj * KT;
          double gl = 0;
          if (pr < PER) {
#pragma unroll
            for (int ww = 0; ww < 4; ++ww) gl += gwb[(qt * 4 + ww) * PER + pr];
          }
          gl += __shfl_xor_sync(0xffffffffu, gl, 1);
          gl += __shfl_xor_sync(0xffffffffu, gl, 2);
          if (pr < PER && jj < ncol && qt == 0) {
            double* st = cst + (size_t)(e0 + jj) * CS;
            double sd = st[0];
#pragma unroll
            for (int kk = 1; kk < KT; ++kk) sd += st[kk];
            const double sig = st[2 * KT], stp = st[2 * KT + 1], hs = stp / 2;
            const double dold = st[k];
            const double post = gl + (dold - sd / Cd) / sig;
            double m = st[KT + k];
            const double kick = __dmul_rn(hs, post);
            if (s >= 1) m = __dsub_rn(m, kick);
            double dnew = dold;
            if (s < L) {
              m = __dsub_rn(m, kick);
              dnew = __dadd_rn(dold, __dmul_rn(stp, m));
            }
            dtb[pr] = dnew;
            st[KT + k] = m;  // momenta are private to (column, class); deltas are committed below
          }
        }
        __syncwarp();
        // commit the new deltas of this tile to the column state (after every lane has read the old ones)
        for (int pr = lane; pr < ncol * KT; pr += 32) {
          const int jj = pr / KT, k = pr - jj * KT;
          cst[(size_t)(e0 + jj) * CS + k] = dtb[pr];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pdone[bsel]);
      }
      if (s < L) { grid_barrier(gcounter, gen, G); grid_barrier(gcounter, gen, G); }
    }
    // final coefficients and momenta back to the compact global arrays
    for (int e = lane; e < c1 - c0; e += 32) {
      const int r = c0 + e;
      const double* st = cst + (size_t)e * CS;
#pragma unroll
      for (int k = 0; k < KT; ++k) { b.dc[(int64_t)r * KT + k] = st[k]; b.momt[(int64_t)r * KT + k] = st[KT + k]; }
    }
  } else {
    // ================================ compute ================================
    const int ct = tid - 64, w = warp - 2;
    const int gj = lane % TT, grs = lane / TT;
    const int tile_stride = TT * S;
    // per-thread constants of the two phases: shared-memory offsets and validity of each row pair
    int goff[NRPM];
    bool gval[NRPM];
#pragma unroll
    for (int m = 0; m < NRPM; ++m) {
      const int rp = grs + RS * (w + kCW * m);
      gval[m] = (m < fg.nrp) && (rp < npairs);
      goff[m] = gj * S + 2 * rp;
    }
    int loff[kNlpMax];
    bool lval[kNlpMax];
#pragma unroll
    for (int pp = 0; pp < kNlpMax; ++pp) {
      const int rp = ct + kCT * pp;
      lval[pp] = (pp < fg.nlp) && (rp < npairs);
      loff[pp] = 2 * rp;
    }
    double Rg[NRPM][2][KT];
    // row phase: this thread's row of the block's row slice; its fixed part, label and one-hot row never
    // change during a trajectory, so they are read once
    const int row0 = cta * fg.rpc, nrows = max(0, min(n, row0 + fg.rpc) - row0);
    double row_fix[KT], row_y[KT];
    int row_yb = -1;
#pragma unroll
    for (int k = 0; k < KT; ++k) { row_fix[k] = 0; row_y[k] = 0; }
    if (ct < nrows) {
      row_yb = b.ybase[row0 + ct];
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        row_fix[k] = b.lv_fix[(int64_t)k * n_s + row0 + ct];
        row_y[k] = b.ymat[(int64_t)k * n_s + row0 + ct];
      }
    }
    const bool timer = (cta == 0 && tid == 64);
    const bool timer2 = (cta == 100 && tid == 64);
    long long dbg[8] = {0,0,0,0,0,0,0,0}, td = 0;
    long long tk0 = 0, cyc_sweep = 0, cyc_b1 = 0, cyc_row = 0, cyc_b2 = 0;
    int tc = 0;               // tile counter (gdone / pdone parity), same sequence as the post warp's
    int stage = 0;            // ring position of the tile whose gradient phase runs next
    uint32_t fph = 0;         // parity to wait for on full[stage]
    for (int s = 0; s <= L; ++s) {
      if (timer) tk0 = clock64();
      // residual rows of this thread into registers (published by the previous softmax phase)
#pragma unroll
      for (int m = 0; m < NRPM; ++m) {
        const int rp = grs + RS * (w + kCW * m);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          double r0v = 0, r1v = 0;
          if (gval[m]) {
            const double* Rk = b.R + (int64_t)k * n_s + 2 * rp;
            r0v = __ldcg(Rk);
            r1v = (2 * rp + 1 < n) ? __ldcg(Rk + 1) : 0.0;   // pad row of an odd n contributes nothing
          }
          Rg[m][0][k] = r0v;
          Rg[m][1][k] = r1v;
        }
      }
      double la[kNlpMax][2][KT];
#pragma unroll
      for (int pp = 0; pp < kNlpMax; ++pp)
#pragma unroll
        for (int k = 0; k < KT; ++k) { la[pp][0][k] = 0; la[pp][1][k] = 0; }
      if (resident) { stage = 0; }

      // lv phase of one tile (runs one tile behind the gradient phase)
      auto lv_phase = [&](int ti_prev, int tc_prev, int stage_prev) {
        const int bsel = tc_prev & 1;
        mbar_wait(&pdone[bsel], (uint32_t)((tc_prev >> 1) & 1));
        if (s < L) {
          const double* tile = ring + (size_t)stage_prev * tile_stride;
          const double* dtb = dts + bsel * TT * KT;
          const int ncol = min(TT, c1 - (c0 + ti_prev * TT));
#pragma unroll
          for (int pp = 0; pp < kNlpMax; ++pp) {
            if (lval[pp]) {
#pragma unroll
              for (int jj = 0; jj < TT; ++jj) {
                if (jj < ncol) {
                  const double2 x = *reinterpret_cast<const double2*>(tile + jj * S + loff[pp]);
#pragma unroll
                  for (int k = 0; k < KT; ++k) {
                    const double c = dtb[jj * KT + k];
                    la[pp][0][k] = fma(x.x, c, la[pp][0][k]);
                    la[pp][1][k] = fma(x.y, c, la[pp][1][k]);
                  }
                }
              }
            }
          }
        }
        if (!resident) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage_prev]);
        }
      };

      int stage_prev = 0;
      for (int ti = 0; ti < ntile; ++ti, ++tc) {
        if (!resident || s == 0) mbar_wait(&full[stage], fph);
        const double* tile = ring + (size_t)stage * tile_stride;
        const int ncol = min(TT, c1 - (c0 + ti * TT));
        // ---- gradient phase
        double acc[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[k] = 0;
        if (gj < ncol) {
#pragma unroll
          for (int m = 0; m < NRPM; ++m) {
            if (gval[m]) {
              const double2 x = *reinterpret_cast<const double2*>(tile + goff[m]);
#pragma unroll
              for (int k = 0; k < KT; ++k) acc[k] = fma(x.x, Rg[m][0][k], fma(x.y, Rg[m][1][k], acc[k]));
            }
          }
        }
#pragma unroll
        for (int o = TT; o < 32; o <<= 1)
#pragma unroll
          for (int k = 0; k < KT; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        const int bsel = tc & 1;
        if (lane < TT) {
          double* gwb = gw + ((size_t)bsel * kCW + w) * TT * KT;
#pragma unroll
          for (int k = 0; k < KT; ++k) gwb[lane * KT + k] = acc[k];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&gdone[bsel]);
        // ---- lv phase of the previous tile
        if (ti > 0) lv_phase(ti - 1, tc - 1, stage_prev);
        stage_prev = stage;
        if (++stage == NS) { stage = 0; fph ^= 1; }
      }
      if (ntile > 0) lv_phase(ntile - 1, tc - 1, stage_prev);
      if (timer) { const long long t = clock64(); cyc_sweep += t - tk0; tk0 = t; }
      if (s == L) break;

      // ---- publish this block's lv partial
      if (cta < gact) {
#pragma unroll
        for (int pp = 0; pp < kNlpMax; ++pp) {
          const int rp = ct + kCT * pp;
          if (pp < fg.nlp && rp < npairs) {
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              double* o = b.lvpart + ((int64_t)cta * KT + k) * n_s + 2 * rp;
              *reinterpret_cast<double2*>(o) = make_double2(la[pp][0][k], la[pp][1][k]);
            }
          }
        }
      }
      grid_barrier(gcounter, gen, G);
      if (timer) { const long long t = clock64(); cyc_b1 += t - tk0; tk0 = t; }
      if (timer || timer2) td = clock64();

      // ---- reduce the partials in block order, softmax, residual, log-likelihood (rows split over all blocks)
      {
        // one (row, class) item per warp pass; all of a lane's partials are requested before any is summed,
        // so the pass costs one L2 round trip instead of one per 32 blocks
        for (int item = w; item < nrows * KT; item += kCW) {
          const int ii = item / KT, k = item - ii * KT;
          const double* src = b.lvpart + (int64_t)k * n_s + row0 + ii;
          double pv[kPartSlots];
#pragma unroll
          for (int q = 0; q < kPartSlots; ++q) {
            const int sl = lane + 32 * q;
            pv[q] = (sl < gact) ? __ldcg(src + (int64_t)sl * KT * n_s) : 0.0;
          }
          double a = pv[0];
#pragma unroll
          for (int q = 1; q < kPartSlots; ++q) a += pv[q];
          a = warp_sum(a);
          if (lane == 0) dl[ii * KT + k] = a;
        }
        if (timer || timer2) { const long long t = clock64(); dbg[0] += t - td; td = t; }
        named_bar(2, kCT);
        if (timer || timer2) { const long long t = clock64(); dbg[1] += t - td; td = t; }
        double ll = 0;
        if (ct < nrows) {
          const int i = row0 + ct;
          double l[KT];
          double m = 0.0;
#pragma unroll
          for (int k = 0; k < KT; ++k) { l[k] = row_fix[k] + dl[ct * KT + k]; m = fmax(m, l[k]); }
          double se = exp(0.0 - m);
#pragma unroll
          for (int k = 0; k < KT; ++k) se += exp(l[k] - m);
          const double lse = log(se) + m;
          if (row_yb == 0) ll = 0.0 - lse;
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const double nl = l[k] - lse;
            if (row_yb == k + 1) ll = nl;
            b.lv[(int64_t)k * n_s + i] = l[k];
            __stcg(&b.R[(int64_t)k * n_s + i], exp(nl) - row_y[k]);
          }
        }
        ll = warp_sum(ll);
        if (timer || timer2) { const long long t = clock64(); dbg[2] += t - td; td = t; }
        if (lane == 0) redsm[w] = ll;
        named_bar(2, kCT);
        if (timer || timer2) { const long long t = clock64(); dbg[3] += t - td; td = t; }
        if (ct == 0) {
          double t = 0;
          for (int i = 0; i < kCW; ++i) t += redsm[i];
          b.red[cta].c = t;
        }
      }
      if (timer) { const long long t = clock64(); cyc_row += t - tk0; tk0 = t; }
      grid_barrier(gcounter, gen, G);
      if (timer) { const long long t = clock64(); cyc_b2 += t - tk0; }
      if (timer || timer2) { const long long t = clock64(); dbg[4] += t - td; td = t; }
    }
    if (timer || timer2) for (int q = 0; q < 8; ++q) ctl->ph_dbg[q + (timer2 ? 8 : 0)] += (unsigned long long)dbg[q];
    if (timer) {
      ctl->ph_cycles[0] += (unsigned long long)cyc_sweep;
      ctl->ph_cycles[1] += (unsigned long long)cyc_b1;
      ctl->ph_cycles[2] += (unsigned long long)cyc_row;
      ctl->ph_cycles[3] += (unsigned long long)cyc_b2;
    }
  }

  // every block has passed the last grid barrier, so all log-likelihood partials are visible
  if (cta == 0 && tid == 64) {
    double A = 0;
    for (unsigned c = 0; c < G; ++c) A += __ldcg(&b.red[c].c);
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
  if (npairs > kCT * kNlpMax) return false;
  const int nrpm = 4;
  const int rpc = (n + g.sm_count - 1) / g.sm_count;
  if (rpc > kCT) return false;
  if (g.sm_count > 32 * kPartSlots) return false;   // row phase: one lane sums at most kPartSlots block partials
  for (int T = 32; T >= 1; T >>= 1) {
    const int RS = 32 / T;
    // half-stride h = S/2 (in 16-byte chunks) must spread 8 consecutive lanes over 8 distinct bank groups
    int h = npairs;
    if (T >= 8) { if (h % 2 == 0) h += 1; }
    else if (T == 4) { while (h % 4 != 2) ++h; }
    else if (T == 2) { while (h % 8 != 4) ++h; }
    const int S = 2 * h;
    const size_t tile_bytes = (size_t)T * S * sizeof(double);
    // worst case (every feature in the restricted set): tiles per block, rounded up, times T columns
    const int tiles_all = (g.nvar + T - 1) / T;
    const int ncmax = ((tiles_all + g.sm_count - 1) / g.sm_count) * T;
    const size_t other = (size_t)(2 * kCW * T * K + 2 * T * K + kCW + rpc * K) * sizeof(double) +
                         (size_t)ncmax * ((2 * K + 2) * sizeof(double) + sizeof(int)) +
                         (2 * 8 + 4) * sizeof(uint64_t) + 256;
    if (other > 96 * 1024) continue;
    if (tile_bytes > (T == 1 ? 100 * 1024 : 48 * 1024)) continue;
    const int nrp = (npairs + RS * kCW - 1) / (RS * kCW);
    if (nrp > nrpm) continue;
    if ((size_t)smem_optin_bytes < other + 1024 + 2 * tile_bytes) continue;
    const size_t budget = (size_t)smem_optin_bytes - other - 1024;
    int NS = (int)std::min<size_t>(8, budget / tile_bytes);
    if (NS < 2) continue;
    out->T = T; out->S = S; out->NS = NS; out->nrp = nrp; out->nlp = (npairs + kCT - 1) / kCT; out->rpc = rpc;
    out->grid = g.sm_count;
    out->ncmax = ncmax;
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
