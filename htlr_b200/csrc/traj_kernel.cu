// Fused leapfrog trajectory: ONE persistent kernel runs all L+1 gradient evaluations of an HMC proposal,
// reading each restricted-set column of X once per leapfrog step (the reference reads it twice:
// UpdatePredProb src/gibbs.cpp:179-188 and UpdateDNlogLike src/gibbs.cpp:239-249).
//
// Why one sweep suffices: MoveMomt of step t (gibbs.cpp:467-471) and UpdateMomtAndDeltas of step t+1
// (gibbs.cpp:291-297) use the same DNlogpost, and both the gradient of column j and column j's
// contribution to the next linear predictor need only column j plus the n x K residual. So while a
// column sits in shared memory (and, within a thread, in registers) its owners compute g_j = X_j' r, apply
// the two half kicks as two separate subtractions (bit-compatible with the reference), drift delta_j, and
// accumulate X_j (delta_j,new - delta_j,old) into their private partial of the CHANGE of lv: the predictor itself is
// carried from step to step by the row owners, and lv / the residual of the current point are kept across
// iterations (k_begin saves them, a rejection restores them), so there is no DetachFixlv sweep
// (gibbs.cpp:197-229) and no softmax before the first gradient. Algorithmic X traffic per iteration:
// 8 n (L+1) u bytes.
//
// Structure: 16 warps per block (4 per SM sub-partition, so a thread may use 128 registers), organised as
// NG = 16/GW independent GROUPS of GW warps. A group owns whole columns (round robin over the block's
// columns) and its own slice of the shared-memory ring:
//   * one lane of the group (the duty rotates over its warps) streams the group's next columns in with
//     cp.async.bulk (1-D TMA copies, completion counted on the stage's `full` mbarrier), SG-1 columns ahead of
//     the one being consumed; the stream simply continues into the next leapfrog step, so the ring is full
//     again when the block comes back from the grid-wide reduction. If the group's columns all fit in its
//     ring they are loaded once and stay resident for the whole trajectory.
//   * thread <-> NRP row pairs of the column, read once with unit-stride 128-bit shared-memory loads and kept
//     in registers for both uses (the stage is handed back as soon as the loads have landed); the residual
//     rows a thread needs live in registers for the sweep. Per column: dot products -> xor-shuffle tree ->
//     (GW > 1) per-warp partials through shared memory and one named barrier of the group -> every thread
//     forms the same gradient, prior gradient, MoveMomt / UpdateMomtAndDeltas (un-fused products) ->
//     X_j (delta_new - delta_old) accumulates into the thread's lv registers.
//   Groups never wait for each other inside a sweep, so up to NG columns are in flight per SM; GW is the
//   smallest group that keeps a thread's row pairs in registers.
// After each sweep the group partials are summed in shared memory, the per-block lv partials are summed in
// block order (deterministic) and added to the rows' running lv, rows are soft-maxed (K >= 3: one lane per class)
// and the new residual is published: two grid-wide barriers per leapfrog step in grid mode. The log-likelihood is
// reduced once, after the last step.
//
// Two flavours (template flag CL):
//   grid mode    — one block per SM, cooperative launch; partials and residual travel through L2 and the two
//                  barriers per step are release/acquire counters in global memory (~1.2 us each). Right when
//                  the restricted set is large (HBM-bound).
//   cluster mode — ONE thread-block cluster (16 CTAs, or 8), everything through distributed shared memory and no
//                  barrier inside the trajectory: a block pushes its lv partial to the owners of the rows and the
//                  owners push the new residual to every block as st.async stores that count themselves in on
//                  a transaction barrier of the receiver (a cluster barrier's release is a MEMBAR.ALL.GPU: an L2
//                  round trip for data that never leaves shared memory). Right for small restricted sets (the
//                  product default, cut = 0.05), where a leapfrog step is pure latency.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "launch.h"

namespace htlr {

namespace {

constexpr int kCW = 16;                 // compute warps
constexpr int kCT = kCW * 32;           // compute threads
constexpr int kFT = kCT;                // every warp computes; TMA copies are issued by the groups themselves
constexpr int kNlpMax = 2;              // row pairs per compute thread when publishing the lv partial (n <= 2048)
constexpr int kPartSlots = 5;           // grids of up to 160 blocks (32 * kPartSlots)
constexpr int kWriterSlots = 10;        // publishing blocks one warp covers in the row phase: 16 * 10 = 160
constexpr int kClMax = 16;              // largest cluster (non-portable size; 8 is the portable fallback)
constexpr int kNsMax = 64;              // ring stages (one column each)
constexpr int kBarGroup = 1;            // named barriers 1..4: the compute groups
constexpr int kBarCompute = 5;          // all compute threads
constexpr int kBarAll = 6;              // every thread of the block (grid barrier)

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
  named_bar(kBarAll, kFT);  // orders every thread's prior writes before thread 0's release below
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    const unsigned target = (gen + 1) * nblocks;
    while (ld_acquire(counter) < target) { __nanosleep(32); }
  }
  named_bar(kBarAll, kFT);
  ++gen;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Asynchronous stores into another block's shared memory that count themselves in on a transaction barrier of the
// RECEIVING block (addresses already mapped with mapa): the data and the signal travel together, so a cluster-mode
// exchange needs no fence and no cluster barrier. (barrier.cluster.arrive.release costs a MEMBAR.ALL.GPU, an L2 round
// trip, for data that never leaves shared memory: 0.35 us per barrier, two per leapfrog step.)
__device__ __forceinline__ void st_async_f64(uint32_t addr, double v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];"
               ::"r"(addr), "d"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_2f64(uint32_t addr, double v0, double v1, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
               ::"r"(addr), "d"(v0), "d"(v1), "r"(mbar) : "memory");
}
__device__ __forceinline__ long long clk() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}

}  // namespace

// Columns per tile: at most 8 (column, class) values per tile, and at most 8 double2 of X per thread. With one class and
// one row pair per thread (K = 1, n <= 64 GW) that can be 8 columns: half as many tiles, i.e. half as many trips through
// the per-tile barriers and reductions, for the binary problems whose sweeps are bound by those (config D, n = 200:
// 4520 -> 5480 it/s). Not for one-warp groups (n <= 64: 16 groups per block): with a dozen columns per block the larger
// tiles keep fewer groups busy (config A: 126 -> 137 us per iteration, measured).
template <int KT, int NRP, int GW>
struct TileCols {
  static constexpr int v = (KT == 1 && NRP == 1 && GW >= 4) ? 8 : (KT <= 2 ? 4 : (KT <= 4 ? 2 : 1));
};

template <int KT, int NRP, int GW, bool CL>
__global__ void __launch_bounds__(kFT, 1) k_traj(Geom g, Bufs b, FusedGeom fg, int L) {
  constexpr int NG = kCW / GW;           // compute groups
  constexpr int GT = 32 * GW;            // threads per group
  constexpr int CS = 3 * KT + 2;         // per-column state: d[K], m[K], prior gradient[K], sig, step
  constexpr int TC = TileCols<KT, NRP, GW>::v;    // columns per tile (= ring stage)
  constexpr int V = TC * KT;             // (column, class) gradient values per tile
  constexpr int VP = V <= 4 ? 4 : 8;     // padded to a power of two for the transposing reduction
  Ctl* ctl = b.ctl;
  timeline_mark(ctl, CL ? 5 : 6);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = g.n, n_s = g.n_s, npairs = (n + 1) >> 1, n_e = 2 * npairs;
  const int SG = fg.NS;                  // ring stages (tiles) per group
  double* ring = reinterpret_cast<double*>(smem_raw);                       // [NG][SG][TC][n_e]
  double* gpart = ring + (size_t)NG * SG * TC * n_e;                         // [2][kCW][VP] per-warp gradient partials
  double* dnsm = gpart + 2 * kCW * VP;                                       // [2][NG][VP] new coefficients of a tile
  double* lasm = dnsm + 2 * NG * VP;                                         // [NG-1][KT][n_e] lv partials of groups 1.. in transit
  double* redsm = lasm + (size_t)(NG - 1) * KT * n_e;                        // [kCW]
  double* dl = redsm + kCW;                                                  // [rpc*KT]
  double* wsum = dl + ((fg.rpc * KT + 1) & ~1);                              // [kCW][rpc*KT] per-warp sums of the partials (grid mode)
  double* rowc = wsum + (CL ? 0 : (size_t)kCW * fg.rpc * KT);                // [rpc][2*KT+2] running lv, ymat, label
  double* cst = rowc + (size_t)fg.rpc * (2 * KT + 2);                        // [ncmax][CS]
  double* prcv = cst + (((size_t)fg.ncmax * CS + 1) & ~(size_t)1);           // cluster mode: lv partials RECEIVED for this block's rows [cl][KT][rpc]
  double* Rs = prcv + (CL ? (size_t)fg.cl * KT * fg.rpc : 0);                // cluster mode: residual copy [KT][n_e]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Rs + (CL ? (size_t)KT * n_e : 0));
  uint64_t* full = bars;                  // [NG*SG]  TMA -> group
  uint64_t* xbar = full + NG * SG;        // [2]      cluster mode: partials received, residuals received
  int* cidx = reinterpret_cast<int*>(xbar + 2);          // [ncmax] X column of each owned restricted-set position

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned G = gridDim.x;
  const int cta = blockIdx.x;
  const int u = ctl->u;
  unsigned* gcounter = &ctl->ticket[4];
  unsigned gen = 0;
  if (ctl->traj_done) return;   // an earlier launch of this thin-step (single-block or cluster kernel) ran the trajectory
  if (CL) {
    // restricted set too large for one cluster (state capacity, or past the size where grid mode is faster):
    // leave the proposal to the grid-mode launch that follows
    if (u > fg.cl_umax) {
      if (fg.cl_only && cta == 0 && tid == 0) set_err(ctl, kErrClusterCap, u);
      return;
    }
  }
  auto sync_all = [&]() {
    if (CL) cluster_barrier(); else grid_barrier(gcounter, gen, G);
  };

  // ---- column ownership: contiguous ranges, only as many blocks as there are columns
  const int gact = min((int)G, u);
  int c0 = 0, c1 = 0;
  if (cta < gact) {
    c0 = (int)((int64_t)cta * u / gact);
    c1 = (int)((int64_t)(cta + 1) * u / gact);
  }
  const int ncols = c1 - c0;
  const uint32_t col_bytes = (uint32_t)npairs << 4;  // n rounded up to even, in bytes
  const double Cd = (double)ctl->C;

  if (tid == 0) {
    for (int i = 0; i < NG * SG; ++i) mbar_init(&full[i], 1);
    if (CL) { mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (cta == 0) ctl->cols_swept += (unsigned long long)u * (unsigned long long)(L + 1);
  }
  // UpdateDNlogPrior / sigmasbt (gibbs.cpp:267-272, 281) depends only on the coefficients, so it is formed once
  // per leapfrog step for all of the block's columns, off the per-tile critical path
  auto prior_gradient = [&](int e) {
    double* st = cst + (size_t)e * CS;
    double sd = st[0];
#pragma unroll
    for (int k = 1; k < KT; ++k) sd += st[k];
    const double sdc = sd / Cd, sig = st[3 * KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) st[2 * KT + k] = (st[k] - sdc) / sig;
  };
  // the ring starts as zeros (the unconditional tile reads below rely on every stage holding finite values)
  for (size_t e = tid; e < (size_t)NG * SG * TC * n_e; e += kFT) ring[e] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zeros before the TMA writes
  // this block's column indices and per-column state live in shared memory for the whole trajectory
  for (int e = tid; e < ncols; e += kFT) {
    const int r = c0 + e;
    cidx[e] = b.iup[r];
    double* st = cst + (size_t)e * CS;
#pragma unroll
    for (int k = 0; k < KT; ++k) { st[k] = b.dc[(int64_t)r * KT + k]; st[KT + k] = b.momt[(int64_t)r * KT + k]; }
    st[3 * KT] = b.sigc[r];
    st[3 * KT + 1] = b.stepc[r];
    prior_gradient(e);
  }
  if (CL) {
    for (int e = tid; e < KT * n_e; e += kFT) {
      const int k = e / n_e, i = e - k * n_e;
      Rs[e] = i < n ? b.R[(int64_t)k * n_s + i] : 0.0;
    }
  }
  __syncthreads();
  // Cluster mode exchanges the lv partials and the residual by asynchronous stores that complete transaction barriers of
  // the receiving block (xbar[0]: partials of this block's rows from every block, xbar[1]: the residual of every row):
  // bytes expected per leapfrog step, armed one phase ahead by thread 0. No cluster barrier inside the trajectory; this
  // one makes every block's barriers (and its residual copy) exist before anybody stores into them.
  const int xrows = CL ? max(0, min(n_e, (cta + 1) * fg.rpc) - cta * fg.rpc) : 0;   // rows of this block's slice, pad row included
  const uint32_t xbytes0 = (uint32_t)(gact * xrows * KT) * 8u, xbytes1 = (uint32_t)(n * KT) * 8u;
  if (CL) {
    if (tid == 0) { mbar_expect_tx(&xbar[0], xbytes0); mbar_expect_tx(&xbar[1], xbytes1); }
    cluster_barrier();
  }

  const int ct = tid;
  const int grp = w / GW, wg = w % GW;
  const int q = wg * 32 + lane;                 // thread index within the group
  // the row pairs this thread covers in every column
  int roff[NRP];
  bool rval[NRP];
#pragma unroll
  for (int m = 0; m < NRP; ++m) {
    const int rp = q + GT * m;
    rval[m] = rp < npairs;
    roff[m] = rval[m] ? 2 * rp : 0;
  }
  // ---- this group's tile stream: the block's columns are cut into tiles of TC; the group takes tiles
  // grp, grp + NG, ...; the sequence repeats every leapfrog step. Tile i of the stream sits in stage i % SG of
  // the group's ring; the copies for tile i + SG are issued into tile i's own stage as soon as tile i sits in
  // registers. No `empty` barrier is needed: a stage is refilled only after the group's first barrier of the tile
  // that used it (program order for a one-warp group), and every thread reads its part of a tile before it. If the
  // group's tiles all fit in its ring they are loaded once and stay resident for the trajectory.
  const int ntiles = (ncols + TC - 1) / TC;
  const int tpg = ntiles > grp ? (ntiles - grp + NG - 1) / NG : 0;     // tiles of this group per step
  const bool resident = tpg <= SG;
  const int64_t loads_total = resident ? tpg : (int64_t)tpg * (L + 1);
  double* gring = ring + (size_t)grp * SG * TC * n_e;
  uint64_t* gfull = full + grp * SG;
  int64_t ld_left = loads_total;   // tiles of the stream still to request
  int ld_tile = 0;                 // next one: position in the group's tile list (0 .. tpg-1)
  int ld_stage = 0;                //           ring stage
  int ld_warp = 0;                 //           issuing warp of the group (the duty rotates)
  auto issue_one = [&]() {
    if (wg == ld_warp) {
      const int e0 = (grp + NG * ld_tile) * TC;
      const int nc = min(TC, ncols - e0);
      if (lane == 0) mbar_expect_tx(&gfull[ld_stage], col_bytes * (uint32_t)nc);
      __syncwarp();
      // neighbouring columns of a dense X (all features in the restricted set, no padding between columns) are one
      // contiguous run in memory and in the stage: one copy instead of TC
      bool run = TC > 1 && g.ld == (int64_t)n_e;
#pragma unroll
      for (int c = 1; c < TC; ++c) run = run && (c >= nc || cidx[e0 + c] == cidx[e0] + c);
      if (run) {
        if (lane == 0)
          bulk_g2s(gring + (size_t)ld_stage * TC * n_e, b.X + (int64_t)cidx[e0] * g.ld, col_bytes * (uint32_t)nc,
                   &gfull[ld_stage]);
      } else if (lane < nc)
        bulk_g2s(gring + ((size_t)ld_stage * TC + lane) * n_e, b.X + (int64_t)cidx[e0 + lane] * g.ld, col_bytes,
                 &gfull[ld_stage]);
    }
    --ld_left;
    if (++ld_tile == tpg) ld_tile = 0;
    if (++ld_stage == SG) ld_stage = 0;
    if (++ld_warp == GW) ld_warp = 0;
  };
  {
    const int pre = (int)(loads_total < SG ? loads_total : SG);
    for (int j = 0; j < pre; ++j) issue_one();
  }

  double Rg[NRP][2][KT];
  // row phase: this thread's row of the block's row slice; its fixed part, label and one-hot row never
  // change during a trajectory, so they are read once (into shared memory: registers are for the sweep)
  const int row0 = cta * fg.rpc, nrows = max(0, min(n, row0 + fg.rpc) - row0);
  if (ct < nrows) {
    double* rc = rowc + (size_t)ct * (2 * KT + 2);
    rc[2 * KT] = (double)b.ybase[row0 + ct];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      rc[k] = b.lv[(int64_t)k * n_s + row0 + ct];   // running linear predictor of the row (updated every step)
      rc[KT + k] = b.ymat[(int64_t)k * n_s + row0 + ct];
    }
  }
  const bool timer = (fg.timing != 0 && cta == 0 && tid == 0);   // only under htlr_cuda_profile_enable
  long long tk0 = 0;   // phase cycle counters go straight to the control block (block 0, thread 0 only) as
                       // fire-and-forget reductions: a read-modify-write would stall the warp on an L2 round trip
  int tcount = 0;           // tiles this group has processed (parity selects the exchange buffers)
  int pw = 0;               // warp of the group that applies the update to the current tile (the duty rotates)
  int stage = 0;            // ring stage of the tile being consumed
  uint32_t fph = 0;         // parity to wait for on its `full` barrier
  // lane -> value index it holds after the transposing reduction, and whether it is the lane that stores it
  const int vmine = VP == 8 ? (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1))
                            : (((lane >> 4) & 1) * 2 + ((lane >> 3) & 1));
  const bool vwriter = (lane & (32 / VP - 1)) == 0;
  for (int s = 0; s <= L; ++s) {
    if (timer) tk0 = clk();
    if (CL && s >= 1) {
      // the residual of step s - 1's softmax has arrived from every row's owner; thread 0 arms the next phase (it cannot
      // complete before this block has sent the partials of step s, i.e. before all its threads are past this wait)
      // (the block barrier orders what the cluster barrier at the end of a step used to order inside the block: the
      // prior-gradient terms and the row state written during the row phase)
      mbar_wait(&xbar[1], (uint32_t)((s - 1) & 1));
      if (tid == 0) mbar_expect_tx(&xbar[1], xbytes1);
      named_bar(kBarAll, kFT);
    }
    // residual rows of this thread into registers (published by the previous softmax phase)
#pragma unroll
    for (int m = 0; m < NRP; ++m) {
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        double r0v = 0, r1v = 0;
        if (rval[m]) {
          if (CL) {
            const double2 rr = *reinterpret_cast<const double2*>(Rs + k * n_e + roff[m]);   // pad row holds 0
            r0v = rr.x;
            r1v = rr.y;
          } else {
            // 16-byte load; the pad row of an odd n holds 0 in R (never written)
            const double2 rr = __ldcg(reinterpret_cast<const double2*>(b.R + (int64_t)k * n_s + roff[m]));
            r0v = rr.x;
            r1v = rr.y;
          }
        }
        Rg[m][0][k] = r0v;
        Rg[m][1][k] = r1v;
      }
    }
    double la[NRP][2][KT];
#pragma unroll
    for (int m = 0; m < NRP; ++m)
#pragma unroll
      for (int k = 0; k < KT; ++k) { la[m][0][k] = 0; la[m][1][k] = 0; }

    if (resident) stage = 0;
    for (int ti = 0; ti < tpg; ++ti, ++tcount) {
      const int e0 = (grp + NG * ti) * TC;          // first block-local column of the tile
      const int nc = min(TC, ncols - e0);           // columns in it (the block's last tile may be short)
      const int buf = tcount & 1;
      if (!resident || s == 0) mbar_wait(&gfull[stage], fph);
      const double* tile = gring + (size_t)stage * TC * n_e;
      // ---- the thread's rows of every column of the tile: the only shared-memory read of X. Unconditional:
      // the ring was zero-filled, so a stage only ever holds X values or zeros; a row pair past the end reads
      // pair 0 and meets a zero residual, a column past the block's last one meets a zero coefficient below.
      double2 x[TC][NRP];
#pragma unroll
      for (int c = 0; c < TC; ++c)
#pragma unroll
        for (int m = 0; m < NRP; ++m) x[c][m] = *reinterpret_cast<const double2*>(tile + c * n_e + roff[m]);
      // ---- gradient partials of the tile (two independent chains per value)
      double vals[VP];
#pragma unroll
      for (int c = 0; c < TC; ++c)
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          double a0 = 0, a1 = 0;
#pragma unroll
          for (int m = 0; m < NRP; ++m) {
            a0 = fma(x[c][m].x, Rg[m][0][k], a0);
            a1 = fma(x[c][m].y, Rg[m][1][k], a1);
          }
          vals[c * KT + k] = a0 + a1;
        }
#pragma unroll
      for (int v = V; v < VP; ++v) vals[v] = 0;
      // ---- transposing butterfly: halve the value set at each of the first log2(VP) levels, then plain xor
      // levels; lane l ends up with the warp total of value vmine(l). Fixed order: deterministic.
      {
        int half = VP / 2;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          if (half >= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < VP / 2; ++i) {
              if (i < half) {
                const double send = upper ? vals[i] : vals[i + half];
                const double keep = upper ? vals[i + half] : vals[i];
                vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            half >>= 1;
          } else {
            vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], off);
          }
        }
      }
      double* gp = gpart + ((size_t)buf * kCW + grp * GW) * VP;
      if (vwriter) gp[wg * VP + vmine] = vals[0];
      if (GW > 1) named_bar(kBarGroup + grp, GT); else __syncwarp();
      // every thread of the group has its part of the tile in registers: the stage is free, request the tile SG
      // ahead into it (SG tiles of the stream are always in flight or landed)
      if (ld_left > 0) issue_one();
      // ---- one warp of the group: DNlogpost, MoveMomt of step s, UpdateMomtAndDeltas of step s+1;
      // lane <-> (column of the tile, class)
      double* dnb = dnsm + ((size_t)buf * NG + grp) * VP;
      if (wg == pw) {
        if (lane < V) {
          const int c = lane / KT, k = lane - c * KT;
          double dnew = 0.0, dinc = 0.0;
          if (c < nc) {
            double gsum = gp[lane];
#pragma unroll
            for (int ww = 1; ww < GW; ++ww) gsum += gp[ww * VP + lane];
            double* st = cst + (size_t)(e0 + c) * CS;
            const double dk = st[k], stp = st[3 * KT + 1], hs = stp / 2;
            double mm = st[KT + k];
            const double post = gsum + st[2 * KT + k];
            const double kick = __dmul_rn(hs, post);
            if (s >= 1) mm = __dsub_rn(mm, kick);
            dnew = dk;
            if (s < L) {
              mm = __dsub_rn(mm, kick);
              dnew = __dadd_rn(dk, __dmul_rn(stp, mm));
              dinc = __dsub_rn(dnew, dk);
            }
            st[k] = dnew;
            st[KT + k] = mm;
          }
          dnb[lane] = dinc;
        }
      }
      if (GW > 1) named_bar(kBarGroup + grp, GT); else __syncwarp();
      // ---- the tile's contribution to the CHANGE of the linear predictor, from the registers: the predictor is
      // carried from step to step (lv += X[:, iup] (delta_new - delta_old)), so no detached fixed part is needed
      if (s < L) {
#pragma unroll
        for (int c = 0; c < TC; ++c) {
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const double dnk = dnb[c * KT + k];   // coefficient increment; 0 for a column past the block's last one
#pragma unroll
            for (int m = 0; m < NRP; ++m) {
              la[m][0][k] = fma(x[c][m].x, dnk, la[m][0][k]);
              la[m][1][k] = fma(x[c][m].y, dnk, la[m][1][k]);
            }
          }
        }
      }
      if (++stage == SG) { stage = 0; fph ^= 1; }
      if (++pw == GW) pw = 0;
    }
    if (timer) { const long long t = clk(); atomicAdd(&ctl->ph_cycles[0], (unsigned long long)(t - tk0)); tk0 = t; }
    if (s == L) break;

    // Cross-group reduction of the lv partial. With 4 or 16 groups (n <= 512) every group leaves its registers in a slot
    // of its own and group 0 adds the slots in group order after ONE barrier (config A 156 -> 149, config B full
    // 551 -> 535 us per iteration). With 2 groups (n = 513..1024, config C) the one-slot hand-over below is kept: it
    // costs two more barriers per step, but the instantiation with the slot writes ahead of the barrier measured
    // 4 % slower on config C (the hot loop's schedule changes), same box, same run.
    if constexpr (NG > 2) {
      // ---- groups 1.. (those that had a tile) leave their lv registers in their own slot; after ONE barrier of the
      // block group 0 adds the slots in group order (every group maps threads to rows the same way)
      if (NG > 1 && grp >= 1 && grp < ntiles) {
        double* slot = lasm + (size_t)(grp - 1) * KT * n_e;
#pragma unroll
        for (int m = 0; m < NRP; ++m) {
          if (rval[m]) {
#pragma unroll
            for (int k = 0; k < KT; ++k)
              *reinterpret_cast<double2*>(slot + (size_t)k * n_e + roff[m]) = make_double2(la[m][0][k], la[m][1][k]);
          }
        }
      }
      named_bar(kBarCompute, kCT);   // every group is done with the sweep
      if (NG > 1 && grp == 0) {
#pragma unroll 1
        for (int gg = 1; gg < NG; ++gg) {
          if (ntiles <= gg) break;   // that group (and every later one) had no tile: nothing to add
          const double* slot = lasm + (size_t)(gg - 1) * KT * n_e;
#pragma unroll
          for (int m = 0; m < NRP; ++m) {
            if (rval[m]) {
#pragma unroll
              for (int k = 0; k < KT; ++k) {
                const double2 t = *reinterpret_cast<const double2*>(slot + (size_t)k * n_e + roff[m]);
                la[m][0][k] += t.x;
                la[m][1][k] += t.y;
              }
            }
          }
        }
      }
    } else {
      // ---- every group is done with the sweep: next step's prior-gradient terms, then this block's lv partial
      named_bar(kBarCompute, kCT);
      // groups 1.. hand their lv registers to group 0 one after the other through a one-group buffer (every group
      // maps threads to rows the same way), group 0 adds them in group order
      if (NG > 1) {
#pragma unroll 1
        for (int gg = 1; gg < NG; ++gg) {
          if (ntiles <= gg) break;   // that group (and every later one) had no tile: nothing to add
          if (grp == gg) {
#pragma unroll
            for (int m = 0; m < NRP; ++m) {
              if (rval[m]) {
#pragma unroll
                for (int k = 0; k < KT; ++k)
                  *reinterpret_cast<double2*>(lasm + (size_t)k * n_e + roff[m]) = make_double2(la[m][0][k], la[m][1][k]);
              }
            }
          }
          named_bar(kBarCompute, kCT);
          if (grp == 0) {
#pragma unroll
            for (int m = 0; m < NRP; ++m) {
              if (rval[m]) {
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                  const double2 t = *reinterpret_cast<const double2*>(lasm + (size_t)k * n_e + roff[m]);
                  la[m][0][k] += t.x;
                  la[m][1][k] += t.y;
                }
              }
            }
          }
          named_bar(kBarCompute, kCT);
        }
      }
    }
    if (cta < gact && grp == 0) {
#pragma unroll
      for (int m = 0; m < NRP; ++m) {
        if (rval[m]) {
          // cluster mode: the pair goes to the block that owns its rows. rpc is even there (fused_plan), so a pair has
          // one owner and a 16-byte aligned slot in [source block][class][row of the owner's slice]
          const int owner = CL ? roff[m] / fg.rpc : 0;
          const uint32_t pdst = CL ? map_to_cta(smem_u32(prcv) + (uint32_t)((cta * KT) * fg.rpc + roff[m] - owner * fg.rpc) * 8u,
                                                (uint32_t)owner) : 0u;
          const uint32_t pbar = CL ? map_to_cta(smem_u32(&xbar[0]), (uint32_t)owner) : 0u;
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            if (CL) {
              st_async_2f64(pdst + (uint32_t)(k * fg.rpc) * 8u, la[m][0][k], la[m][1][k], pbar);
            } else {
              double* o = b.lvpart + ((int64_t)cta * KT + k) * n_s + roff[m];
              *reinterpret_cast<double2*>(o) = make_double2(la[m][0][k], la[m][1][k]);
            }
          }
        }
      }
    }
    if (!CL) sync_all();
    if (timer) { const long long t = clk(); atomicAdd(&ctl->ph_cycles[1], (unsigned long long)(t - tk0)); tk0 = t; }

    // ---- sum the partials in block order, softmax, residual, log-likelihood (rows split over all blocks)
    {
      const bool last_step = (s == L - 1);
      // Row softmax with one LANE per (row, class), class 0 (lv = 0) included: the C exponentials of a row, and
      // then its K residual exponentials, run side by side instead of one after the other in one thread — the chain
      // per row is exp, log, exp whatever K is (a thread per row needs 2K+1 exps and a log in sequence, ~160 cycles
      // each). Rows of a warp: RPW = 32 / C; used when the block's row slice fits 16 warps that way.
      constexpr int LPR = KT + 1;
      constexpr int RPW = 32 / LPR;
      // (only a cluster's row slice with K >= 4 may not fit - n > 1536 at K = 4, n > 768 at K = 8: the thread-per-row
      // form below is compiled for those instantiations alone)
      constexpr bool kMayNotFit = CL && KT >= 4;
      // Measured on one box (tools/timeline.py, us per iteration): K = 3 (config B) 235 -> 223 default cut, 569 -> 555
      // full; K = 1 and K = 2 are slower this way (config A 163.6 -> 167.5, config C default 408 -> 419), so they
      // keep the thread-per-row form.
      constexpr int kByClassMinK = 3;
      const bool by_class = KT >= kByClassMinK && (!kMayNotFit || fg.rpc <= kCW * RPW);
      const int cr = lane / LPR, cc = lane - cr * LPR;       // this lane's row within the warp, and class
      const int crow = w * RPW + cr;                          // row of the block's slice
      const bool chave = by_class && lane < RPW * LPR && crow < nrows;
      double l[KT];
      double lc = 0.0;   // by_class: this lane's linear predictor (class 0: 0)
      // the partials of one (row, class) item of the cluster's blocks, added in block order
      auto gather = [&](int ii, int k) {
        const double* src = prcv + k * fg.rpc + ii;   // [source block][class][row]: received, local
        double pv[kClMax];
#pragma unroll
        for (int qq = 0; qq < kClMax; ++qq) pv[qq] = (qq < gact) ? src[qq * KT * fg.rpc] : 0.0;
        double a = pv[0];
#pragma unroll
        for (int qq = 1; qq < kClMax; ++qq) a += pv[qq];
        return rowc[(size_t)ii * (2 * KT + 2) + k] + a;
      };
      if (CL) {
        // the partials of this block's rows have arrived from every block that owns columns; thread 0 arms the next
        // phase (it cannot complete before this block's residuals of this step are out, i.e. after every wait here)
        if (xrows > 0) {
          mbar_wait(&xbar[0], (uint32_t)(s & 1));
          if (tid == 0) mbar_expect_tx(&xbar[0], xbytes0);
        }
        if (by_class) {
          if (chave && cc >= 1) lc = gather(crow, cc - 1);
        } else if (KT == 1) {
          if (ct < nrows) l[0] = gather(ct, 0);
        } else {
          // thread <-> (row, class) item first (one round of remote loads for all classes), then thread <-> row
          for (int item = ct; item < nrows * KT; item += kCT) {
            const int k = item / nrows, ii = item - k * nrows;
            dl[ii * KT + k] = gather(ii, k);
          }
          named_bar(kBarCompute, kCT);
          if (ct < nrows) {
#pragma unroll
            for (int k = 0; k < KT; ++k) l[k] = dl[ct * KT + k];
          }
        }
      } else {
        // lane <-> (row, class) item of this block's row slice, warp <-> every 16th publishing block: one warp
        // instruction reads the block's whole slice of one partial (contiguous rows: full sectors, each fetched
        // once), a lane has all its loads in flight before the first add; per-warp sums go through shared memory
        // and are added in warp order. Fixed order throughout: deterministic.
        {
          const int nitems = nrows * KT;
          double a = 0;
          if (lane < nitems) {
            const int k = lane / nrows, ii = lane - k * nrows;
            const double* src = b.lvpart + (int64_t)k * n_s + row0 + ii;
            double pv[kWriterSlots];
#pragma unroll
            for (int qq = 0; qq < kWriterSlots; ++qq) {
              const int sl = w + kCW * qq;
              pv[qq] = (sl < gact) ? __ldcg(src + (int64_t)sl * KT * n_s) : 0.0;
            }
            a = pv[0];
#pragma unroll
            for (int qq = 1; qq < kWriterSlots; ++qq) a += pv[qq];
          }
          // items beyond one warp's lanes (more than 32 (row, class) pairs per block: small grids only)
          for (int item = lane + 32; item < nitems; item += 32) {
            const int k = item / nrows, ii = item - k * nrows;
            const double* src = b.lvpart + (int64_t)k * n_s + row0 + ii;
            double t = 0;
            for (int sl = w; sl < gact; sl += kCW) t += __ldcg(src + (int64_t)sl * KT * n_s);
            wsum[(size_t)w * fg.rpc * KT + item] = t;
          }
          if (lane < nitems) wsum[(size_t)w * fg.rpc * KT + lane] = a;
        }
        named_bar(kBarCompute, kCT);
        if (ct < nrows * KT) {
          double t = wsum[ct];
#pragma unroll
          for (int ww = 1; ww < kCW; ++ww) t += wsum[(size_t)ww * fg.rpc * KT + ct];
          const int k = ct / nrows, ii = ct - k * nrows;
          dl[ii * KT + k] = t;
        }
        named_bar(kBarCompute, kCT);
        if (by_class) {
          if (chave && cc >= 1) lc = rowc[(size_t)crow * (2 * KT + 2) + cc - 1] + dl[crow * KT + cc - 1];
        } else if (ct < nrows) {
#pragma unroll
          for (int k = 0; k < KT; ++k) l[k] = rowc[(size_t)ct * (2 * KT + 2) + k] + dl[ct * KT + k];
        }
      }
      // next step's prior-gradient terms, by the warps that hold no row of the softmax below (needed only
      // after the second barrier)
      {
        int base = by_class ? 32 * ((nrows + RPW - 1) / RPW) : ((nrows + 31) & ~31);
        if (base >= kCT) base = 0;
        if (ct >= base)
          for (int e = ct - base; e < ncols; e += kCT - base) prior_gradient(e);
      }
      double ll = 0;
      // cluster mode: the new residual of row i, class k into every block's copy, counted in on the receivers' barriers
      // (one 8-byte store per value and block: pairing two rows into a 16-byte store measured slower, and so did a
      // staging area with one bulk copy per class and block)
      auto push_residual = [&](int k, int i, double res) {
        const uint32_t dst = smem_u32(Rs) + (uint32_t)(k * n_e + i) * 8u, bar1 = smem_u32(&xbar[1]);
#pragma unroll
        for (int qq = 0; qq < kClMax; ++qq)
          if (qq < (int)G) st_async_f64(map_to_cta(dst, (uint32_t)qq), res, map_to_cta(bar1, (uint32_t)qq));
      };
      if (by_class) {
        if (w * RPW < nrows) {   // warp-uniform: all 32 lanes take part in the shuffles
          // find_normlv (utils.cpp:10-26) with the same operations in the same order as the thread-per-row form:
          // m = max(0, l_1..l_K); se = exp(0 - m) + exp(l_1 - m) + ... (left to right); lse = log(se) + m
          const int rb = min(cr, RPW - 1) * LPR;   // first lane of this lane's row
          double m = 0.0;
#pragma unroll
          for (int q = 0; q < LPR; ++q) m = fmax(m, __shfl_sync(0xffffffffu, lc, rb + q));
          const double e = exp(lc - m);
          double se = __shfl_sync(0xffffffffu, e, rb);
#pragma unroll
          for (int q = 1; q < LPR; ++q) se += __shfl_sync(0xffffffffu, e, rb + q);
          const double lse = log(se) + m;
          const double nl = lc - lse;
          if (chave) {
            const int i = row0 + crow;
            double* rc = rowc + (size_t)crow * (2 * KT + 2);
            if ((int)rc[2 * KT] == cc) ll = nl;
            if (cc >= 1) {
              const int k = cc - 1;
              const double res = exp(nl) - rc[KT + k];
              rc[k] = lc;                                       // the row's running linear predictor
              if (last_step) {
                b.lv[(int64_t)k * n_s + i] = lc;                // the proposal's lv and residual: only the final ones are state
                if (CL) b.R[(int64_t)k * n_s + i] = res;
              }
              if (CL) push_residual(k, i, res);
              else __stcg(&b.R[(int64_t)k * n_s + i], res);
            }
          }
        }
      } else if (ct < nrows) {
        const int i = row0 + ct;
        double* rc = rowc + (size_t)ct * (2 * KT + 2);
        const int row_yb = (int)rc[2 * KT];
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k) m = fmax(m, l[k]);
        double se = exp(0.0 - m);
#pragma unroll
        for (int k = 0; k < KT; ++k) se += exp(l[k] - m);
        const double lse = log(se) + m;
        if (row_yb == 0) ll = 0.0 - lse;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const double nl = l[k] - lse;
          if (row_yb == k + 1) ll = nl;
          const double res = exp(nl) - rc[KT + k];
          rc[k] = l[k];                                       // the row's running linear predictor
          if (last_step) {
            b.lv[(int64_t)k * n_s + i] = l[k];                // the proposal's lv and residual: only the final ones are state
            if (CL) b.R[(int64_t)k * n_s + i] = res;
          }
          if (CL) {
            push_residual(k, i, res);
          } else {
            __stcg(&b.R[(int64_t)k * n_s + i], res);
          }
        }
      }
      // UpdateLogLike (gibbs.cpp:254-261) is only read after the last step of the trajectory (CompNegEnergy,
      // gibbs.cpp:418-419): the block reduction runs once per launch, not once per step
      if (last_step) {
        ll = warp_sum(ll);
        if (lane == 0) redsm[w] = ll;
        named_bar(kBarCompute, kCT);
        if (ct == 0) {
          double t = 0;
          for (int i = 0; i < kCW; ++i) t += redsm[i];
          b.red[cta].c = t;
        }
      }
    }
    if (timer) { const long long t = clk(); atomicAdd(&ctl->ph_cycles[2], (unsigned long long)(t - tk0)); tk0 = t; }
    if (!CL) sync_all();
    if (timer) { const long long t = clk(); atomicAdd(&ctl->ph_cycles[3], (unsigned long long)(t - tk0)); }
  }
  // cluster mode: the one barrier after the trajectory - the log-likelihood partials below are in global memory, and
  // no block may leave while another can still store into its shared memory
  if (CL) cluster_barrier();
  // final coefficients and momenta back to the compact global arrays (every group has finished its tiles)
  named_bar(kBarCompute, kCT);
  for (int e = ct; e < ncols; e += kCT) {
    const int r = c0 + e;
    const double* st = cst + (size_t)e * CS;
#pragma unroll
    for (int k = 0; k < KT; ++k) { b.dc[(int64_t)r * KT + k] = st[k]; b.momt[(int64_t)r * KT + k] = st[KT + k]; }
  }

  // every block has passed the last grid-wide barrier, so all log-likelihood partials are visible
  if (cta == 0 && tid == 0) {
    double A = 0;
    for (unsigned c = 0; c < G; ++c) A += __ldcg(&b.red[c].c);
    ctl->loglike_new = A;
    if (CL) ctl->traj_done = 1;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
namespace {

template <int KT, int NRP, int GW, bool CL>
cudaError_t launch_one(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st) {
  static PerDeviceOnce configured;
  auto kern = k_traj<KT, NRP, GW, CL>;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fg.smem_bytes_max);
    if (e != cudaSuccess) return e;
    if (CL && fg.cl > 8) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
    }
    configured.mark();
  }
  Geom gg = g;
  Bufs bb = b;
  FusedGeom ff = fg;
  int LL = L;
  void* args[] = {&gg, &bb, &ff, &LL};
  if (!CL) return cudaLaunchCooperativeKernel((void*)kern, dim3(fg.grid), dim3(kFT), args, fg.smem_bytes, st);
  // one cluster: its CTAs are co-scheduled by construction, no cooperative launch needed
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(fg.grid);
  cfg.blockDim = dim3(kFT);
  cfg.dynamicSmemBytes = fg.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = fg.cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelExC(&cfg, (void*)kern, args);
}

template <int KT, int NRP, bool CL>
cudaError_t launch_gw(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st) {
  switch (fg.gw) {
    case 1: return launch_one<KT, NRP, 1, CL>(g, b, fg, L, st);
    case 4: return launch_one<KT, NRP, 4, CL>(g, b, fg, L, st);
    case 8: return launch_one<KT, NRP, 8, CL>(g, b, fg, L, st);
    case 16: return launch_one<KT, NRP, 16, CL>(g, b, fg, L, st);
  }
  return cudaErrorInvalidValue;
}

template <int KT, bool CL>
cudaError_t launch_k(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st) {
  if (fg.nrp == 1) return launch_gw<KT, 1, CL>(g, b, fg, L, st);
  if constexpr (KT <= 4) {   // two row pairs per thread only fit the register file up to K = 4
    if (fg.nrp == 2) return launch_gw<KT, 2, CL>(g, b, fg, L, st);
  }
  return cudaErrorInvalidValue;
}

template <bool CL>
cudaError_t launch_mode(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st) {
  switch (g.K) {
    case 1: return launch_k<1, CL>(g, b, fg, L, st);
    case 2: return launch_k<2, CL>(g, b, fg, L, st);
    case 3: return launch_k<3, CL>(g, b, fg, L, st);
    case 4: return launch_k<4, CL>(g, b, fg, L, st);
    case 5: return launch_k<5, CL>(g, b, fg, L, st);
    case 6: return launch_k<6, CL>(g, b, fg, L, st);
    case 7: return launch_k<7, CL>(g, b, fg, L, st);
    case 8: return launch_k<8, CL>(g, b, fg, L, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

// Picks the group width and ring depth for (n, K); returns false when the fused path does not apply (a column
// must fit in shared memory with room for a pipeline, K <= 8). cl = 0: grid mode (one block per SM); cl > 0:
// one cluster of cl blocks, which keeps per-column state for at most cl * ncmax restricted columns.
bool fused_plan(const Geom& g, int smem_optin_bytes, int cl, FusedGeom* out) {
  const int n = g.n, K = g.K;
  if (K < 1 || K > 8) return false;
  const int npairs = (n + 1) / 2, n_e = 2 * npairs;
  if (npairs > kCT * kNlpMax) return false;
  const int G = cl > 0 ? cl : g.sm_count;
  if (cl > kClMax) return false;
  int rpc = (n + G - 1) / G;
  if (cl > 0) rpc = (rpc + 1) & ~1;   // cluster mode: a row pair has one owner (16-byte pushes of the lv partials)
  if (rpc > kCT) return false;
  if (G > 32 * kPartSlots) return false;   // row phase: one lane sums at most kPartSlots block partials
  // smallest group whose threads keep their row pairs (x, residual, lv partial) in registers
  // row pairs per thread: x (TC x NRP double2), residual and lv partial (4 NRP K doubles) stay in registers
  const int nrp_max = K <= 4 ? 2 : 1;
  int gw = 0;
  const char* force = std::getenv("HTLR_FUSED_GW");   // experiments only
  for (int cand : {1, 4, 8, 16}) {
    if (force && std::atoi(force) != cand) continue;
    if ((npairs + 32 * cand - 1) / (32 * cand) <= nrp_max) { gw = cand; break; }
  }
  if (!gw) return false;
  const int nrp = (npairs + 32 * gw - 1) / (32 * gw);
  const int TC = (K == 1 && nrp == 1 && gw >= 4) ? 8 : (K <= 2 ? 4 : (K <= 4 ? 2 : 1));   // TileCols<K, NRP, GW>
  const int ng = kCW / gw;
  const size_t col_bytes = (size_t)npairs * 16;
  int ncmax;
  if (cl > 0) ncmax = (int)(24 * 1024 / ((3 * K + 2) * sizeof(double) + sizeof(int))) & ~7;   // columns per block the
                                                                       // cluster keeps state for (24 KB)
  else ncmax = (g.nvar + G - 1) / G + 1;   // worst case: every feature in the restricted set
  const size_t other = (size_t)(2 * kCW * 8 + 2 * kCW * 8 + kCW + ((rpc * K + 1) & ~1) + rpc * (2 * K + 2) + 2 +
                                (cl > 0 ? 0 : kCW * rpc * K)) * sizeof(double) +
                       (size_t)(ng - 1) * K * n_e * sizeof(double) +
                       (size_t)ncmax * ((3 * K + 2) * sizeof(double) + sizeof(int)) +
                       (cl > 0 ? (size_t)(cl * K * rpc + K * n_e) * sizeof(double) : 0) + 2 * kNsMax * sizeof(uint64_t) + 256;
  const size_t tile_bytes = (size_t)TC * col_bytes;
  if ((size_t)smem_optin_bytes < other + 1024 + 2 * ng * tile_bytes) return false;
  const size_t budget = (size_t)smem_optin_bytes - other - 1024;
  const int SG = (int)std::min<size_t>(kNsMax / ng, budget / tile_bytes / ng);   // ring stages (tiles) per group
  if (SG < 2) return false;
  const int NS = SG;
  out->T = TC; out->S = n_e; out->NS = NS; out->nrp = nrp; out->gw = gw; out->nlp = (npairs + kCT - 1) / kCT;
  out->rpc = rpc;
  out->grid = G;
  out->cl = cl;
  out->ncmax = ncmax;
  out->smem_bytes = (size_t)NS * ng * tile_bytes + other;
  out->smem_bytes_max = (size_t)smem_optin_bytes;
  return true;
}

// Largest cluster (16, else 8) that can be resident with this shared-memory footprint, or 0.
int fused_cluster_size(const Geom& g, int smem_optin_bytes) {
  int top = kClMax, bottom = 8;
  if (const char* force = std::getenv("HTLR_CLUSTER_SIZE")) top = bottom = std::atoi(force);   // experiments only
  for (int cl = top; cl >= bottom; cl >>= 1) {
    FusedGeom fg{};
    if (!fused_plan(g, smem_optin_bytes, cl, &fg)) continue;
    int nclusters = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl);
    cfg.blockDim = dim3(kFT);
    cfg.dynamicSmemBytes = fg.smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    // every instantiation has the same launch bounds and dynamic shared memory, so one probe is enough
    auto kern = k_traj<1, 1, 1, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_bytes) != cudaSuccess) continue;
    if (cl > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) continue;
    if (cudaOccupancyMaxActiveClusters(&nclusters, (void*)kern, &cfg) == cudaSuccess && nclusters >= 1) return cl;
    (void)cudaGetLastError();
  }
  (void)cudaGetLastError();
  return 0;
}

cudaError_t launch_traj_fused(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st) {
  return fg.cl > 0 ? launch_mode<true>(g, b, fg, L, st) : launch_mode<false>(g, b, fg, L, st);
}

}  // namespace htlr
