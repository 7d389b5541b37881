// Row-partitioned cluster kernel for restricted sets with FEW columns (u K of the order of 100 or less): the regime of
// config B with the product defaults (n = 500, u ~ 11) and of strongly sparse fits on larger n, where k_traj's cluster
// mode still exchanges n K partial predictors and n K residuals per block and leapfrog step.
//
// Here a block of the cluster owns a slab of ROWS of every restricted column (X[rows, iup] resident in its shared memory
// for the whole trajectory), so the linear predictor, the softmax and the residual of its rows never leave the block.
// What is exchanged per leapfrog step is the gradient: every block forms the partial sums X[rows, j]' r[rows, :] of
// all u columns over its rows and sends the u K numbers to every block (counted asynchronous stores: st.async +
// mbarrier complete_tx on the receiver, no fence, no cluster barrier); every block then adds the partials in block order
// and applies the leapfrog update to its own copy of the coefficients — all copies stay bit-identical because all blocks
// add the same numbers in the same order. One exchange of 8 u K bytes per block pair and step instead of two of 8 n K / 16.
//
// Same arithmetic as k_traj and k_traj_small (Fit::Traject, gibbs.cpp:390-419): all L+1 gradient evaluations in one
// launch, the predictor carried from step to step. Launched ahead of the other trajectory kernels; it decides on the
// device whether the restricted set is small enough (then sets Ctl::traj_done) or leaves the proposal to them.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "launch.h"

namespace htlr {

namespace {

constexpr int kRT = 512;      // threads per block
constexpr int kRCI = 4;       // columns a warp reduces at a time in the gradient phase
constexpr int kRowsClMax = 16;

__device__ __forceinline__ uint32_t r_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t r_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void r_st_async(uint32_t addr, double v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];"
               ::"r"(addr), "d"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void r_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(r_u32(bar)), "r"(count));
}
__device__ __forceinline__ void r_mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(r_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void r_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(r_u32(bar)), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void r_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace

// smem_doubles: dynamic shared memory available per block, in doubles; uk_max: largest u * K the kernel takes
template <int KT>
__global__ void __launch_bounds__(kRT, 1) k_traj_rows(Geom g, Bufs b, int L, int smem_doubles, int uk_max) {
  Ctl* ctl = b.ctl;
  if (ctl->traj_done) return;
  const int u = ctl->u;
  const int n = g.n, n_s = g.n_s, n_e = (n + 1) & ~1;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int NW = kRT / 32;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int UK = u * KT;
  // rows of this block: an even number per block, so slabs start on 16-byte boundaries and a row pair has one owner
  const int rpc = (((n + G - 1) / G) + 1) & ~1;
  const int row0 = cta * rpc;
  const int nr = max(0, min(n, row0 + rpc) - row0);        // rows of the slab
  const int nre = max(0, min(n_e, row0 + rpc) - row0);     // ... with the pad row of an odd n (X and the residual hold 0 there)
  const int gsrc = (n + rpc - 1) / rpc;                    // blocks that own rows (the others only keep the barriers company)
  const int npl = nre >> 1;                                // row pairs of the slab
  // lv phase: CH threads share a row when the slab has fewer rows than the block has threads
  const int n32 = (rpc + 31) & ~31;
  const int CH = n32 <= kRT / 2 ? kRT / n32 : 1;
  if (UK > uk_max || rpc > kRT) return;
  // the gradient exchange has two hops: item i = (column, class) is summed by block i % gsrc (its owner), which then
  // sends the sum to every block - 2 u K stores sent and received per block and step instead of 16 u K (a counted store
  // costs the receiver ~4.5 cycles: all-to-all was 1.3 us per step at u K = 36)
  const int IPO = (UK + gsrc - 1) / gsrc, IPOe = (IPO + 1) & ~1;      // items per owner
  const int mine = (cta < gsrc && UK > cta) ? (UK - 1 - cta) / gsrc + 1 : 0;   // items this block owns
  const size_t need = (size_t)u * rpc + (size_t)KT * rpc + (CH > 1 ? (size_t)KT * CH * n32 : 0) + (size_t)u * (4 * KT + 2) +
                      2 * (size_t)((UK + 1) & ~1) + 2 * (size_t)G * IPOe + 2 * NW + 10 + (size_t)(u + 1) / 2;
  if (need > (size_t)smem_doubles) return;       // does not fit: the other kernels take the proposal
  extern __shared__ __align__(128) unsigned char sm_raw[];
  const int UKe = (UK + 1) & ~1;
  double* Xs = reinterpret_cast<double*>(sm_raw);            // [u][rpc]   this block's rows of every restricted column
  double* Rs = Xs + (size_t)u * rpc;                          // [KT][rpc]  residual of the rows
  double* part = Rs + (size_t)KT * rpc;                       // [KT][CH][n32] (CH > 1)
  double* st = part + (CH > 1 ? (size_t)KT * CH * n32 : 0);   // [u][4*KT+2]: d, m, prior gradient, sig, step, increment of d
  double* recv = st + (size_t)u * (4 * KT + 2);                                    // [2][G][IPOe] partials of the items this block owns (by step parity)
  double* gfull = recv + 2 * (size_t)G * IPOe;                // [2][UKe] summed gradient, as sent by the owners (by step parity)
  double* red = gfull + 2 * (size_t)UKe;                      // [2*NW]
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 2 * NW);  // [0] TMA, [1], [2] partials arrived, [3], [4] sums arrived (by step parity)
  int* cidx = reinterpret_cast<int*>(bar + 6);                // [u]
  constexpr int CS = 4 * KT + 2;
  const double Cd = (double)ctl->C;
  const uint32_t abytes = (uint32_t)(gsrc * mine) * 8u, bbytes = (uint32_t)UK * 8u;

  if (tid == 0) {
    r_mbar_init(&bar[0], 1);
    r_mbar_init(&bar[1], 1);
    r_mbar_init(&bar[2], 1);
    r_mbar_init(&bar[3], 1);
    r_mbar_init(&bar[4], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (nr > 0) {   // the first two steps' partials and sums are expected from the start
      if (mine > 0) { r_mbar_expect(&bar[1], abytes); r_mbar_expect(&bar[2], abytes); }
      r_mbar_expect(&bar[3], bbytes);
      r_mbar_expect(&bar[4], bbytes);
    }
    if (cta == 0) ctl->cols_swept += (unsigned long long)u;   // X[:, iup] is read from memory once per trajectory here
  }
  for (int e = tid; e < u; e += kRT) cidx[e] = b.iup[e];
  __syncthreads();
  // X[rows, iup] -> shared memory, one bulk copy per column
  const uint32_t col_bytes = (uint32_t)nre * 8u;
  if (nre > 0) {
    if (tid == 0) r_mbar_expect(&bar[0], col_bytes * (uint32_t)u);
    for (int e = tid; e < u; e += kRT)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       r_u32(Xs + (size_t)e * rpc)),
                   "l"(b.X + (int64_t)cidx[e] * g.ld + row0), "r"(col_bytes), "r"(r_u32(&bar[0]))
                   : "memory");
  }
  // per-column state (every block keeps all of it) and the first prior-gradient term (gibbs.cpp:267-272, 281)
  for (int e = tid; e < u; e += kRT) {
    double* s = st + (size_t)e * CS;
#pragma unroll
    for (int k = 0; k < KT; ++k) { s[k] = b.dc[(int64_t)e * KT + k]; s[KT + k] = b.momt[(int64_t)e * KT + k]; }
    const double sig = b.sigc[e];
    s[3 * KT] = sig;
    s[3 * KT + 1] = b.stepc[e];
    double sd = s[0];
#pragma unroll
    for (int k = 1; k < KT; ++k) sd += s[k];
    const double sdc = sd / Cd;
#pragma unroll
    for (int k = 0; k < KT; ++k) s[2 * KT + k] = (s[k] - sdc) / sig;
  }
  for (int e = tid; e < KT * rpc; e += kRT) {
    const int k = e / rpc, i = e - k * rpc;
    Rs[e] = i < nr ? b.R[(int64_t)k * n_s + row0 + i] : 0.0;
  }
  // lv phase ownership: thread -> (row of the slab, column chunk); per-row constants in registers
  const int ch = CH > 1 ? tid / n32 : 0;
  const int rloc = CH > 1 ? tid - ch * n32 : tid;   // row of the slab
  const bool owner = ch == 0;                        // applies the softmax
  const bool have_row = rloc < nr && (CH == 1 || tid < CH * n32);
  double fixv[KT], yv[KT];                           // fixv: the row's running linear predictor
  int ybv = -1;
#pragma unroll
  for (int k = 0; k < KT; ++k) { fixv[k] = 0; yv[k] = 0; }
  if (owner && have_row) {
    ybv = b.ybase[row0 + rloc];
#pragma unroll
    for (int k = 0; k < KT; ++k) { fixv[k] = b.lv[(int64_t)k * n_s + row0 + rloc]; yv[k] = b.ymat[(int64_t)k * n_s + row0 + rloc]; }
  }
  const int cpc = (u + CH - 1) / CH;
  const int cl0 = min(u, ch * cpc), cl1 = min(u, cl0 + cpc);
  // Alternative mapping of the lv phase, used when the slab fits it: one LANE per (row, class), class 0 (lv = 0) included.
  // The row's C exponentials, and then its K residual exponentials, run side by side instead of one after the other in
  // one thread (2K+1 exps and a log in sequence): the chain per row is exp, log, exp whatever K is (like k_traj's row phase).
  constexpr int LPR = KT + 1;
  constexpr int RPW = 32 / LPR;
  const bool by_class = KT >= 2 && rpc <= NW * RPW;
  const int cr = lane / LPR, cc = lane - cr * LPR;       // this lane's row within the warp, and class
  const int crow = w * RPW + cr;                          // row of the slab
  const bool chave = by_class && lane < RPW * LPR && crow < nr;
  double cfix = 0.0, cy = 0.0;                            // by_class: running predictor and one-hot entry of (row, class)
  int cyb = -1;
  if (chave) {
    cyb = b.ybase[row0 + crow];
    if (cc >= 1) { cfix = b.lv[(int64_t)(cc - 1) * n_s + row0 + crow]; cy = b.ymat[(int64_t)(cc - 1) * n_s + row0 + crow]; }
  }
  if (nre > 0) r_mbar_wait(&bar[0], 0);
  __syncthreads();
  // every block's barriers exist (and are armed) before anybody stores into them
  r_cluster_sync();

  double ll_total = 0;
  for (int s = 0; s <= L; ++s) {
    const int par = s & 1;
    // ================= gradient partial over this block's rows: warp <-> kRCI columns at a time =========
    if (nr > 0) {
      for (int c0 = w; c0 < u; c0 += NW * kRCI) {
        double acc[kRCI][KT];
#pragma unroll
        for (int ci = 0; ci < kRCI; ++ci)
#pragma unroll
          for (int k = 0; k < KT; ++k) acc[ci][k] = 0;
        for (int m = lane; m < npl; m += 32) {
          double2 r[KT];
#pragma unroll
          for (int k = 0; k < KT; ++k) r[k] = *reinterpret_cast<const double2*>(Rs + k * rpc + 2 * m);   // pad row holds 0
#pragma unroll
          for (int ci = 0; ci < kRCI; ++ci) {
            if (c0 + NW * ci < u) {
              const double2 x = *reinterpret_cast<const double2*>(Xs + (size_t)(c0 + NW * ci) * rpc + 2 * m);
#pragma unroll
              for (int k = 0; k < KT; ++k) acc[ci][k] = fma(x.x, r[k].x, fma(x.y, r[k].y, acc[ci][k]));
            }
          }
        }
        // xor-shuffle trees of the columns the warp really has, interleaved (one fully unrolled variant per count, chosen
        // by a warp-uniform branch): an SM retires about one warp shuffle per cycle, and 16 warps reducing kRCI x K values
        // each whatever the number of columns were the longest phase of a leapfrog step with few columns
        auto tree = [&](auto nc_) {
          constexpr int nc = decltype(nc_)::value;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int ci = 0; ci < nc; ++ci)
#pragma unroll
              for (int k = 0; k < KT; ++k) acc[ci][k] += __shfl_xor_sync(0xffffffffu, acc[ci][k], o);
        };
        const int nci = (u - c0 + NW - 1) / NW;   // columns c0, c0 + NW, ... below u
        if (nci >= 4) tree(std::integral_constant<int, 4>{});
        else if (nci == 3) tree(std::integral_constant<int, 3>{});
        else if (nci == 2) tree(std::integral_constant<int, 2>{});
        else tree(std::integral_constant<int, 1>{});
        // ---- exchange, hop 1, straight from the registers (every lane holds the totals after the xor trees): lane
        // ci * KT + k sends this block's partial of item (column c0 + NW ci, class k) to the item's owner, block item % gsrc
        {
          double v = 0.0;
#pragma unroll
          for (int ci = 0; ci < kRCI; ++ci)
#pragma unroll
            for (int k = 0; k < KT; ++k)
              if (lane == ci * KT + k) v = acc[ci][k];
          const int ci = lane / KT;
          if (ci < kRCI && c0 + NW * ci < u) {
            const int item = (c0 + NW * ci) * KT + (lane - ci * KT);
            const int slot = item / gsrc, own = item - slot * gsrc;
            r_st_async(r_mapa(r_u32(recv + ((size_t)par * G + cta) * IPOe) + (uint32_t)slot * 8u, (uint32_t)own), v,
                       r_mapa(r_u32(&bar[1 + par]), (uint32_t)own));
          }
        }
      }
    }
    // ================= exchange, hop 2 ============
    if (nr > 0) {
      const uint32_t ph = (uint32_t)((s >> 1) & 1);
      // ---- hop 2: the owner adds the partials of its items in block order and sends the sums to every block. A barrier
      // is armed for step s + 2 by thread 0 right after its wait for step s: stores of step s + 2 can only come from blocks
      // that are past step s + 1, which needs this block's contribution to step s + 1, sent after every wait of step s.
      if (mine > 0) {
        r_mbar_wait(&bar[1 + par], ph);
        if (tid == 0 && s + 2 <= L) r_mbar_expect(&bar[1 + par], abytes);
        const uint32_t gb = r_u32(gfull + (size_t)par * UKe), xb = r_u32(&bar[3 + par]);
        for (int idx = tid; idx < mine * gsrc; idx += kRT) {
          const int slot = idx / gsrc, dest = idx - slot * gsrc;
          const double* rp = recv + (size_t)par * G * IPOe + slot;
          double pv[kRowsClMax];   // every load is issued before the first add
#pragma unroll
          for (int q = 0; q < kRowsClMax; ++q) pv[q] = q < gsrc ? rp[(size_t)q * IPOe] : 0.0;
          double gsum = pv[0];
#pragma unroll
          for (int q = 1; q < kRowsClMax; ++q) gsum += pv[q];
          r_st_async(r_mapa(gb + (uint32_t)(slot * gsrc + cta) * 8u, (uint32_t)dest), gsum, r_mapa(xb, (uint32_t)dest));
        }
      }
      r_mbar_wait(&bar[3 + par], ph);
      if (tid == 0 && s + 2 <= L) r_mbar_expect(&bar[3 + par], bbytes);
    }
    // ================= update: lane <-> (column, class), the same in every block (DNlogpost, MoveMomt,
    // UpdateMomtAndDeltas, next prior gradient). A column's K lanes are neighbours inside a warp; the sum of the new
    // coefficients over the classes goes through shuffles, in class order.
    {
      constexpr int CPW = 32 / KT;                          // columns per warp
      const int ucl = lane / KT, uk = lane - ucl * KT;
      for (int cb = w * CPW; cb < u; cb += NW * CPW) {      // warp-uniform
        const int c = cb + ucl;
        const bool act = lane < CPW * KT && c < u;
        double* sc = st + (size_t)(act ? c : 0) * CS;
        double dnk = 0.0, dold = 0.0, sig = 1.0;
        if (act) {
          const double stp = sc[3 * KT + 1], hs = stp / 2;
          sig = sc[3 * KT];
          const double post = gfull[(size_t)par * UKe + (size_t)c * KT + uk] + sc[2 * KT + uk];   // the same number in every block
          const double kick = __dmul_rn(hs, post);
          double mm = sc[KT + uk];
          if (s >= 1) mm = __dsub_rn(mm, kick);
          dold = sc[uk];
          dnk = dold;
          if (s < L) {
            mm = __dsub_rn(mm, kick);
            dnk = __dadd_rn(dold, __dmul_rn(stp, mm));
          }
          sc[KT + uk] = mm;
        }
        if (s < L) {
          const int base = ucl * KT;
          double sd = __shfl_sync(0xffffffffu, dnk, base);
#pragma unroll
          for (int q = 1; q < KT; ++q) sd += __shfl_sync(0xffffffffu, dnk, base + q);
          if (act) {
            const double sdc = sd / Cd;
            sc[3 * KT + 2 + uk] = __dsub_rn(dnk, dold);   // what the lv phase adds
            sc[uk] = dnk;
            sc[2 * KT + uk] = (dnk - sdc) / sig;
          }
        }
      }
    }
    __syncthreads();
    if (s == L) break;

    // ================= linear predictor, softmax, residual of this block's rows =================
    const bool last = (s == L - 1);
    double ll = 0;
    if (by_class) {
      if (w * RPW < nr) {   // warp-uniform: all 32 lanes take part in the shuffles
        // increment of the predictor of (row, class): four strided chains over the columns, added in a fixed order
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (chave && cc >= 1) {
          const double* xr = Xs + crow;
          const double* sc = st + 3 * KT + 2 + (cc - 1);
          int c = 0;
          for (; c + 3 < u; c += 4) {
            a0 = fma(xr[(size_t)c * rpc], sc[(size_t)c * CS], a0);
            a1 = fma(xr[(size_t)(c + 1) * rpc], sc[(size_t)(c + 1) * CS], a1);
            a2 = fma(xr[(size_t)(c + 2) * rpc], sc[(size_t)(c + 2) * CS], a2);
            a3 = fma(xr[(size_t)(c + 3) * rpc], sc[(size_t)(c + 3) * CS], a3);
          }
          for (; c < u; ++c) a0 = fma(xr[(size_t)c * rpc], sc[(size_t)c * CS], a0);
        }
        const double lc = cc >= 1 ? cfix + ((a0 + a1) + (a2 + a3)) : 0.0;
        cfix = lc;
        // find_normlv (utils.cpp:10-26): m = max(0, l_1..l_K); se = exp(0 - m) + exp(l_1 - m) + ... (left to right);
        // lse = log(se) + m - the reference's operations in the reference's order
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
          if (cyb == cc) ll = nl;
          if (cc >= 1) {
            const double res = exp(nl) - cy;
            Rs[(cc - 1) * rpc + crow] = res;
            if (last) {
              b.lv[(int64_t)(cc - 1) * n_s + row0 + crow] = lc;
              b.R[(int64_t)(cc - 1) * n_s + row0 + crow] = res;
            }
          }
        }
      }
    } else {
      double a[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) a[k] = 0;
      if (have_row) {
        for (int c = cl0; c < cl1; ++c) {
          const double x = Xs[(size_t)c * rpc + rloc];
          const double* sc = st + (size_t)c * CS + 3 * KT + 2;
#pragma unroll
          for (int k = 0; k < KT; ++k) a[k] = fma(x, sc[k], a[k]);
        }
      }
      if (CH > 1) {
        if (have_row) {
#pragma unroll
          for (int k = 0; k < KT; ++k) part[((size_t)k * CH + ch) * n32 + rloc] = a[k];
        }
        __syncthreads();
        if (have_row && owner) {
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            double t = part[(size_t)k * CH * n32 + rloc];
            for (int cc = 1; cc < CH; ++cc) t += part[((size_t)k * CH + cc) * n32 + rloc];
            a[k] = t;
          }
        }
      }
      if (have_row && owner) {
        const int i = row0 + rloc;
        double l[KT];
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k) { l[k] = fixv[k] + a[k]; fixv[k] = l[k]; m = fmax(m, l[k]); }
        // find_normlv (utils.cpp:10-26), the reference's operations in the reference's order (forming the probabilities
        // as e / se instead would save a log and an exp per step, 4-8 % of an iteration, and costs four orders of
        // magnitude of parity margin over a chain: measured, not adopted)
        double se = exp(0.0 - m);
#pragma unroll
        for (int k = 0; k < KT; ++k) se += exp(l[k] - m);
        const double lse = log(se) + m;
        if (ybv == 0) ll += 0.0 - lse;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const double nl = l[k] - lse;
          if (ybv == k + 1) ll += nl;
          const double res = exp(nl) - yv[k];
          Rs[k * rpc + rloc] = res;
          if (last) { b.lv[(int64_t)k * n_s + i] = l[k]; b.R[(int64_t)k * n_s + i] = res; }
        }
      }
    }
    if (last) ll_total = ll;
    __syncthreads();
  }

  // ---- log-likelihood of the proposal: block partials in global memory, added in block order by block 0 after the one
  // cluster barrier that also keeps every block alive until nobody can store into it any more
  {
    double v = warp_sum(ll_total);
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int i = 0; i < NW; ++i) t += red[i];
      b.red[cta].c = t;
    }
  }
  r_cluster_sync();
  if (cta == 0) {
    if (tid == 0) {
      double t = 0;
      for (int c = 0; c < gsrc; ++c) t += __ldcg(&b.red[c].c);
      ctl->loglike_new = t;
      ctl->traj_done = 1;
      ctl->small_runs += 1;
    }
    for (int e = tid; e < u; e += kRT) {
      const double* s = st + (size_t)e * CS;
#pragma unroll
      for (int k = 0; k < KT; ++k) { b.dc[(int64_t)e * KT + k] = s[k]; b.momt[(int64_t)e * KT + k] = s[KT + k]; }
    }
  }
}

namespace {

// Dynamic shared memory per block: the device maximum, like k_traj - a launch with another shared-memory carve-out
// between two k_traj launches costs ~10 us of reconfiguration per iteration even when the kernel returns at once.

template <int KT>
cudaError_t launch_rows_k(const Geom& g, const Bufs& b, int L, int cl, int uk_max, int kRowsSmem, cudaStream_t st) {
  static PerDeviceOnce configured;
  auto kern = k_traj_rows<KT>;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowsSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    configured.mark();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl);
  cfg.blockDim = dim3(kRT);
  cfg.dynamicSmemBytes = kRowsSmem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, g, b, L, kRowsSmem / (int)sizeof(double), uk_max);
}

}  // namespace

// Cluster size the row-partitioned kernel can run with on this device (16, else 8), or 0: K <= 4, a slab of at most
// kRT rows per block.
int rows_cluster_size(const Geom& g, int kRowsSmem) {
  if (g.K < 1 || g.K > 4) return 0;
  for (int cl = kRowsClMax; cl >= 8; cl >>= 1) {
    if ((g.n + cl - 1) / cl + 1 > kRT) continue;
    auto kern = k_traj_rows<1>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowsSmem) != cudaSuccess) continue;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) continue;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl);
    cfg.blockDim = dim3(kRT);
    cfg.dynamicSmemBytes = kRowsSmem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, (void*)kern, &cfg) == cudaSuccess && nclusters >= 1) return cl;
    (void)cudaGetLastError();
  }
  (void)cudaGetLastError();
  return 0;
}

cudaError_t launch_traj_rows(const Geom& g, const Bufs& b, int L, int cl, int uk_max, int smem_bytes, cudaStream_t st) {
  switch (g.K) {
    case 1: return launch_rows_k<1>(g, b, L, cl, uk_max, smem_bytes, st);
    case 2: return launch_rows_k<2>(g, b, L, cl, uk_max, smem_bytes, st);
    case 3: return launch_rows_k<3>(g, b, L, cl, uk_max, smem_bytes, st);
    case 4: return launch_rows_k<4>(g, b, L, cl, uk_max, smem_bytes, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace htlr
