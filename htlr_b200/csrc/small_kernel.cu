// Single-block trajectory kernel for SMALL restricted sets: when X[:, iup] (a few dozen columns) fits in one SM's
// shared memory, a leapfrog step needs no inter-block exchange at all — two __syncthreads instead of two grid or
// cluster barriers — and costs well under a microsecond. This is the regime of the reference's own example data
// (config A, colon: n = 62, u ~ 80) and of config B with the product defaults (n = 500, u ~ 11), where the CPU is
// fast and every microsecond of synchronisation shows.
//
// Same arithmetic as k_traj (traj_kernel.cu): Fit::Traject, gibbs.cpp:390-419, all L+1 gradient evaluations in one
// launch, X[:, iup] loaded once (TMA bulk copies) and resident for the whole trajectory.
//   gradient phase  warp <-> 4 columns at a time: lane-strided dot products with the residual (128-bit
//                   shared-memory loads), interleaved xor-shuffle trees;
//   update phase    thread <-> column: DNlogpost / MoveMomt / UpdateMomtAndDeltas and the next prior-gradient term;
//   lv phase        thread <-> (row, column chunk): lv += sum_j X[i, j] (delta_new - delta_old)[j, :] over the chunk's
//                   columns in ascending order, chunks summed in order (the predictor is carried from step to step,
//                   like k_traj: no detached fixed part), row softmax, residual; log-likelihood at the end.
// Launched ahead of the cluster-mode and grid-mode kernels; it decides on the device whether the restricted set
// fits (then sets Ctl::traj_done so the others return at once) or leaves the proposal to them.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "launch.h"

namespace htlr {

namespace {

constexpr int kST = 512;   // threads
constexpr int kCI = 4;     // columns a warp reduces at a time in the gradient phase

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

}  // namespace

// smem_doubles: dynamic shared memory available, in doubles
template <int KT>
__global__ void __launch_bounds__(kST, 1) k_traj_small(Geom g, Bufs b, int L, int smem_doubles, int u_max) {
  Ctl* ctl = b.ctl;
  timeline_mark(ctl, 4);
  if (ctl->traj_done) return;
  const int u = ctl->u;
  const int n = g.n, n_s = g.n_s, npairs = (n + 1) >> 1, n_e = 2 * npairs;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int NW = kST / 32;
  // column chunks of the lv phase: CH threads share a row when there are fewer rows than threads
  const int n32 = (n + 31) & ~31;
  const int CH = n32 <= kST / 2 ? kST / n32 : 1;
  const int RPT = (n + kST - 1) / kST;           // rows per thread when CH == 1
  constexpr int kRptMax = 2;
  if (RPT > kRptMax || u > u_max) return;
  // ---- shared-memory layout (doubles)
  const size_t need = (size_t)u * n_e + (size_t)KT * n_e + (CH > 1 ? (size_t)KT * CH * n32 : 0) +
                      (size_t)u * (5 * KT + 2) + (size_t)(u + 1) / 2 + 2 * NW + 8;
  if (need > (size_t)smem_doubles) return;       // does not fit: the cluster / grid kernels take the proposal
  extern __shared__ __align__(128) unsigned char sm_raw[];
  double* Xs = reinterpret_cast<double*>(sm_raw);            // [u][n_e]
  double* Rs = Xs + (size_t)u * n_e;                          // [KT][n_e]
  double* part = Rs + (size_t)KT * n_e;                       // [KT][CH][n32] (CH > 1)
  double* st = part + (CH > 1 ? (size_t)KT * CH * n32 : 0);   // [u][4*KT+2]: d, m, prior gradient, sig, step, increment of d
  double* gsm = st + (size_t)u * (4 * KT + 2);                // [u][KT] gradient of the step
  double* red = gsm + (size_t)u * KT;                         // [2*NW]
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 2 * NW);  // one mbarrier
  int* cidx = reinterpret_cast<int*>(bar + 1);                // [u]
  constexpr int CS = 4 * KT + 2;
  const double Cd = (double)ctl->C;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl->cols_swept += (unsigned long long)u;   // X[:, iup] is read from memory once per trajectory here
  }
  for (int e = tid; e < u; e += kST) cidx[e] = b.iup[e];
  __syncthreads();
  // X[:, iup] -> shared memory, one bulk copy per column
  const uint32_t col_bytes = (uint32_t)npairs << 4;
  if (tid == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(col_bytes * (uint32_t)u)
                 : "memory");
  for (int e = tid; e < u; e += kST)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     s_u32(Xs + (size_t)e * n_e)),
                 "l"(b.X + (int64_t)cidx[e] * g.ld), "r"(col_bytes), "r"(s_u32(bar))
                 : "memory");
  // per-column state and the first prior-gradient term (gibbs.cpp:267-272, 281)
  for (int e = tid; e < u; e += kST) {
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
  for (int e = tid; e < KT * n_e; e += kST) {
    const int k = e / n_e, i = e - k * n_e;
    Rs[e] = i < n ? b.R[(int64_t)k * n_s + i] : 0.0;
  }
  // lv phase ownership: thread -> (row(s), chunk); per-row constants in registers
  const int ch = CH > 1 ? tid / n32 : 0;
  const int row_first = CH > 1 ? tid - ch * n32 : tid;
  const bool owner = ch == 0;    // applies the softmax
  double fixv[kRptMax][KT], yv[kRptMax][KT];   // fixv: the row's running linear predictor
  int ybv[kRptMax];
#pragma unroll
  for (int q = 0; q < kRptMax; ++q) {
    const int i = row_first + q * kST;
    ybv[q] = -1;
#pragma unroll
    for (int k = 0; k < KT; ++k) { fixv[q][k] = 0; yv[q][k] = 0; }
    if (owner && i < n && (q == 0 || CH == 1)) {
      ybv[q] = b.ybase[i];
#pragma unroll
      for (int k = 0; k < KT; ++k) { fixv[q][k] = b.lv[(int64_t)k * n_s + i]; yv[q][k] = b.ymat[(int64_t)k * n_s + i]; }
    }
  }
  // columns of this thread's chunk in the lv phase
  const int cpc = (u + CH - 1) / CH;
  const int cl0 = min(u, ch * cpc), cl1 = min(u, cl0 + cpc);
  // wait for the columns
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok)
          : "r"(s_u32(bar)), "r"(0)
          : "memory");
    }
  }
  __syncthreads();

  double ll_total = 0;
  for (int s = 0; s <= L; ++s) {
    // ================= gradient: warp <-> kCI columns at a time (independent chains hide the latencies) =========
    // the warp's kCI columns of a round are NW apart, so few columns still spread over all warps
    for (int c0 = w; c0 < u; c0 += NW * kCI) {
      double acc[kCI][KT];
#pragma unroll
      for (int ci = 0; ci < kCI; ++ci)
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[ci][k] = 0;
      for (int m = lane; m < npairs; m += 32) {
        double2 r[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) r[k] = *reinterpret_cast<const double2*>(Rs + k * n_e + 2 * m);   // pad row holds 0
#pragma unroll
        for (int ci = 0; ci < kCI; ++ci) {
          if (c0 + NW * ci < u) {
            const double2 x = *reinterpret_cast<const double2*>(Xs + (size_t)(c0 + NW * ci) * n_e + 2 * m);
#pragma unroll
            for (int k = 0; k < KT; ++k) acc[ci][k] = fma(x.x, r[k].x, fma(x.y, r[k].y, acc[ci][k]));
          }
        }
      }
      // xor-shuffle trees of the columns the warp really has, interleaved (one fully unrolled variant per count, chosen
      // by a warp-uniform branch): an SM retires about one warp shuffle per cycle, and 16 warps reducing kCI x K values
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
      if (lane == 0) {
#pragma unroll
        for (int ci = 0; ci < kCI; ++ci)
          if (c0 + NW * ci < u) {
#pragma unroll
            for (int k = 0; k < KT; ++k) gsm[(size_t)(c0 + NW * ci) * KT + k] = acc[ci][k];
          }
      }
    }
    __syncthreads();
    // ================= update: thread <-> column (DNlogpost, MoveMomt, UpdateMomtAndDeltas, next prior gradient) ====
    for (int c = tid; c < u; c += kST) {
      double* sc = st + (size_t)c * CS;
      const double stp = sc[3 * KT + 1], hs = stp / 2, sig = sc[3 * KT];
      double dn[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const double post = gsm[(size_t)c * KT + k] + sc[2 * KT + k];
        const double kick = __dmul_rn(hs, post);
        double mm = sc[KT + k];
        if (s >= 1) mm = __dsub_rn(mm, kick);
        double dnew = sc[k];
        if (s < L) {
          mm = __dsub_rn(mm, kick);
          dnew = __dadd_rn(dnew, __dmul_rn(stp, mm));
        }
        sc[KT + k] = mm;
        dn[k] = dnew;
      }
      if (s < L) {
        double sd = dn[0];
#pragma unroll
        for (int k = 1; k < KT; ++k) sd += dn[k];
        const double sdc = sd / Cd;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          sc[3 * KT + 2 + k] = __dsub_rn(dn[k], sc[k]);   // what the lv phase adds
          sc[k] = dn[k];
          sc[2 * KT + k] = (dn[k] - sdc) / sig;
        }
      }
    }
    __syncthreads();
    if (s == L) break;

    // ================= linear predictor, softmax, residual =================
    const bool last = (s == L - 1);
    double ll = 0;
#pragma unroll
    for (int q = 0; q < kRptMax; ++q) {
      const int i = row_first + q * kST;
      if (q > 0 && CH > 1) break;
      const bool have_row = i < n && (CH == 1 || tid < CH * n32);
      double a[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) a[k] = 0;
      if (have_row) {
        for (int c = cl0; c < cl1; ++c) {
          const double x = Xs[(size_t)c * n_e + i];
          const double* sc = st + (size_t)c * CS + 3 * KT + 2;
#pragma unroll
          for (int k = 0; k < KT; ++k) a[k] = fma(x, sc[k], a[k]);
        }
      }
      if (CH > 1) {
        if (have_row) {
#pragma unroll
          for (int k = 0; k < KT; ++k) part[((size_t)k * CH + ch) * n32 + i] = a[k];
        }
        __syncthreads();
        if (have_row && owner) {
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            double t = part[(size_t)k * CH * n32 + i];
            for (int cc = 1; cc < CH; ++cc) t += part[((size_t)k * CH + cc) * n32 + i];
            a[k] = t;
          }
        }
      }
      if (have_row && owner) {
        double l[KT];
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k) { l[k] = fixv[q][k] + a[k]; fixv[q][k] = l[k]; m = fmax(m, l[k]); }
        double se = exp(0.0 - m);
#pragma unroll
        for (int k = 0; k < KT; ++k) se += exp(l[k] - m);
        const double lse = log(se) + m;
        if (ybv[q] == 0) ll += 0.0 - lse;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const double nl = l[k] - lse;
          if (ybv[q] == k + 1) ll += nl;
          const double res = exp(nl) - yv[q][k];
          Rs[k * n_e + i] = res;
          if (last) { b.lv[(int64_t)k * n_s + i] = l[k]; b.R[(int64_t)k * n_s + i] = res; }
        }
      }
    }
    if (last) ll_total = ll;
    __syncthreads();
  }

  // ---- log-likelihood of the proposal (fixed order: lanes by xor tree, then warps left to right)
  {
    double v = warp_sum(ll_total);
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int i = 0; i < NW; ++i) t += red[i];
      ctl->loglike_new = t;
      ctl->traj_done = 1;
      ctl->small_runs += 1;
    }
  }
  for (int e = tid; e < u; e += kST) {
    const double* s = st + (size_t)e * CS;
#pragma unroll
    for (int k = 0; k < KT; ++k) { b.dc[(int64_t)e * KT + k] = s[k]; b.momt[(int64_t)e * KT + k] = s[KT + k]; }
  }
}

namespace {
template <int KT>
cudaError_t launch_small_k(const Geom& g, const Bufs& b, int L, int smem_bytes, int u_max, cudaStream_t st) {
  static PerDeviceOnce configured;
  auto kern = k_traj_small<KT>;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    configured.mark();
  }
  kern<<<1, kST, smem_bytes, st>>>(g, b, L, smem_bytes / (int)sizeof(double), u_max);
  return cudaGetLastError();
}
}  // namespace

// The single-block kernel applies when a thread owns at most 2 rows (n <= 1024) and K <= 4.
bool small_applies(const Geom& g) { return g.K >= 1 && g.K <= 4 && g.n <= 2 * kST; }

cudaError_t launch_traj_small(const Geom& g, const Bufs& b, int L, int smem_bytes, int u_max, cudaStream_t st) {
  switch (g.K) {
    case 1: return launch_small_k<1>(g, b, L, smem_bytes, u_max, st);
    case 2: return launch_small_k<2>(g, b, L, smem_bytes, u_max, st);
    case 3: return launch_small_k<3>(g, b, L, smem_bytes, u_max, st);
    case 4: return launch_small_k<4>(g, b, L, smem_bytes, u_max, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace htlr
