// Adaptive rejection sampling, one GROUP of W lanes per draw (W = 32: a warp; W = 16: two draws per warp): lane h of
// the group owns tangent hull h (so at most W hulls — the reference allows 1000, `ars.h:61`, but its own targets
// never need more than ~13; with a full table insertion simply stops, exactly what the reference does when
// `no_hulls_ == max_nhull_`, ars.cpp:26, so this is the reference engine run with max_nhull = W). Every scalar of
// the engine is computed redundantly by all lanes of the group and the FP64 pipe is what bounds the kernel, so
// two draws per warp (the halves run the same code on different features, diverging only when their accept /
// reject histories differ) nearly halve the cost per draw.
//
// Replaces, for n = 1 draw (the only way gibbs.cpp uses it, gibbs.cpp:344-345, 358, 367):
//   ARS::ARS         src/ars.cpp:204-242   first tangent at x0, error if logf(x0) is not finite
//   ARS::Initialize  src/ars.cpp:138-202   step out left/right by 10 until the slope changes sign
//   update_hulls     src/ars.cpp:22-119    insert a tangent, fix the two neighbours
//   ARS::Sample      src/ars.cpp:259-318   pick hull, pick point, squeeze test, accept test
//   sample_disc / sample_elin / logint_elin / interc   src/ars.cpp:334-462
// All scalar quantities are group-uniform (every lane evaluates the same expression on the same
// operands); per-hull fields live one per lane and are broadcast with shuffles inside the group. The cumulative hull
// weights are summed left to right like the reference so that, fed the same uniforms, the draw
// matches the CPU engine to the last few ulps of libm.
//
// The log-density is a functor `eval(x, logf, dlogf)` that must be warp-uniform; for the logw target it
// is a block-collective reduction over the p features (all warps then run the engine redundantly).
#pragma once
#include <cstdint>

#include "rng.cuh"

namespace htlr {

enum ArsStatus : int {
  kArsOk = 0,
  kArsBadStart = 1,      // "the first tangent point doesn't have positive probability" (ars.cpp:217-222)
  kArsHullOverflow = 2,  // step-out ran out of hulls (ars.cpp:155-160, 187-193)
  kArsBadPiece = 3,      // "the exp linear function integrates to NAN/INFINITY" (ars.cpp:391-398)
  kArsNoUniforms = 4,    // injected uniform block exhausted
  kArsStuck = 5,         // guard: more than kArsMaxTries proposals
  kArsOkTableFull = 64,  // standalone hook only: a valid draw, but the hull table filled up on the way (see insert())
};

constexpr int kArsMaxTries = 4096;
constexpr double kArsStepout = 10.0, kArsTolD = 1e-5, kArsTolDD = 1e-5;

// exp / log are called from ~40 places in the engine and the targets; inlining them all makes a 220 KB kernel
// whose warps (each at a different point of its own accept/reject loop) thrash the instruction cache
// (ncu: 43 % of stalls were `no_instruction`). One out-of-line copy of each keeps the loop cache-resident.
__device__ __noinline__ double ars_exp(double x) { return exp(x); }
__device__ __noinline__ double ars_log(double x) { return log(x); }


// ars.cpp:424-450
__device__ __forceinline__ double ars_logint(double f, double df, double t, double lo, double hi) {
  double dx = hi - lo;
  double adf = fabs(df);
  if (adf <= kArsTolD) return f + ars_log(dx);
  double edge = df > kArsTolD ? hi : lo;
  return f + df * (edge - t) - ars_log(adf) + ars_log(1 - ars_exp(-adf * dx));
}
// ars.cpp:453-462
__device__ __forceinline__ double ars_cross(double t1, double t2, double f1, double f2, double d1, double d2) {
  if (fabs(d1 - d2) > kArsTolDD) return (f2 - f1 - d2 * t2 + d1 * t1) / (d1 - d2);
  return (t1 + t2) / 2.0;
}

// Uniform sources (warp-uniform `next`).
struct ArsKeyedUniforms {
  uint64_t seed, step;
  uint32_t purpose, j, draw;
  __device__ __forceinline__ bool next(double& u) {
    PhiloxOut r = philox4x32_10(j, draw++, purpose | ((uint32_t)(step >> 32) << 8), (uint32_t)step,
                                (uint32_t)seed, (uint32_t)(seed >> 32));
    u = bits_to_u01(r.x, r.y);
    return true;
  }
};
struct ArsArrayUniforms {
  const double* u;
  int n, c;
  __device__ __forceinline__ bool next(double& v) {
    if (c >= n) return false;
    v = u[c++];
    return true;
  }
};

// One source type for both random-number modes (a runtime switch instead of two instantiations of the engine).
struct ArsUniforms {
  bool inject;          // host-supplied block (u, n) or the keyed Philox stream
  const double* u;
  int n, c;
  ArsKeyedUniforms keyed;
  __device__ __forceinline__ bool next(double& v) {
    if (inject) {
      if (c >= n) return false;
      v = u[c++];
      return true;
    }
    return keyed.next(v);
  }
};

struct ArsHull {
  double t, f, df, lo, hi, lw, sl, sr;
  int left, right;
};

template <class Eval, class Unif, int W = 32>
struct ArsWarp {
  static_assert(W == 32 || W == 16, "a group is a warp or half a warp");
  static constexpr int kArsCap = W;   // hull table size; also the "no right neighbour" sentinel
  ArsHull H;  // this lane's hull
  int nh;
  int lane;        // lane within the group
  int base;        // first lane of the group within the warp
  unsigned mask;   // the group's lanes
  Eval& ev;
  Unif& un;
  int n_eval, n_unif;
  bool full;       // the table filled up during this draw: later rejected points were not inserted

  __device__ ArsWarp(Eval& e, Unif& u) : ev(e), un(u), full(false) {
    lane = threadIdx.x & (W - 1);
    base = (threadIdx.x & 31) & ~(W - 1);
    mask = W == 32 ? 0xffffffffu : (0xffffu << base);
    nh = 0;
    n_eval = n_unif = 0;
  }
  __device__ __forceinline__ double shfl_d(double v, int src) const { return __shfl_sync(mask, v, src, W); }
  __device__ __forceinline__ int shfl_i(int v, int src) const { return __shfl_sync(mask, v, src, W); }

  __device__ __forceinline__ void relog() { H.lw = ars_logint(H.f, H.df, H.t, H.lo, H.hi); }

  // update_hulls(h, x, f, df), ars.cpp:22-119
  // The reference's table holds max_nhull = 1000 hulls and update_hulls() returns silently once it is full
  // (ars.cpp:26): the sampler stays exact, the envelope just stops tightening. Same semantics here with a table of
  // W hulls (16 per feature, 32 for logw; the reference's targets need at most 13 - asserted in the tests); draws that
  // got there are counted (Ctl::ars_full, htlr_cuda_ars_table_full) instead of passing unnoticed.
  __device__ void insert(int h, double x, double f, double df) {
    if (nh == kArsCap) { full = true; return; }
    const double th = shfl_d(H.t, h);
    int lh, rh;
    if (x > th) {
      lh = h;
      rh = shfl_i(H.right, h);
      if ((rh == kArsCap) & (f == -INFINITY)) {
        if (lane == h && H.hi != x) { H.hi = x; relog(); }
        return;
      }
    } else {
      lh = shfl_i(H.left, h);
      rh = h;
      if ((lh == -1) & (f == -INFINITY)) {
        if (lane == h && H.lo != x) { H.lo = x; relog(); }
        return;
      }
    }
    const int ni = nh++;
    // neighbours' tangent data (clamped source lanes; values unused when the neighbour is absent)
    const int ls = lh < 0 ? 0 : lh, rs = rh >= kArsCap ? 0 : rh;
    const double lt = shfl_d(H.t, ls), lf = shfl_d(H.f, ls), ld = shfl_d(H.df, ls);
    const double rt = shfl_d(H.t, rs), rf = shfl_d(H.f, rs), rd = shfl_d(H.df, rs);
    const double h_lo = shfl_d(H.lo, h), h_hi = shfl_d(H.hi, h);
    double nlo, nsl, nhi, nsr;
    if (lh == -1) { nlo = h_lo; nsl = INFINITY; }
    else { nlo = ars_cross(lt, x, lf, f, ld, df); nsl = (f - lf) / (x - lt); }
    if (rh == kArsCap) { nhi = h_hi; nsr = -INFINITY; }
    else { nhi = ars_cross(x, rt, f, rf, df, rd); nsr = (f - rf) / (x - rt); }
    if (lane == ni) {
      H.t = x; H.f = f; H.df = df; H.left = lh; H.right = rh;
      H.lo = nlo; H.sl = nsl; H.hi = nhi; H.sr = nsr;
    }
    if (lane == lh) { H.hi = nlo; H.right = ni; H.sr = nsl; }
    if (lane == rh) { H.lo = nhi; H.left = ni; H.sl = nsr; }
    if (lane == ni || lane == lh || lane == rh) relog();
  }

  // Returns status; x_out is the draw.
  __device__ int sample(double x0, double& x_out) {
    double f, df;
    // ---- ctor, ars.cpp:211-241
    ev(x0, f, df); ++n_eval;
    if (!isfinite(f)) return kArsBadStart;
    H.t = x0; H.f = f; H.df = df; H.lo = -INFINITY; H.hi = INFINITY;
    H.left = -1; H.right = kArsCap; H.sl = INFINITY; H.sr = -INFINITY; H.lw = INFINITY;
    if (lane != 0) { H.t = 0; H.f = 0; H.df = 0; H.lo = 0; H.hi = 0; H.lw = 0; H.sl = 0; H.sr = 0; H.left = 0; H.right = 0; }
    nh = 1;
    // ---- step out to the left, always from hull 0 (ars.cpp:152-169)
    double x = x0 - kArsStepout;
    do {
      if (nh == kArsCap) return kArsHullOverflow;
      ev(x, f, df); ++n_eval;
      insert(0, x, f, df);
      if (f == -INFINITY) break;
      x -= kArsStepout;
    } while (df < kArsTolD);
    // ---- step out to the right (ars.cpp:183-200)
    int h = 0;
    x = x0 + kArsStepout;
    do {
      if (nh == kArsCap) return kArsHullOverflow;
      ev(x, f, df); ++n_eval;
      insert(h, x, f, df);
      if (!isfinite(f)) break;
      x += kArsStepout;
      h = nh - 1;
    } while (df > -kArsTolD);

    // ---- rejection loop (ars.cpp:275-309)
    for (int tries = 0; tries < kArsMaxTries; ++tries) {
      // sample_disc, ars.cpp:334-356: max-shifted weights, cumulative sum left to right
      double mx = lane < nh ? H.lw : -INFINITY;
#pragma unroll
      for (int o = W / 2; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(mask, mx, o, W));
      double e = lane < nh ? ars_exp(H.lw - mx) : 0.0;
      double acc = 0.0, cw = 0.0;
      for (int i = 0; i < nh; ++i) {
        acc += shfl_d(e, i);
        if (lane == i) cw = acc;
      }
      double u;
      if (!un.next(u)) return kArsNoUniforms;
      ++n_unif;
      u *= acc;
      const unsigned hit = __ballot_sync(mask, lane < nh && u <= cw) >> base;   // bits of this group only
      h = hit ? (__ffs(hit) - 1) : nh - 1;
      // sample_elin, ars.cpp:359-417
      const double lo = shfl_d(H.lo, h), hi = shfl_d(H.hi, h), dh = shfl_d(H.df, h);
      const double th = shfl_d(H.t, h), fh = shfl_d(H.f, h);
      int type = -1;
      bool fault = false;
      if (fabs(dh) <= kArsTolD) { if (!(isfinite(lo) & isfinite(hi))) fault = true; else type = 0; }
      if (dh > kArsTolD) { if (!isfinite(hi)) fault = true; else type = 1; }
      if (dh < -kArsTolD) { if (!isfinite(lo)) fault = true; else type = 2; }
      if (fault || type < 0) return kArsBadPiece;
      const double dx = hi - lo;
      double y;
      if (!un.next(y)) return kArsNoUniforms;
      ++n_unif;
      if (type == 0) x = lo + y * dx;
      else if (type == 1) x = hi + ars_log((1 - y) * ars_exp(-dh * dx) + y) / dh;
      else x = lo + ars_log(1 - y + y * ars_exp(dh * dx)) / dh;
      const double upper = (x - th) * dh + fh;
      double ua;
      if (!un.next(ua)) return kArsNoUniforms;
      ++n_unif;
      const double logacc = upper + ars_log(ua);
      const double slope = x >= th ? shfl_d(H.sr, h) : shfl_d(H.sl, h);
      const double lower = (x - th) * slope + fh;
      if (logacc <= lower) { x_out = x; return kArsOk; }
      ev(x, f, df); ++n_eval;
      insert(h, x, f, df);
      if (logacc <= f) { x_out = x; return kArsOk; }
    }
    return kArsStuck;
  }
};

// ---- targets of sampler.cpp, x = log sigma^2_j --------------------------------------------------------
// NEG: src/sampler.cpp:18-30; GHS: src/sampler.cpp:37-49.
struct SgmTarget {
  int kind;  // 1 ghs, 2 neg
  double v, K, alpha, log_aw;
  __device__ __forceinline__ void operator()(double x, double& logf, double& dlogf) const {
    const double vexi = v / 2.0 / ars_exp(x);
    const double e = ars_exp(x - log_aw);
    if (kind == 2) {
      logf = -(K / 2.0 - 1.0) * x;
      dlogf = -(K / 2.0 - 1.0);
      logf -= vexi;
      dlogf += vexi;
      logf -= (alpha / 2.0 + 1.0) * ars_log(1.0 + 2.0 * e);
      dlogf -= (alpha + 2.0) * e / (1.0 + 2.0 * e);
    } else {
      logf = -(K - 1.0) / 2.0 * x;
      dlogf = -(K - 1.0) / 2.0;
      logf -= vexi;
      dlogf += vexi;
      logf -= (alpha + 1.0) / 2.0 * ars_log(1.0 + e);
      dlogf -= (alpha + 1.0) / 2.0 * e / (1.0 + e);
    }
  }
};

}  // namespace htlr
