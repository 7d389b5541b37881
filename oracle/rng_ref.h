// TEST INFRASTRUCTURE — part of the CPU oracle, never shipped or measured as product.
//
// Counter-based random numbers for the oracle sampler. The reference draws from R's global
// Mersenne-Twister stream (R::rnorm gibbs.cpp:450, R::runif gibbs.cpp:90, Rcpp::rgamma
// utils.cpp:31, unif_rand ars.cpp:284,344, R::runif ars.cpp:402); R is not available and a serial
// stream cannot be reproduced on a GPU, so the oracle and the CUDA library share this *keyed* scheme
// instead: every variate is a pure function of (seed, purpose, thin-step, feature, sub-index).
// The same law as the reference, a different stream ("parity unpinned" for the RNG stream itself,
// see DESIGN.md); exact cross-checks use injected variates instead.
//
// Philox4x32-10 (Salmon et al., SC'11). The device twin lives in htlr_b200/csrc/rng.cuh; the two are
// written independently and cross-checked by tests/test_rng.py.
#pragma once
#include <cmath>
#include <cstdint>

namespace oracle_rng {

enum Purpose : uint32_t {
  kMomentum = 0,   // (step, j, k)         -> N(0,1)
  kAccept = 1,     // (step)               -> U(0,1)
  kGamma = 2,      // (step, j, attempt)   -> Gamma(shape,1) via Marsaglia-Tsang
  kArs = 3,        // (step, j, draw)      -> U(0,1)
  kLogwArs = 4,    // (step, draw)         -> U(0,1)
};

struct U4 { uint32_t v[4]; };

inline U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * ctr.v[0];
    uint64_t p1 = (uint64_t)M1 * ctr.v[2];
    U4 n;
    n.v[0] = (uint32_t)(p1 >> 32) ^ ctr.v[1] ^ k0;
    n.v[1] = (uint32_t)p1;
    n.v[2] = (uint32_t)(p0 >> 32) ^ ctr.v[3] ^ k1;
    n.v[3] = (uint32_t)p0;
    ctr = n;
    k0 += W0;
    k1 += W1;
  }
  return ctr;
}

// 52 random bits -> (m + 0.5) * 2^-52, exactly representable, strictly inside (0,1).
inline double bits_to_u01(uint32_t hi, uint32_t lo) {
  uint64_t m = (((uint64_t)hi << 32) | lo) >> 12;
  return ((double)m + 0.5) * 2.220446049250313e-16;
}

struct Keyed {
  uint64_t seed = 0;

  U4 block(uint32_t purpose, uint64_t step, uint32_t a, uint32_t b) const {
    // counter = (a, b, purpose | step_hi << 8, step_lo); key = seed
    U4 c;
    c.v[0] = a;
    c.v[1] = b;
    c.v[2] = purpose | ((uint32_t)(step >> 32) << 8);
    c.v[3] = (uint32_t)step;
    return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  }
  double unif(uint32_t purpose, uint64_t step, uint32_t a, uint32_t b) const {
    U4 r = block(purpose, step, a, b);
    return bits_to_u01(r.v[0], r.v[1]);
  }
  // Box-Muller, cosine branch only.
  double norm(uint32_t purpose, uint64_t step, uint32_t a, uint32_t b) const {
    U4 r = block(purpose, step, a, b);
    double u1 = bits_to_u01(r.v[0], r.v[1]);
    double u2 = bits_to_u01(r.v[2], r.v[3]);
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  // Marsaglia & Tsang (2000) for shape >= 1; shape < 1 via Gamma(shape+1) * U^(1/shape).
  // Attempt t uses sub-counter b = t (normal) and b = t | 0x80000000 (uniform); the boost
  // uniform uses b = 0x7fffffff.
  double gamma(uint64_t step, uint32_t j, double shape) const {
    double a = shape < 1.0 ? shape + 1.0 : shape;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / std::sqrt(9.0 * d);
    double g = 0.0;
    for (uint32_t t = 0;; ++t) {
      double x = norm(kGamma, step, j, t);
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      double u = unif(kGamma, step, j, t | 0x80000000u);
      if (std::log(u) < 0.5 * x * x + d - d * v + d * std::log(v)) {
        g = d * v;
        break;
      }
    }
    if (shape < 1.0) {
      double u = unif(kGamma, step, j, 0x7fffffffu);
      g *= std::pow(u, 1.0 / shape);
    }
    return g;
  }
};

}  // namespace oracle_rng
