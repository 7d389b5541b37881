// Counter-based random variates on the device: Philox4x32-10 keyed by
// (seed, purpose, thin-step, feature, sub-index). A variate never depends on the order in which the
// restricted set is enumerated or on which thread/rank asks for it, so chains are reproducible and
// row-shard ranks stay bit-identical without a broadcast.
//
// Replaces the reference's use of R's global RNG stream: R::rnorm (src/gibbs.cpp:450), R::runif
// (gibbs.cpp:90), Rcpp::rgamma / R::rgamma (utils.cpp:31, gibbs.cpp:320), unif_rand / R::runif in the
// ARS engine (ars.cpp:284,344,402). Same laws, different stream.
#pragma once
#include <cstdint>

namespace htlr {

enum RngPurpose : uint32_t { kRngMomentum = 0, kRngAccept = 1, kRngGamma = 2, kRngArs = 3, kRngLogwArs = 4 };

struct PhiloxOut { uint32_t x, y, z, w; };

__device__ __forceinline__ PhiloxOut philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return PhiloxOut{c0, c1, c2, c3};
}

// 52 random bits -> (m + 0.5) * 2^-52: exact, strictly inside (0, 1).
__device__ __forceinline__ double bits_to_u01(uint32_t hi, uint32_t lo) {
  uint64_t m = (((uint64_t)hi << 32) | lo) >> 12;
  return ((double)m + 0.5) * 2.220446049250313e-16;
}

struct KeyedRng {
  uint64_t seed;
  __device__ __forceinline__ PhiloxOut block(uint32_t purpose, uint64_t step, uint32_t a, uint32_t b) const {
    return philox4x32_10(a, b, purpose | ((uint32_t)(step >> 32) << 8), (uint32_t)step, (uint32_t)seed,
                         (uint32_t)(seed >> 32));
  }
  __device__ __forceinline__ double unif(uint32_t purpose, uint64_t step, uint32_t a, uint32_t b) const {
    PhiloxOut r = block(purpose, step, a, b);
    return bits_to_u01(r.x, r.y);
  }
  __device__ __forceinline__ double norm(uint32_t purpose, uint64_t step, uint32_t a, uint32_t b) const {
    PhiloxOut r = block(purpose, step, a, b);
    double u1 = bits_to_u01(r.x, r.y), u2 = bits_to_u01(r.z, r.w);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  }
  // Marsaglia & Tsang (2000); shape < 1 boosted by U^(1/shape). Sub-counter scheme documented in
  // oracle/rng_ref.h (the independently written host twin).
  __device__ double gamma(uint64_t step, uint32_t j, double shape) const {
    double a = shape < 1.0 ? shape + 1.0 : shape;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double g = 0.0;
    for (uint32_t t = 0;; ++t) {
      double x = norm(kRngGamma, step, j, t);
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      double u = unif(kRngGamma, step, j, t | 0x80000000u);
      if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
        g = d * v;
        break;
      }
    }
    if (shape < 1.0) g *= pow(unif(kRngGamma, step, j, 0x7fffffffu), 1.0 / shape);
    return g;
  }
};

}  // namespace htlr
