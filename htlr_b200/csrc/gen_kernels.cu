// Synthetic-data generators on the device: the LAWS of the reference's gendata_MLR (R/gendata.r:26-46)
// and gendata_FAM (R/gendata.r:105-113 + gendata_FAM_helper, src/utils.cpp:72-105), written straight into the layout
// htlr_cuda_create takes with x_on_device (column-major n x (p+1) with the intercept column, one-hot ymat, ybase), so
// tall data never exists on the host. Variates are keyed Philox draws (seed, purpose, GLOBAL row, column): a row block
// [row0, row0 + n) generated on any GPU is the same slice of one data set, whatever the sharding.
//   gendata_MLR: X ~ N(0,1); sigmasbt_j = w / Gamma(nu/2, rate nu/2); betas = N(0,1) * c(0, sqrt(sigmasbt));
//                y ~ categorical(softmax(cbind(1, X) %*% betas))
//   gendata_FAM: y = rep(1:C, len = n); X = A z + muj[, y] + sd_g e with A = I_p whose leading pb x kb block is
//                replaced (the vignette's factor pattern, vignettes/simu.Rmd:49-84: A is never materialised);
//                stdx: (X - rowMeans(muj)) / sqrt(diag(A A') + sd_g^2 + var_pop(muj))
#include <algorithm>

#include "common.cuh"
#include "launch.h"
#include "rng.cuh"

namespace htlr {

namespace {

constexpr uint32_t kGenX = 8, kGenBeta = 9, kGenY = 11, kGenZ = 12, kGenE = 13;
constexpr unsigned long long kGenGammaStep = 0x6d6c7267616dULL;   // "mlrgam": step key of the sigmasbt gammas
constexpr int kGenChunk = 256;                                    // columns per block of the MLR generator
constexpr int kGenMaxC = kMaxK + 1;

__global__ void k_gen_betas(int p, int NC, double nu, double w, unsigned long long seed, double* __restrict__ betas) {
  const KeyedRng rng{seed};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= p; j += gridDim.x * blockDim.x) {
    // 1 / rgamma(p, nu/2, rate = nu/2) * w
    const double scale = j == 0 ? 0.0 : sqrt(w * (nu / 2) / rng.gamma(kGenGammaStep, (uint32_t)(j - 1), nu / 2));
    for (int c = 0; c < NC; ++c) betas[(int64_t)c * (p + 1) + j] = rng.norm(kGenBeta, (uint64_t)c, (uint32_t)j, 0u) * scale;
  }
}

// block (row tile of 256, column chunk): writes its tile of X and its partial of cbind(1, X) %*% betas
__global__ void __launch_bounds__(256) k_gen_mlr_x(int n, int p, int NC, long long row0, unsigned long long seed,
                                                    double* __restrict__ X, int64_t ld, const double* __restrict__ betas,
                                                    double* __restrict__ part) {
  __shared__ double sb[kGenChunk * kGenMaxC];
  const KeyedRng rng{seed};
  const int nvar = p + 1, j0 = blockIdx.y * kGenChunk, jn = min(kGenChunk, nvar - j0);
  for (int e = threadIdx.x; e < jn * NC; e += blockDim.x) {
    const int jj = e / NC, c = e - jj * NC;
    sb[e] = betas[(int64_t)c * nvar + j0 + jj];
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc[kGenMaxC];
#pragma unroll
  for (int c = 0; c < kGenMaxC; ++c) acc[c] = 0;
  for (int jj = 0; jj < jn; ++jj) {
    const int j = j0 + jj;
    const double x = j == 0 ? 1.0 : rng.norm(kGenX, (uint64_t)(row0 + i), (uint32_t)j, 0u);
    X[(int64_t)j * ld + i] = x;
#pragma unroll
    for (int c = 0; c < kGenMaxC; ++c)
      if (c < NC) acc[c] = fma(x, sb[jj * NC + c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < kGenMaxC; ++c)
    if (c < NC) part[((int64_t)blockIdx.y * NC + c) * n + i] = acc[c];
}

__global__ void __launch_bounds__(256) k_gen_mlr_y(int n, int NC, int nchunks, long long row0, unsigned long long seed,
                                                    const double* __restrict__ part, double* __restrict__ ymat,
                                                    long long* __restrict__ ybase) {
  const KeyedRng rng{seed};
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double lv[kGenMaxC], m = -INFINITY;
#pragma unroll
  for (int c = 0; c < kGenMaxC; ++c) {
    lv[c] = 0;
    if (c < NC) {
      for (int q = 0; q < nchunks; ++q) lv[c] += part[((int64_t)q * NC + c) * n + i];   // chunk order: deterministic
      m = fmax(m, lv[c]);
    }
  }
  double tot = 0;
#pragma unroll
  for (int c = 0; c < kGenMaxC; ++c)
    if (c < NC) { lv[c] = exp(lv[c] - m); tot += lv[c]; }
  const double u = rng.unif(kGenY, (uint64_t)(row0 + i), 0u, 0u) * tot;
  int y = NC - 1;
  double cum = 0;
#pragma unroll
  for (int c = 0; c < kGenMaxC; ++c)
    if (c < NC) { cum += lv[c]; if (u <= cum && y == NC - 1 && c < NC - 1) y = c; }
  ybase[i] = y;
  for (int k = 0; k < NC - 1; ++k) ymat[(int64_t)k * n + i] = (y == k + 1) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256) k_gen_fam(int n, int p, int NC, int pb, int kb, const double* __restrict__ muj,
                                                 const double* __restrict__ Ab, double sd_g, int stdx, long long row0,
                                                 unsigned long long seed, double* __restrict__ X, int64_t ld,
                                                 double* __restrict__ ymat, long long* __restrict__ ybase) {
  const KeyedRng rng{seed};
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long row = row0 + i;
  const int y = (int)(row % NC);
  if (blockIdx.y == 0) {
    ybase[i] = y;
    for (int k = 0; k < NC - 1; ++k) ymat[(int64_t)k * n + i] = (y == k + 1) ? 1.0 : 0.0;
    X[i] = 1.0;   // intercept column
  }
  const int j0 = blockIdx.y * kGenChunk, j1 = min(p, j0 + kGenChunk);
  for (int j = j0; j < j1; ++j) {
    double az, diag;
    if (j < pb) {
      az = 0; diag = 0;
      for (int f = 0; f < kb; ++f) {
        const double a = Ab[(int64_t)f * pb + j];
        az = fma(a, rng.norm(kGenZ, (uint64_t)row, (uint32_t)f, 0u), az);
        diag = fma(a, a, diag);
      }
      if (j >= kb) { az += rng.norm(kGenZ, (uint64_t)row, (uint32_t)j, 0u); diag += 1.0; }
    } else {
      az = rng.norm(kGenZ, (uint64_t)row, (uint32_t)j, 0u);
      diag = 1.0;
    }
    double x = az + muj[(int64_t)y * p + j] + sd_g * rng.norm(kGenE, (uint64_t)row, (uint32_t)j, 0u);
    if (stdx) {
      double mu = 0;
      for (int c = 0; c < NC; ++c) mu += muj[(int64_t)c * p + j];
      mu /= NC;
      double var = 0;
      for (int c = 0; c < NC; ++c) { const double d = muj[(int64_t)c * p + j] - mu; var = fma(d, d, var); }
      var /= NC;                                           // arma::var(muj, 1, 1): population variance
      x = (x - mu) / sqrt(diag + sd_g * sd_g + var);
    }
    X[(int64_t)(j + 1) * ld + i] = x;
  }
}

}  // namespace

void launch_gen_mlr(int n, int p, int NC, double nu, double w, unsigned long long seed, long long row0, double* X,
                    int64_t ld, double* ymat, long long* ybase, double* betas, double* part, cudaStream_t st) {
  const int nchunks = (p + 1 + kGenChunk - 1) / kGenChunk;
  k_gen_betas<<<std::max(1, std::min((p + 256) / 256, 1024)), 256, 0, st>>>(p, NC, nu, w, seed, betas);
  dim3 grid((n + 255) / 256, nchunks);
  k_gen_mlr_x<<<grid, 256, 0, st>>>(n, p, NC, row0, seed, X, ld, betas, part);
  k_gen_mlr_y<<<(n + 255) / 256, 256, 0, st>>>(n, NC, nchunks, row0, seed, part, ymat, ybase);
}
int gen_mlr_chunks(int p) { return (p + 1 + kGenChunk - 1) / kGenChunk; }

void launch_gen_fam(int n, int p, int NC, int pb, int kb, const double* muj, const double* Ab, double sd_g, int stdx,
                    unsigned long long seed, long long row0, double* X, int64_t ld, double* ymat, long long* ybase,
                    cudaStream_t st) {
  dim3 grid((n + 255) / 256, (p + kGenChunk - 1) / kGenChunk);
  k_gen_fam<<<grid, 256, 0, st>>>(n, p, NC, pb, kb, muj, Ab, sd_g, stdx, row0, seed, X, ld, ymat, ybase);
}

}  // namespace htlr
