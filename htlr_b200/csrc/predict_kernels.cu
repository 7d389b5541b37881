// On-device prediction from the resident trace: replaces the arithmetic of the
// reference's `htlr_predict` (R/predict.R:28-92) —
//   longlv   = cbind(1, X_ts) %*% matrix(mcdeltas[, , usedmc], nrow = p + 1)        (R/predict.R:77-78)
//   normlv   = lv - log_sum_exp(cbind(0, lv))  per case and posterior sample        (comp_lsl, R/util.r)
//   probs    = mean over samples of exp(normlv);  class-1 column = pmax(0, 1 - rowSums)   (:82-86)
// so the (p+1) x K x H cube never leaves HBM (config C: 0.5 GB for 1600 used samples).
//
// This is the one dense contraction of the path (n_ts x (p+1) x K*S): an FP64 shared-memory tiled GEMM
// (64 x 64 tiles, 4 x 4 per thread, k in ascending order: deterministic), one column per (sample, class),
// followed by a row-wise kernel that turns linear predictors into averaged probabilities. Samples are
// processed in chunks so the intermediate stays small.
#include <algorithm>

#include "common.cuh"
#include "launch.h"

namespace htlr {

namespace {

constexpr int kTM = 64, kTN = 64, kTK = 16;

// LV[i, j] = sum_r A[i, r] * D_j[r],  j = (t, k) -> column k of trace slice first + (t0 + t) * thin
__global__ void __launch_bounds__(256) k_predict_gemm(int M, int nvar, int K, int ncolsN, const double* __restrict__ A,
                                                       int64_t lda, const double* __restrict__ mcdeltas, int first,
                                                       int thin, int t0, double* __restrict__ LV) {
  __shared__ double As[kTK][kTM + 1];
  __shared__ double Bs[kTK][kTN + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.x * kTM, j0 = blockIdx.y * kTN;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0;
  for (int r0 = 0; r0 < nvar; r0 += kTK) {
    // A tile: 64 rows x 16 k (coalesced over rows)
    for (int e = threadIdx.x; e < kTM * kTK; e += 256) {
      const int ii = e % kTM, rr = e / kTM;
      const int i = i0 + ii, r = r0 + rr;
      As[rr][ii] = (i < M && r < nvar) ? A[(int64_t)r * lda + i] : 0.0;
    }
    // B tile: 16 k x 64 columns (coalesced over k)
    for (int e = threadIdx.x; e < kTK * kTN; e += 256) {
      const int rr = e % kTK, jj = e / kTK;
      const int j = j0 + jj, r = r0 + rr;
      double v = 0.0;
      if (j < ncolsN && r < nvar) {
        const int t = j / K, k = j - t * K;
        v = mcdeltas[((int64_t)(first + (int64_t)(t0 + t) * thin) * K + k) * nvar + r];
      }
      Bs[rr][jj] = v;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kTK; ++rr) {
      double a[4], bb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[rr][tx + 16 * q]; bb[q] = Bs[rr][ty + 16 * q]; }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[q][c] = fma(a[q], bb[c], acc[q][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + tx + 16 * q, j = j0 + ty + 16 * c;
      if (i < M && j < ncolsN) LV[(int64_t)j * M + i] = acc[q][c];
    }
}

// acc[i, k] += sum over the chunk's samples of exp(lv[i, t, k] - lsl[i, t]); one thread per case
__global__ void __launch_bounds__(256) k_predict_accum(int M, int K, int nt, const double* __restrict__ LV,
                                                        double* __restrict__ accum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double a[kMaxK];
  for (int k = 0; k < K; ++k) a[k] = accum[(int64_t)k * M + i];
  for (int t = 0; t < nt; ++t) {
    double l[kMaxK];
    double m = 0.0;
    for (int k = 0; k < K; ++k) { l[k] = LV[((int64_t)t * K + k) * M + i]; m = fmax(m, l[k]); }
    double se = exp(0.0 - m);
    for (int k = 0; k < K; ++k) se += exp(l[k] - m);
    const double lsl = log(se) + m;
    for (int k = 0; k < K; ++k) a[k] += exp(l[k] - lsl);
  }
  for (int k = 0; k < K; ++k) accum[(int64_t)k * M + i] = a[k];
}

// probs n_ts x C col-major: columns 1..K = accum / S, column 0 = max(0, 1 - row sum)
__global__ void k_predict_finish(int M, int K, int S, const double* __restrict__ accum, double* __restrict__ probs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double sum = 0;
  for (int k = 0; k < K; ++k) {
    const double pk = accum[(int64_t)k * M + i] / (double)S;
    probs[(int64_t)(k + 1) * M + i] = pk;
    sum += pk;
  }
  probs[i] = fmax(0.0, 1.0 - sum);
}

}  // namespace

int launch_predict(const Geom& g, const Bufs& b, int M, const double* A, int64_t lda, int first, int thin, int S,
                   double* LV, int chunk, double* accum, double* probs, cudaStream_t st) {
  int launches = 0;
  cudaMemsetAsync(accum, 0, sizeof(double) * (size_t)M * g.K, st);
  for (int t0 = 0; t0 < S; t0 += chunk) {
    const int nt = std::min(chunk, S - t0);
    const int N = nt * g.K;
    dim3 grid((M + kTM - 1) / kTM, (N + kTN - 1) / kTN);
    k_predict_gemm<<<grid, 256, 0, st>>>(M, g.nvar, g.K, N, A, lda, b.mcdeltas, first, thin, t0, LV);
    k_predict_accum<<<(M + 255) / 256, 256, 0, st>>>(M, g.K, nt, LV, accum);
    launches += 2;
  }
  k_predict_finish<<<(M + 255) / 256, 256, 0, st>>>(M, g.K, S, accum, probs);
  return launches + 1;
}

}  // namespace htlr
