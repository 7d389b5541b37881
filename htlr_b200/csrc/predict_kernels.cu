// On-device prediction from the resident trace: replaces the arithmetic of the
// reference's `htlr_predict` (R/predict.R:28-92) —
//   longlv   = cbind(1, X_ts) %*% matrix(mcdeltas[, , usedmc], nrow = p + 1)        (R/predict.R:77-78)
//   normlv   = lv - log_sum_exp(cbind(0, lv))  per case and posterior sample        (comp_lsl, R/util.r)
//   probs    = mean over samples of exp(normlv);  class-1 column = pmax(0, 1 - rowSums)   (:82-86)
// so the (p+1) x K x H cube never leaves HBM (config C: 0.5 GB for 1600 used samples).
//
// This is the one dense contraction of the path (n_ts x (p+1) x K*S; config C with 1000 test cases and 900 samples:
// 72 GFLOP in FP64, compute-bound: 0.5 GB of trace is 80 us of HBM time). It runs on the FP64 tensor-core path,
// mma.sync.m8n8k4.f64 (DMMA): with plain DFMA a thread needs one shared-memory operand per 4-8 FMAs and the SM's
// shared-memory pipe saturates before its FP64 pipe; a DMMA takes its 8x4 / 4x8 operands from registers and a warp
// reuses every fragment four times. Block tile 128 x 64, 8 warps of 32 x 32 (4 x 4 DMMA tiles, 32 accumulators per
// thread), k tiles of 16 double-buffered with cp.async (zero-filled past the edges), k in ascending order:
// deterministic. A second kernel turns linear predictors into averaged probabilities row by row. Samples are
// processed in chunks so the intermediate stays small.
#include <algorithm>

#include "common.cuh"
#include "launch.h"

namespace htlr {

namespace {

constexpr int kTM = 128, kTN = 64, kTK = 16;
constexpr int kLdA = kTM + 8;    // As[k][i]: rows 8 doubles apart modulo the 32 banks (a 64-bit warp load is 2 wavefronts anyway)
constexpr int kLdB = kTK + 4;    // Bs[j][k]

__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gsrc, int bytes, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int src = valid ? bytes : 0;   // src-size 0: the destination is zero-filled, the source is not touched
  if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src) : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// LV[i, j] = sum_r A[i, r] * D_j[r],  j = (t, k) -> column k of trace slice first + (t0 + t) * thin.
// A: M x nvar column-major, lda even and A 16-byte aligned (the caller pads).
__global__ void __launch_bounds__(256) k_predict_gemm(int M, int nvar, int K, int ncolsN, const double* __restrict__ A,
                                                       int64_t lda, const double* __restrict__ mcdeltas, int first,
                                                       int thin, int t0, double* __restrict__ LV) {
  extern __shared__ __align__(16) double pg_smem[];   // 54 KB: past the 48 KB a kernel gets without opting in
  double (*As)[kTK][kLdA] = reinterpret_cast<double (*)[kTK][kLdA]>(pg_smem);
  double (*Bs)[kTN][kLdB] = reinterpret_cast<double (*)[kTN][kLdB]>(pg_smem + 2 * kTK * kLdA);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * kTM, j0 = blockIdx.y * kTN;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;   // this warp's 32 x 32 corner inside the block tile
  // column base pointers of this thread's B copies (4 per k tile: columns jj = tid / 8 + 32 q?, see below)
  auto load_tiles = [&](int buf, int r0) {
    // A tile: 16 k x 128 i, 16-byte copies (2 rows of i): 1024 copies, 4 per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + 256 * q, rr = e >> 6, ii = (e & 63) * 2;
      const int i = i0 + ii, r = r0 + rr;
      const bool ok = r < nvar && i < M;    // (lda is even and rows beyond M inside the padded column hold zeros)
      cp_async_zfill(&As[buf][rr][ii], A + (int64_t)(ok ? r : 0) * lda + (ok ? i : 0), 16, ok);
    }
    // B tile: 64 columns x 16 k, 8-byte copies (a trace column starts at an odd multiple of 8 bytes when nvar is odd)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + 256 * q, jj = e >> 4, rr = e & 15;
      const int j = j0 + jj, r = r0 + rr;
      const bool ok = j < ncolsN && r < nvar;
      const double* src = mcdeltas;
      if (ok) {
        const int t = j / K, k = j - t * K;
        src = mcdeltas + ((int64_t)(first + (int64_t)(t0 + t) * thin) * K + k) * nvar + r;
      }
      cp_async_zfill(&Bs[buf][jj][rr], src, 8, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[a][c][0] = 0; acc[a][c][1] = 0; }
  const int nkt = (nvar + kTK - 1) / kTK;
  load_tiles(0, 0);
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) load_tiles(buf ^ 1, (kt + 1) * kTK);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kTK; ks += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        af[q] = As[buf][ks + (lane & 3)][wm + 8 * q + (lane >> 2)];   // A fragment: row lane/4, k lane%4
        bf[q] = Bs[buf][wn + 8 * q + (lane >> 2)][ks + (lane & 3)];   // B fragment: k lane%4, column lane/4
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) dmma(acc[a][c], af[a], bf[c]);
    }
    __syncthreads();
  }
  // C fragment: row lane/4, columns (lane%4)*2 + {0, 1}
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = i0 + wm + 8 * a + (lane >> 2), j = j0 + wn + 8 * c + (lane & 3) * 2 + h;
        if (i < M && j < ncolsN) LV[(int64_t)j * M + i] = acc[a][c][h];
      }
}

// FP64 peak of the device, measured (bench.py's prediction block states its fraction of it): independent DFMA chains
// and independent DMMA chains, nothing else in the loop.
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, int use_dmma, double* sink) {
  double a[8][2];
#pragma unroll
  for (int q = 0; q < 8; ++q) { a[q][0] = 1e-3 * (threadIdx.x + q); a[q][1] = 2e-3 * q; }
  const double x = 1.0000001, y = 1e-9 * threadIdx.x;
  if (use_dmma) {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int q = 0; q < 8; ++q) dmma(a[q], x, y);
  } else {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int q = 0; q < 8; ++q) { a[q][0] = fma(a[q][0], x, y); a[q][1] = fma(a[q][1], x, y); }
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += a[q][0] + a[q][1];
  if (s == 123.456) sink[0] = s;
}

// acc[i, k] += sum over the chunk's samples of exp(lv[i, t, k] - lsl[i, t]); one thread per case
__global__ void __launch_bounds__(256) k_predict_accum(int M, int K, int nt, const double* __restrict__ LV,
                                                        double* __restrict__ accum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  if (K > kMaxK) {   // many classes: no register arrays, three passes over the classes per sample
    for (int t = 0; t < nt; ++t) {
      double m = 0.0;
      for (int k = 0; k < K; ++k) m = fmax(m, LV[((int64_t)t * K + k) * M + i]);
      double se = exp(0.0 - m);
      for (int k = 0; k < K; ++k) se += exp(LV[((int64_t)t * K + k) * M + i] - m);
      const double lsl = log(se) + m;
      for (int k = 0; k < K; ++k) accum[(int64_t)k * M + i] += exp(LV[((int64_t)t * K + k) * M + i] - lsl);
    }
    return;
  }
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

// TFLOP/s of `iters` rounds of 8 independent chains per thread on every SM (FMA = 2 flops; a DMMA m8n8k4 = 512).
void launch_fp64_peak(int sm_count, int iters, int use_dmma, double* sink, cudaStream_t st) {
  k_fp64_peak<<<sm_count * 8, 256, 0, st>>>(iters, use_dmma, sink);
}

int launch_predict(const Geom& g, const Bufs& b, int M, const double* A, int64_t lda, int first, int thin, int S,
                   double* LV, int chunk, double* accum, double* probs, cudaStream_t st) {
  int launches = 0;
  constexpr size_t kGemmSmem = sizeof(double) * (2 * kTK * kLdA + 2 * kTN * kLdB);
  static PerDeviceOnce configured;
  if (configured.needed()) {
    cudaFuncSetAttribute(k_predict_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
    configured.mark();
  }
  cudaMemsetAsync(accum, 0, sizeof(double) * (size_t)M * g.K, st);
  for (int t0 = 0; t0 < S; t0 += chunk) {
    const int nt = std::min(chunk, S - t0);
    const int N = nt * g.K;
    dim3 grid((M + kTM - 1) / kTM, (N + kTN - 1) / kTN);
    k_predict_gemm<<<grid, 256, kGemmSmem, st>>>(M, g.nvar, g.K, N, A, lda, b.mcdeltas, first, thin, t0, LV);
    k_predict_accum<<<(M + 255) / 256, 256, 0, st>>>(M, g.K, nt, LV, accum);
    launches += 2;
  }
  k_predict_finish<<<(M + 255) / 256, 256, 0, st>>>(M, g.K, S, accum, probs);
  return launches + 1;
}

}  // namespace htlr
