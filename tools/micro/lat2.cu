// What makes the first global reads after a grid barrier slow? Variants of the spin / fence / load flavour.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ double ld_relaxed_d(const double* p) {
  double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v;
}
// SPIN: 0 acquire+nanosleep, 1 relaxed spin + one fence.acq_rel, 2 relaxed spin + one ld.acquire, 3 acquire spin no sleep
// SRC: 0 partials written this round by other blocks, 1 static buffer
// LD: 0 ld.cg, 1 ld.relaxed.gpu
template <int SPIN, int SRC, int LD, int REL>
__global__ void __launch_bounds__(576, 1) k_partial(unsigned* counter, int iters, long long* cyc, double* part, const double* stat, int stride) {
  unsigned gen = 0;
  const unsigned G = gridDim.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long acc_load = 0, acc_bar = 0;
  double keep = 0;
  auto barrier = [&]() {
    asm volatile("bar.sync 1, 576;" ::: "memory");
    if (threadIdx.x == 0) {
      if (REL == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      else { __threadfence(); asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory"); }
      const unsigned target = (gen + 1) * G;
      if (SPIN == 0) { while (ld_acquire(counter) < target) { __nanosleep(32); } }
      else if (SPIN == 1) { while (ld_relaxed(counter) < target) { } asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
      else if (SPIN == 2) { while (ld_relaxed(counter) < target) { } (void)ld_acquire(counter); }
      else { while (ld_acquire(counter) < target) { } }
    }
    asm volatile("bar.sync 1, 576;" ::: "memory");
    ++gen;
  };
  for (int it = 0; it < iters; ++it) {
    for (int e = threadIdx.x; e < 2016; e += blockDim.x) part[(size_t)blockIdx.x * stride + e] = it + e;
    long long tb = clock64();
    barrier();
    long long t0 = clock64();
    acc_bar += t0 - tb;
    double pv[5];
    const double* base = SRC == 0 ? part : stat;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int sl = lane + 32 * q;
      const double* a = base + (size_t)sl * stride + blockIdx.x * 7 + (w % 14);
      pv[q] = sl < (int)G ? (LD == 0 ? __ldcg(a) : ld_relaxed_d(a)) : 0.0;
    }
    double a = pv[0] + pv[1] + pv[2] + pv[3] + pv[4];
    keep += a;
    long long t1 = clock64();
    if (keep == 1.2345) printf("x");
    acc_load += t1 - t0;
    barrier();
  }
  if (threadIdx.x == 64 && blockIdx.x < 4) { cyc[2 * blockIdx.x] = acc_load; cyc[2 * blockIdx.x + 1] = acc_bar; }
  if (keep == 1.2345) printf("y");
}

template <int SPIN, int SRC, int LD, int REL>
int run(const char* name, unsigned* flag, long long* cyc, double* data, double* stat) {
  CK(cudaMemset(flag, 0, 256));
  int stride = 2016, it2 = 500;
  void* args[] = {&flag, &it2, &cyc, &data, &stat, &stride};
  CK(cudaLaunchCooperativeKernel((void*)k_partial<SPIN, SRC, LD, REL>, dim3(148), dim3(576), args, 0, 0));
  CK(cudaDeviceSynchronize());
  long long h[8];
  CK(cudaMemcpy(h, cyc, 8 * 8, cudaMemcpyDeviceToHost));
  printf("%-70s read %.0f cyc (%.2f us)  barrier %.0f cyc (%.2f us) [blk1 read %.0f]\n", name, h[0] / (double)it2, h[0] / (double)it2 / 1965.0,
         h[1] / (double)it2, h[1] / (double)it2 / 1965.0, h[2] / (double)it2);
  return 0;
}

int main() {
  long long* cyc; unsigned* flag; double *data, *stat;
  CK(cudaMalloc(&cyc, 64 * 8)); CK(cudaMalloc(&flag, 256)); CK(cudaMalloc(&data, 1 << 24)); CK(cudaMalloc(&stat, 1 << 24));
  CK(cudaMemset(data, 0, 1 << 24)); CK(cudaMemset(stat, 0, 1 << 24));
  run<0, 0, 0, 0>("acquire spin+nanosleep, fresh partials, ld.cg", flag, cyc, data, stat);
  run<0, 1, 0, 0>("acquire spin+nanosleep, static buffer, ld.cg", flag, cyc, data, stat);
  run<1, 0, 0, 0>("relaxed spin + fence.acq_rel, fresh partials, ld.cg", flag, cyc, data, stat);
  run<1, 1, 0, 0>("relaxed spin + fence.acq_rel, static buffer, ld.cg", flag, cyc, data, stat);
  run<2, 0, 0, 0>("relaxed spin + one ld.acquire, fresh partials, ld.cg", flag, cyc, data, stat);
  run<3, 0, 0, 0>("acquire spin no sleep, fresh partials, ld.cg", flag, cyc, data, stat);
  run<0, 0, 1, 0>("acquire spin+nanosleep, fresh partials, ld.relaxed.gpu", flag, cyc, data, stat);
  run<2, 0, 1, 0>("relaxed spin + one ld.acquire, fresh partials, ld.relaxed.gpu", flag, cyc, data, stat);
  run<2, 0, 0, 1>("threadfence+red.relaxed / relaxed spin + one ld.acquire, fresh, ld.cg", flag, cyc, data, stat);
  run<2, 1, 0, 1>("threadfence+red.relaxed / relaxed spin + one ld.acquire, static, ld.cg", flag, cyc, data, stat);
  return 0;
}
