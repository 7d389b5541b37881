// Latency microbenchmarks for the row phase / grid barrier design (B200). nvcc -arch=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, b, a);
  }
  long long t1 = clock64();
  out[0] = x; cyc[0] = t1 - t0;
  double y = a;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) y = y + b;
  }
  t1 = clock64();
  out[1] = y; cyc[1] = t1 - t0;
  double z = a;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) z = exp(z * 1e-3 - 0.5);
  t1 = clock64();
  out[2] = z; cyc[2] = t1 - t0;
  z = a + 2.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) z = log(z + 3.0);
  t1 = clock64();
  out[3] = z; cyc[3] = t1 - t0;
  z = a + 2.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) z = 1.0 / (z + 3.0);
  t1 = clock64();
  out[4] = z; cyc[4] = t1 - t0;
  z = a + 2.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) z = sqrt(z + 3.0);
  t1 = clock64();
  out[5] = z; cyc[5] = t1 - t0;
  // shuffle + dadd chain
  z = a;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) z += __shfl_xor_sync(0xffffffffu, z, 1);
  t1 = clock64();
  out[6] = z; cyc[6] = t1 - t0;
}

// pointer chase through an L2-resident buffer with ld.cg
__global__ void k_chase(const unsigned* next, int iters, long long* cyc, unsigned* sink) {
  unsigned p = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) p = __ldcg(next + p);
  long long t1 = clock64();
  cyc[0] = t1 - t0; sink[0] = p;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

__global__ void k_fence(unsigned* flag, double* data, long long* cyc) {
  // single thread: cost of individual sync primitives
  long long t0 = clock64();
  for (int i = 0; i < 100; ++i) (void)ld_acquire(flag);
  long long t1 = clock64(); cyc[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < 100; ++i) (void)ld_relaxed(flag);
  t1 = clock64(); cyc[1] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < 100; ++i) __threadfence();
  t1 = clock64(); cyc[2] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < 100; ++i) { data[i] = i; asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory"); }
  t1 = clock64(); cyc[3] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < 100; ++i) { data[i] = i; asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory"); }
  t1 = clock64(); cyc[4] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < 100; ++i) { atomicAdd(flag, 1u); }
  t1 = clock64(); cyc[5] = t1 - t0;
  // load after acquire: does the acquire slow down the next plain L2 load?
  double s = 0;
  t0 = clock64();
  for (int i = 0; i < 100; ++i) { (void)ld_acquire(flag); s += __ldcg(data + 32 * i); }
  t1 = clock64(); cyc[6] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < 100; ++i) { s += __ldcg(data + 32 * i + 4096); }
  t1 = clock64(); cyc[7] = t1 - t0;
  data[0] = s;
}

// grid barrier variants, 148 blocks x 576 threads
template <int MODE>
__global__ void __launch_bounds__(576, 1) k_gbar(unsigned* counter, int iters, long long* cyc, double* buf) {
  unsigned gen = 0;
  const unsigned G = gridDim.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) buf[(size_t)blockIdx.x * 1024 + threadIdx.x] = it;   // stores in flight before the barrier
    asm volatile("bar.sync 1, 576;" ::: "memory");
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      const unsigned target = (gen + 1) * G;
      if (MODE == 0 || MODE == 2) { while (ld_acquire(counter) < target) { __nanosleep(32); } }
      else { while (ld_relaxed(counter) < target) { } __threadfence(); }
    }
    asm volatile("bar.sync 1, 576;" ::: "memory");
    ++gen;
  }
  long long t1 = clock64();
  if (threadIdx.x == 64 && blockIdx.x < 4) cyc[blockIdx.x] = t1 - t0;
}

// reading partials written by other blocks right after a barrier
__global__ void __launch_bounds__(576, 1) k_partial(unsigned* counter, int iters, long long* cyc, double* part, int stride) {
  unsigned gen = 0;
  const unsigned G = gridDim.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long acc_load = 0;
  double keep = 0;
  for (int it = 0; it < iters; ++it) {
    // publish
    for (int e = threadIdx.x; e < 2016; e += blockDim.x) part[(size_t)blockIdx.x * stride + e] = it + e;
    asm volatile("bar.sync 1, 576;" ::: "memory");
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      const unsigned target = (gen + 1) * G;
      while (ld_acquire(counter) < target) { __nanosleep(32); }
    }
    asm volatile("bar.sync 1, 576;" ::: "memory");
    ++gen;
    long long t0 = clock64();
    double pv[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int sl = lane + 32 * q;
      pv[q] = sl < (int)G ? __ldcg(part + (size_t)sl * stride + blockIdx.x * 7 + (w % 14)) : 0.0;
    }
    double a = pv[0] + pv[1] + pv[2] + pv[3] + pv[4];
    keep += a;
    long long t1 = clock64();
    if (keep == 1.2345) printf("x");
    acc_load += t1 - t0;
    // second barrier so writes of the next round do not race the reads
    asm volatile("bar.sync 1, 576;" ::: "memory");
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      const unsigned target = (gen + 1) * G;
      while (ld_acquire(counter) < target) { __nanosleep(32); }
    }
    asm volatile("bar.sync 1, 576;" ::: "memory");
    ++gen;
  }
  if (threadIdx.x == 64 && blockIdx.x < 4) cyc[blockIdx.x] = acc_load;
  if (keep == 1.2345) printf("y");
}

int main() {
  double* out; long long* cyc; unsigned* flag; double* data;
  CK(cudaMalloc(&out, 64 * 8)); CK(cudaMalloc(&cyc, 64 * 8)); CK(cudaMalloc(&flag, 256)); CK(cudaMalloc(&data, 1 << 24));
  CK(cudaMemset(flag, 0, 256)); CK(cudaMemset(data, 0, 1 << 24));
  long long h[16];
  k_dfma<<<1, 32>>>(out, cyc, 0.5, 0.999);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, cyc, 8 * 8, cudaMemcpyDeviceToHost));
  printf("dependent DFMA %.1f cyc, DADD %.1f cyc, exp %.0f cyc, log %.0f cyc, div %.0f cyc, sqrt %.0f cyc, shfl+dadd %.0f cyc\n",
         h[0] / 1024.0, h[1] / 1024.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0, h[5] / 256.0, h[6] / 256.0);
  // pointer chase: 8 MB buffer, stride pattern
  {
    const int N = 1 << 21; unsigned* hn = new unsigned[N];
    for (int i = 0; i < N; ++i) hn[i] = (unsigned)(((long long)i * 40503 + 977) % N);
    unsigned* dn; CK(cudaMalloc(&dn, N * 4)); CK(cudaMemcpy(dn, hn, N * 4, cudaMemcpyHostToDevice));
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    k_chase<<<1, 1>>>(dn, 2000, cyc, sink); CK(cudaDeviceSynchronize());
    k_chase<<<1, 1>>>(dn, 2000, cyc, sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("ld.cg chase (L2 resident, 8MB) %.0f cyc/load\n", h[0] / 2000.0);
  }
  k_fence<<<1, 1>>>(flag, data, cyc); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, cyc, 8 * 8, cudaMemcpyDeviceToHost));
  printf("ld.acquire.gpu %.0f, ld.relaxed.gpu %.0f, threadfence %.0f, st+red.release %.0f, st+red.relaxed %.0f, atomicAdd %.0f, acquire+ldcg %.0f, ldcg %.0f cyc\n",
         h[0] / 100.0, h[1] / 100.0, h[2] / 100.0, h[3] / 100.0, h[4] / 100.0, h[5] / 100.0, h[6] / 100.0, h[7] / 100.0);
  int iters = 1000;
  for (int mode = 0; mode < 3; ++mode) {
    CK(cudaMemset(flag, 0, 256));
    void* args[] = {&flag, &iters, &cyc, &data};
    void* fn = mode == 0 ? (void*)k_gbar<0> : mode == 1 ? (void*)k_gbar<1> : (void*)k_gbar<2>;
    CK(cudaLaunchCooperativeKernel(fn, dim3(148), dim3(576), args, 0, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, 8 * 4, cudaMemcpyDeviceToHost));
    printf("grid barrier mode %d: %.0f cyc per barrier (%.2f us)\n", mode, h[0] / (double)iters, h[0] / (double)iters / 1965.0);
  }
  {
    CK(cudaMemset(flag, 0, 256));
    int stride = 2016, it2 = 500;
    void* args[] = {&flag, &it2, &cyc, &data, &stride};
    CK(cudaLaunchCooperativeKernel((void*)k_partial, dim3(148), dim3(576), args, 0, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, 8 * 4, cudaMemcpyDeviceToHost));
    printf("post-barrier read of 148 partials (5 independent ld.cg per lane): %.0f cyc (%.2f us) blk0, %.0f blk1\n",
           h[0] / (double)it2, h[0] / (double)it2 / 1965.0, h[1] / (double)it2);
  }
  return 0;
}
