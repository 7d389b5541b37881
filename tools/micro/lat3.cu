// Grid barrier variants, timed as whole-kernel time / iterations (148 blocks x 512 threads, cooperative).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// MODE 0: counter, acquire poll + nanosleep(32) by thread 0
// MODE 1: counter, acquire poll, no sleep
// MODE 2: flag per block (128 B apart): thread 0 st.release own flag; warp 0 polls all flags (relaxed) then fence
// MODE 3: flag per block; every warp's lane set polls? no: warp 0 polls with ld.acquire
// MODE 4: counter, relaxed poll + fence.acq_rel after
// STORES: bytes of partial stores per block before the barrier (0 or 16384)
template <int MODE, int STORES>
__global__ void __launch_bounds__(512, 1) k_gbar(unsigned* counter, unsigned* flags, int iters, double* buf) {
  unsigned gen = 0;
  const unsigned G = gridDim.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int it = 0; it < iters; ++it) {
    if (STORES) {
      for (int e = threadIdx.x; e < STORES / 8; e += 512) buf[(size_t)blockIdx.x * 4096 + e] = it + e;
    }
    ++gen;
    if (MODE == 0 || MODE == 1 || MODE == 4) {
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        const unsigned target = gen * G;
        if (MODE == 0) { while (ld_acquire(counter) < target) { __nanosleep(32); } }
        else if (MODE == 1) { while (ld_acquire(counter) < target) { } }
        else { while (ld_relaxed(counter) < target) { } asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
    } else {
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (w == 0) {
        if (lane == 0) st_release(flags + 32 * blockIdx.x, gen);
        bool done;
        do {
          done = true;
          for (unsigned q = lane; q < G; q += 32) {
            unsigned v = MODE == 2 ? ld_relaxed(flags + 32 * q) : ld_acquire(flags + 32 * q);
            done = done && (v >= gen);
          }
          done = __all_sync(0xffffffffu, done);
        } while (!done);
        if (MODE == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
  }
}
template <int MODE, int STORES>
int run(const char* name, unsigned* counter, unsigned* flags, double* buf) {
  CK(cudaMemset(counter, 0, 256)); CK(cudaMemset(flags, 0, 148 * 128));
  int iters = 2000;
  void* args[] = {&counter, &flags, &iters, &buf};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaLaunchCooperativeKernel((void*)k_gbar<MODE, STORES>, dim3(148), dim3(512), args, 0, 0));
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(counter, 0, 256)); CK(cudaMemset(flags, 0, 148 * 128));
  cudaEventRecord(e0);
  CK(cudaLaunchCooperativeKernel((void*)k_gbar<MODE, STORES>, dim3(148), dim3(512), args, 0, 0));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-64s %.3f us per barrier\n", name, ms * 1e3 / iters);
  return 0;
}
int main() {
  unsigned *counter, *flags; double* buf;
  CK(cudaMalloc(&counter, 256)); CK(cudaMalloc(&flags, 148 * 128)); CK(cudaMalloc(&buf, 148 * 4096 * 8));
  run<0, 0>("counter, acquire poll + nanosleep(32), no stores", counter, flags, buf);
  run<1, 0>("counter, acquire poll, no sleep, no stores", counter, flags, buf);
  run<4, 0>("counter, relaxed poll + fence, no stores", counter, flags, buf);
  run<2, 0>("flag per block, relaxed poll + fence, no stores", counter, flags, buf);
  run<3, 0>("flag per block, acquire poll, no stores", counter, flags, buf);
  run<0, 16384>("counter, acquire poll + nanosleep(32), 16 KB stores", counter, flags, buf);
  run<1, 16384>("counter, acquire poll, no sleep, 16 KB stores", counter, flags, buf);
  run<4, 16384>("counter, relaxed poll + fence, 16 KB stores", counter, flags, buf);
  run<2, 16384>("flag per block, relaxed poll + fence, 16 KB stores", counter, flags, buf);
  run<3, 16384>("flag per block, acquire poll, 16 KB stores", counter, flags, buf);
  return 0;
}
