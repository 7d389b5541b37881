// Shared device-side definitions for the HTLR sampler kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace htlr {

constexpr int kMaxK = 16;  // classes - 1 the generic kernels hold in registers at a time (more: kMaxK per pass over X)

// Error codes stored in Ctl::err (sticky, first one wins).
enum DevErr : int {
  kErrNone = 0,
  kErrInjectNormals = 1,
  kErrInjectUniforms = 2,
  kErrInjectGammas = 3,
  kErrInjectArs = 4,
  kErrClusterCap = 5,   // CLUSTER_ONLY: restricted set larger than the cluster keeps state for
  kErrArs = 16,  // + ArsStatus in err_detail
};

// Device-resident control block: everything that varies from thin-step to thin-step lives here so a
// single captured CUDA graph can be replayed for every iteration with no host round trip.
struct Ctl {
  // static configuration
  int nvar, K, C, n, p;
  int iters_h, iters_rmc, thin, keep_warmup_hist, ptype, rng_mode;
  double alpha, s, eta, leap_step, sgmsq_cut;
  unsigned long long seed;
  // progress
  int i_mc, i_thin;
  unsigned long long step;       // global thin-step index (Philox key)
  long long step_in_call;        // thin-steps since the current run() call (inject block index)
  // restricted set
  int u, nfix, init_all;
  // chain scalars
  double logw, loglike, loglike_new, loglike_old;
  double nenergy_old, nenergy_new;
  int accept;
  unsigned long long small_runs;   // trajectories run by the single-block kernel (profiling / tests)
  int traj_done;   // cluster-mode trajectory kernel ran this thin-step (the grid-mode launch then exits)
  double uvar_acc, rej_acc;
  // inject cursors and sizes
  long long cur_norm, cur_unif, cur_gam, n_norm, n_unif, n_gam, n_ars_blocks;
  int ars_width;
  // errors
  int err, err_detail;
  unsigned long long ars_full;   // ARS draws (ghs / neg variances, logw) whose hull table filled up (ars.cuh, insert())
  // X columns swept so far (lv + gradient sweeps), for the roofline arithmetic
  unsigned long long cols_swept;
  // fused trajectory kernel, block 0: SM cycles spent in the column sweep, the two grid barriers and the
  // row phase (partial reduction + softmax), accumulated over launches (htlr_profile.fused_*_ms)
  unsigned long long ph_cycles[4];
  // per-iteration timeline (HTLR_TIMELINE=1): tl_acc[k] accumulates the %globaltimer nanoseconds between the
  // start of the previous marked kernel of the iteration and the start of kernel k (0 select, 1 begin,
  // 2 detach sweep, 3 softmax, 4 single-block trajectory, 5 cluster-mode, 6 grid-mode, 7 energy, 8 tail)
  int timeline;
  unsigned long long tl_prev, tl_acc[12];
  // scratch counters for "last block finishes" reductions
  unsigned int ticket[8];
  // k_select over several blocks: per-block counts chained through (epoch << 32 | count) words
  unsigned long long sel_flag[32];
  unsigned int sel_epoch;
};

// Start-of-kernel time stamp for the per-iteration timeline; the marked kernels of an iteration run one after
// the other, so the read-modify-write needs no atomics.
__device__ __forceinline__ void timeline_mark(Ctl* ctl, int k) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && ctl->timeline) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    ctl->tl_acc[k] += t - ctl->tl_prev;
    ctl->tl_prev = t;
  }
}

// End-of-kernel stamp, called by the one thread that finishes the kernel's last-block epilogue
// (9 tail, 10 energy, 11 begin): separates a kernel's own time from the gap to the next kernel's start.
__device__ __forceinline__ void timeline_mark_end(Ctl* ctl, int k) {
  if (ctl->timeline) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    ctl->tl_acc[k] += t - ctl->tl_prev;
    ctl->tl_prev = t;
  }
}

struct Red4 {
  double a, b, c, d;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block in a fixed order (lanes by xor-tree, then warps left to right). Result valid in
// every thread. `sm` must hold 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double t = 0;
  for (int i = 0; i < nw; ++i) t += sm[i];
  return t;
}

__device__ __forceinline__ void set_err(Ctl* ctl, int code, int detail) {
  if (atomicCAS(&ctl->err, 0, code) == 0) ctl->err_detail = detail;
}

// Grid-wide "last block done" election. Returns true in exactly one block (all its threads), after the
// writes of every other block are visible.
__device__ __forceinline__ bool last_block(unsigned int* ticket, unsigned int nblocks) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == nblocks - 1);
    if (s_last) *ticket = 0;  // re-arm for the next launch
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

}  // namespace htlr
