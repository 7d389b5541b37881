// Host-callable launch wrappers (defined in the kernel translation units, used by api.cu).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdint>

#include "common.cuh"

namespace htlr {

// cudaFuncSetAttribute applies to the CURRENT device only, and a process may drive several GPUs (one host thread
// per chain / fold): one "configured" bit per device for every kernel instantiation that needs opt-in shared memory.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  static unsigned long long bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
  }
  bool needed() const { return (done.load(std::memory_order_acquire) & bit()) == 0; }
  void mark() { done.fetch_or(bit(), std::memory_order_acq_rel); }
};

// Geometry of the sweep kernels, fixed at create time (grids never depend on the restricted-set size,
// which only the device knows; surplus blocks exit immediately).
struct Geom {
  int n, nvar, K;
  int64_t ld;       // leading dimension of X on the device (doubles), even
  int n_s;          // row stride of n-length arrays (lv, R, ymat), even
  // lv sweep: 256 threads x 2 rows per block
  int lv_rt;        // row tiles
  int lv_cs;        // max column splits (partial-sum slots)
  int lv_grid;      // blocks: one full wave (a block walks (row tile, split) items with the grid's stride)
  // gradient sweep: 8 warps per block, one column per warp at a time
  int g_tr;         // rows per tile (a block: 8 warps x PP*64 rows)
  int g_rt;         // row tiles
  int g_gx;         // column groups per row tile
  int g_grid;       // blocks: one full wave (a block walks (row tile, column group) items with the grid's stride)
  // elementwise kernels over the restricted set
  int e_grid;
  int sm_count;
};

// Shape of the fused trajectory kernel (traj_kernel.cu), chosen once per handle by fused_plan().
struct FusedGeom {
  int T;       // columns per ring stage (1)
  int S;       // column stride in shared memory (doubles), even
  int NS;      // ring stages per compute group
  int nrp;     // row pairs per thread within a column (template parameter: 1, 2 or 4)
  int gw;      // warps per compute group (template parameter: 1, 4, 8 or 16)
  int nlp;     // row pairs per thread when publishing the block's lv partial
  int rpc;     // rows per block in the softmax phase
  int grid;    // blocks (= SMs, or the cluster size)
  int cl;      // 0: grid mode (cooperative launch); > 0: one thread-block cluster of this many blocks
  int ncmax;   // most restricted-set columns one block can own (state kept in shared memory)
  int cl_umax; // cluster mode runs the trajectory only while the restricted set has at most this many columns
  int cl_only; // no grid-mode launch follows: exceeding cl_umax is an error
  int timing;  // 1: block 0 accumulates per-phase cycle counts into Ctl::ph_cycles (profiling runs only)
  size_t smem_bytes, smem_bytes_max;
};

struct Bufs {
  // data
  const double* X;
  const double* ymat;   // [k][i], stride n_s
  const int* ybase;
  const double* ddn;    // colsum(X^2)/4
  // chain state
  double *deltas, *sigmasbt, *var_deltas;   // nvar x K col-major, nvar, nvar
  double *lv, *lv_old, *lv_fix, *R, *R_old;  // [k][i], stride n_s (class columns 1..K; class 0 is 0)
  // restricted set + compact per-feature work arrays (index r = position in iup)
  int *iup, *ifix;
  int* posof;           // inverse map: position of feature j in iup, or -1 (written by k_select)
  double *dc, *momt, *stepc, *sigc, *vdc;   // u x K, u x K, u, u, u
  double* g;            // gradient, u x K (+1 trailing slot for the local log-likelihood when sharded)
  double* gpart;        // [g_rt][nvar*K] partials when g_rt > 1
  double* lvpart;       // [lv_cs][K][n_s]
  Red4* red;            // per-block partial sums
  Red4* red_begin;      // k_begin's per-block energy partials (.a, .b; [0].c = number of blocks), read by k_energy
  Ctl* ctl;
  // history
  double *mcdeltas, *mcsigmasbt, *mcvardeltas, *mclogw, *mcloglike, *mcuvar, *mchmcrej;
  htlr::Red4* stats;    // per outer iteration: loglike, logw, uvar, rej
  // inject
  const double *inj_norm, *inj_unif, *inj_gam, *inj_ars;
  // optional user momenta for htlr_cuda_trajectory (nvar x K col-major) or null
  const double* user_momt;
};

void launch_colsumsq(const Geom& g, const double* X, double* ddn, cudaStream_t st);
void launch_select(const Geom& g, const Bufs& b, cudaStream_t st);
void launch_begin(const Geom& g, const Bufs& b, cudaStream_t st);
// mode 0: detach sweep (iup/compact coefficients if u <= nvar/2 else ifix/global deltas)
// mode 1: trajectory sweep over iup with compact coefficients
// mode 2: explicit index list + global coefficient matrix (eval hook, Initialize)
void launch_lv_sweep(const Geom& g, const Bufs& b, int mode, const int* idx, const int* cnt,
                     const double* coef_global, cudaStream_t st);
// what 0: detach finish (lv_old = lv, lv_fix, R from current lv)
// what 1: trajectory softmax (lv = lv_fix + sum of partials, R, log-likelihood -> ctl.loglike_new)
// what 2: like 1 but lv_fix taken as zero and also writes pred (n x C col-major) if non-null
void launch_softmax(const Geom& g, const Bufs& b, int what, const int* cnt, double* lv_out, double* R_out,
                    double* pred_out, double* loglike_out, cudaStream_t st);
void launch_grad_sweep(const Geom& g, const Bufs& b, const int* idx, const int* cnt, const double* R,
                       double* gout, cudaStream_t st);
// sums row-tile partials into g (only needed before an all-reduce; k_post does it itself otherwise)
int lv_sweep_blocks_per_sm(int K);     // resident blocks per SM (occupancy query; current device)
int grad_sweep_blocks_per_sm(int K);
int grad_sweep_tile_rows(int K);       // rows one block of the gradient sweep covers
void launch_gsum(const Geom& g, const Bufs& b, const int* cnt, double* gout, cudaStream_t st);
void launch_post(const Geom& g, const Bufs& b, int s, int L, int g_is_summed, cudaStream_t st);
void launch_energy(const Geom& g, const Bufs& b, int g_has_loglike, cudaStream_t st);
void launch_commit(const Geom& g, const Bufs& b, cudaStream_t st);
void launch_record(const Geom& g, const Bufs& b, cudaStream_t st);
void launch_vardeltas_all(const Geom& g, const Bufs& b, cudaStream_t st);
void launch_scatter_proposal(const Geom& g, const Bufs& b, double* deltas_out, double* momt_out,
                             double* vd_out, cudaStream_t st);

// posterior summaries from the device-resident trace (mean over slices first..last step thin, sdb, flags)
void launch_trace_summary(const Geom& g, const Bufs& b, int first, int last, int thin, double cut, double* mean_out,
                          double* sdb_out, int* flags_out, cudaStream_t st);

// predict_kernels.cu: averaged predictive probabilities of M cases from S trace slices (first, first+thin, ...);
// LV is scratch for M x K x chunk doubles, accum for M x K; returns the number of kernels launched
int launch_predict(const Geom& g, const Bufs& b, int M, const double* A, int64_t lda, int first, int thin, int S,
                   double* LV, int chunk, double* accum, double* probs, cudaStream_t st);

// gen_kernels.cu: gendata_MLR / gendata_FAM laws on the device (betas: (p+1) x NC scratch/output, part: chunks x NC x n)
void launch_gen_mlr(int n, int p, int NC, double nu, double w, unsigned long long seed, long long row0, double* X,
                    int64_t ld, double* ymat, long long* ybase, double* betas, double* part, cudaStream_t st);
int gen_mlr_chunks(int p);
void launch_gen_fam(int n, int p, int NC, int pb, int kb, const double* muj, const double* Ab, double sd_g, int stdx,
                    unsigned long long seed, long long row0, double* X, int64_t ld, double* ymat, long long* ybase,
                    cudaStream_t st);
void launch_fp64_peak(int sm_count, int iters, int use_dmma, double* sink, cudaStream_t st);

// traj_kernel.cu
bool fused_plan(const Geom& g, int smem_optin_bytes, int cl, FusedGeom* out);
int fused_cluster_size(const Geom& g, int smem_optin_bytes);
cudaError_t launch_traj_fused(const Geom& g, const Bufs& b, const FusedGeom& fg, int L, cudaStream_t st);

// std_kernels.cu: std_helper (src/utils.cpp:36-48) on device-resident column-major data; out may be null or == A
void launch_std(int n, int p, const double* A, int64_t lda, double* out, int64_t ldo, double* med, double* sd,
                int sm_count, cudaStream_t st);

// small_kernel.cu: single-block trajectory for restricted sets that fit in one SM's shared memory
bool small_applies(const Geom& g);
cudaError_t launch_traj_small(const Geom& g, const Bufs& b, int L, int smem_bytes, int u_max, cudaStream_t st);

// rows_kernel.cu: row-partitioned cluster kernel for restricted sets with few columns (only the gradient is exchanged)
int rows_cluster_size(const Geom& g, int smem_bytes);  // cluster size it can run with on the current device, or 0
cudaError_t launch_traj_rows(const Geom& g, const Bufs& b, int L, int cl, int uk_max, int smem_bytes, cudaStream_t st);

// sigma_kernels.cu (compiled with --fmad=false)
void launch_sigma(const Geom& g, const Bufs& b, cudaStream_t st);      // t prior: gamma draws
void launch_tail_t(const Geom& g, const Bufs& b, cudaStream_t st);     // t prior, eta = 0: commit + sigma + record in one
void launch_sigma_ars(const Geom& g, const Bufs& b, cudaStream_t st);  // ghs / neg: warp-per-feature ARS
void launch_logw(const Geom& g, const Bufs& b, cudaStream_t st);
void launch_ars_standalone(int kind, int m, const double* vard, double K, double alpha, double log_aw,
                           const double* uniforms, int width, unsigned long long seed,
                           unsigned long long step, double* x_out, int* n_unif, int* n_eval, int* status,
                           cudaStream_t st);
void launch_rng_dump(unsigned long long seed, unsigned int purpose, unsigned long long step, int n_a, int n_b,
                     int kind, double shape, double* out, cudaStream_t st);

}  // namespace htlr
