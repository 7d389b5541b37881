/*
 * htlr_cuda.h — C ABI of the B200 (sm_100a) replacement for HTLR's C++ `Fit` sampler.
 *
 * What it replaces in the reference (longhaiSK/HTLR, paths relative to its root):
 *   src/main.cpp:7-25      htlr_fit_helper(): construct Fit, StartSampling(), OutputR()
 *   src/gibbs.cpp:4-530    class Fit — restricted-Gibbs / HMC sampler
 *   src/utils.cpp:3-33     log_sum_exp / find_normlv / spl_sgm_ig
 *   src/sampler.cpp:1-72   ARS targets (NEG, GHS, logw)
 *   src/ars.cpp:1-462      adaptive rejection sampling engine
 * The R side (R/core.r:198-213 -> .Call("_HTLR_htlr_fit_helper")) and the Rcpp glue keep their
 * signatures; the glue's body calls htlr_cuda_fit() (or create/run/fetch in chunks so that R stays
 * interruptible, mirroring R_CheckUserInterrupt every 256 iterations, gibbs.cpp:124).
 * INTEGRATION.md shows the binding.
 *
 * Conventions: plain C types only; every function returns 0 on success and a non-zero HTLR_E_* code
 * otherwise, never throws; htlr_cuda_last_error() gives the message. All matrices are FP64,
 * column-major (Armadillo / R layout). Inputs are borrowed for the duration of the call only and never
 * written; outputs are caller-allocated. A handle is not re-entrant; independent handles (one per GPU /
 * chain) may be used from different host threads.
 *
 * There is no CPU fallback: if no sm_100-class device is visible every entry point fails with
 * HTLR_E_CUDA.
 */
#ifndef HTLR_CUDA_H
#define HTLR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTLR_CUDA_ABI_VERSION 2

enum { HTLR_OK = 0, HTLR_E_ARG = 1, HTLR_E_CUDA = 2, HTLR_E_NCCL = 3, HTLR_E_SAMPLER = 4, HTLR_E_STATE = 5 };

/* ptype of htlr_fit_helper: "t" / "ghs" / "neg" (src/gibbs.cpp:299-309). */
enum { HTLR_PRIOR_T = 0, HTLR_PRIOR_GHS = 1, HTLR_PRIOR_NEG = 2 };

/* Where the random variates come from.
 * DEVICE: Philox4x32-10 keyed by (seed, purpose, thin-step, feature, sub-index) — reproducible,
 *         independent of the restricted-set order, identical on every row-shard rank.
 * INJECT: the host supplies them through htlr_inject in the reference's consumption order
 *         (src/gibbs.cpp:441-460 normals j-major, :90 one uniform, utils.cpp:31 p gammas). Used by the
 *         parity tests and by an Rcpp host that wants R's own RNG stream (set.seed semantics). */
enum { HTLR_RNG_DEVICE = 0, HTLR_RNG_INJECT = 1 };

/* Trajectory algorithm.
 * TWO_SWEEP: the reference's structure, one X sweep for lv and one for the gradient per leapfrog step.
 * FUSED    : persistent cooperative kernel over all SMs, one X sweep per step (gradient, both half kicks, drift
 *            and the column's lv contribution from one staged copy of the column). The fused kernels carry the
 *            linear predictor and the residual from step to step and from iteration to iteration, so they need
 *            no DetachFixlv sweep (gibbs.cpp:197-229) and no softmax before the first gradient. Needs a column to
 *            fit in shared memory (n <= 2048 for K <= 4, n <= 1024 for K = 5..8; one GPU); AUTO picks it whenever it applies.
 * FUSED_CLUSTER: the same kernel run by ONE thread-block cluster that exchanges partial sums through
 *            distributed shared memory and synchronises with hardware cluster barriers: lowest latency per
 *            leapfrog step, for restricted sets of a few hundred columns (falls back to FUSED by itself when
 *            the set is too large). AUTO launches both; the device picks one per iteration from the size of the
 *            restricted set (cluster up to ~1.5 MB of X[:, iup]).
 * Ahead of both, AUTO and CLUSTER_ONLY launch a single-block kernel that runs the trajectory
 * entirely inside one SM (two __syncthreads per leapfrog step) whenever X[:, iup] fits in its shared memory
 * and it is the faster choice (measured: n <= 256 and u <= max(4, 46000 / n^2), e.g. the first iterations of config A);
 * FUSED_NO_SMALL is AUTO without it, SMALL_FIRST uses it whenever it fits (tests).
 * For n >= 300 and K <= 4 AUTO also launches, ahead of everything, a row-partitioned cluster kernel for restricted
 * sets of a handful of columns (8 .. 64 depending on n and K, measured): every block owns a slab of rows, only the u K
 * gradient partials are exchanged per leapfrog step; ROWS_FIRST uses it whenever it fits (tests).
 * CLUSTER_ONLY: FUSED_CLUSTER without the all-SM fallback launch, so that several chains (handles) can run
 *            concurrently on one GPU, each on its own 16-SM cluster (config D: 4 chains / folds per GPU). A
 *            restricted set beyond the cluster's capacity is HTLR_E_SAMPLER. */
enum { HTLR_ALGO_AUTO = 0, HTLR_ALGO_TWO_SWEEP = 1, HTLR_ALGO_FUSED = 2, HTLR_ALGO_FUSED_CLUSTER = 3,
       HTLR_ALGO_CLUSTER_ONLY = 4, HTLR_ALGO_FUSED_NO_SMALL = 5, HTLR_ALGO_SMALL_FIRST = 6, HTLR_ALGO_ROWS_FIRST = 7 };

typedef struct htlr_handle htlr_handle;

typedef struct htlr_cfg {
  uint32_t struct_size;  /* = sizeof(htlr_cfg) */
  uint32_t abi_version;  /* = HTLR_CUDA_ABI_VERSION */
  /* --- the scalar arguments of htlr_fit_helper (src/main.cpp:7-14), same names and meaning --- */
  int32_t p, K, n;       /* n = rows held by THIS rank when row-sharded */
  int32_t ptype;
  double alpha, s, eta;
  int32_t iters_rmc, iters_h, thin, leap_L, leap_L_h;
  double leap_step;
  double hmc_sgmcut;     /* sd cut; squared inside when > 0, <= 0 updates every feature (gibbs.cpp:14) */
  int32_t keep_warmup_hist, silence, legacy;
  /* --- device-side additions --- */
  int32_t device;        /* CUDA device ordinal */
  int32_t rng_mode;      /* HTLR_RNG_* */
  int32_t algo;          /* HTLR_ALGO_* */
  uint64_t seed;         /* DEVICE rng: chain key */
  int32_t use_graph;     /* 1: replay one CUDA graph per thin-step (default), 0: eager launches */
  int32_t x_on_device;   /* 1: X/ymat/ybase pointers passed to create are device pointers. X is then BORROWED for
                            the life of the handle (not copied);
                            it must be 16-byte aligned with an even leading dimension, and for an odd n the
                            pad row n of every column must hold a finite value (the fused kernels copy whole row
                            pairs: create zeroes it when the leading dimension leaves room, i.e. ldx > n) */
  /* --- row sharding (tall X): all ranks pass the same p, K, seed; X/ymat/ybase hold local rows --- */
  void* nccl_comm;       /* ncclComm_t from htlr_cuda_nccl_init, or NULL for a single GPU */
  int32_t ars_inject_width; /* INJECT rng with ghs/neg: uniforms per (thin-step, feature) block */
  int32_t reserved;
} htlr_cfg;

/* Host-supplied variates for one htlr_cuda_run() call (INJECT mode). Consumed by sequential cursors
 * that restart at 0 on every call; running out is HTLR_E_SAMPLER. */
typedef struct htlr_inject {
  const double* normals;  int64_t n_normals;   /* momenta, u*K per thin-step, j-major (gibbs.cpp:445-452) */
  const double* uniforms; int64_t n_uniforms;  /* accept test, 1 per thin-step (gibbs.cpp:90) */
  const double* gammas;   int64_t n_gammas;    /* t prior, p per thin-step, feature order (utils.cpp:31) */
  const double* ars_uniforms; int64_t n_ars_blocks; /* ghs/neg: block (step*(p+1) + j-1) of ars_inject_width;
                                                       logw ARS: block step*(p+1) + p */
} htlr_inject;

/* Per outer iteration scalars, the values the reference prints (gibbs.cpp:119-121) and records. */
typedef struct htlr_iter_stats {
  double loglike, logw, uvar, rej;
} htlr_iter_stats;

/* Device-time summary, accumulated since create or the last reset (CUDA events on the handle's stream;
 * only collected when profiling is enabled — it forces eager launches). */
typedef struct htlr_profile {
  double sweep_ms;        /* total time of the X-sweep kernels (lv + gradient, or the fused trajectory) */
  double sweep_bytes;     /* algorithmic bytes those launches covered: 8*n*columns swept */
  int64_t sweep_launches;
  double total_ms;        /* all kernels of the timed thin-steps */
  int64_t kernel_launches;/* launches issued by this library since create / reset */
  double sigma_ms;        /* per-feature variance update (gamma draws or warp-collective ARS) */
  /* fused trajectory kernel, block 0's view (SM cycle counter / clock rate): time in the column sweeps,
   * waiting at the two grid-wide barriers of each leapfrog step, and in the row phase between them */
  double fused_sweep_ms, fused_barrier_ms, fused_rowphase_ms;
  double small_kernel_trajectories;   /* proposals run by the single-block or the row-partitioned kernel since create / reset */
  /* per-iteration timeline, collected only when the handle was created with HTLR_TIMELINE=1 in the environment
   * (graph replay stays on): microseconds from the start of the previous kernel of the iteration to the start of
   * kernel k, summed over iterations — 0 select (i.e. the tail of the previous iteration), 1 begin, 2 detach
   * sweep, 3 softmax, 4 single-block trajectory, 5 cluster-mode trajectory, 6 grid-mode trajectory, 7 energy,
   * 8 tail */
  double timeline_us[12];
  double allreduce_ms;    /* row-sharded X: total time of the NCCL all-reduces of the gradient (ABI version 2) */
} htlr_profile;

/* Replaces `Fit::Fit` + `Fit::Initialize` (gibbs.cpp:4-50, 519-530).
 * X: n x (p+1) column-major with leading dimension ldx >= n, column 0 = ones (R/core.r:144);
 * ymat: n x K one-hot of classes 2..C; ybase: n class indices in [0, C) (validated);
 * deltas0: (p+1) x K; sigmasbt0: p+1. */
int htlr_cuda_create(const htlr_cfg* cfg, const double* X, int64_t ldx, const double* ymat,
                     const int64_t* ybase, const double* deltas0, const double* sigmasbt0,
                     htlr_handle** out);
void htlr_cuda_destroy(htlr_handle* h);
/* Destroyed handles leave their device memory in a per-process cache for the next create on the same device
 * (bounded by HTLR_CACHE_MB, default 4096 MB; 0 disables). This returns it to the driver. */
void htlr_cuda_release_cache(void);

/* Replaces n_iters passes of the loop body of `Fit::StartSampling` (gibbs.cpp:57-125). Call in chunks
 * (<= 256 keeps R interruptible). stats (nullable) receives n_iters entries. Synchronises before returning. */
int htlr_cuda_run(htlr_handle* h, int32_t n_iters, const htlr_inject* inject, htlr_iter_stats* stats);

/* Replaces `Fit::OutputR` (gibbs.cpp:128-152). H = iters_rmc + 1 (+ iters_h if keep_warmup_hist).
 * mcdeltas (p+1) x K x H, mcsigmasbt / mcvardeltas (p+1) x H, the four vectors H, ddnloglike p+1.
 * Any pointer may be NULL. */
int htlr_cuda_fetch(htlr_handle* h, double* mcdeltas, double* mcsigmasbt, double* mcvardeltas,
                    double* mclogw, double* mcloglike, double* mcuvar, double* mchmcrej,
                    double* ddnloglike);
int32_t htlr_cuda_hist_len(const htlr_handle* h);

/* Pipelined form of run + fetch for hosts that want the trace download hidden behind the sampling
 * (`Fit::OutputR` copies `mcdeltas` etc. after the loop, gibbs.cpp:128-152; here the copy of slice i overlaps
 * iteration i+1..): waits until the first `upto_iters` iterations (counted since create, enqueued earlier with
 * htlr_cuda_run_async) have finished — work enqueued after them keeps running —, then copies
 *  - the per-iteration stats of iterations [stats_from, upto_iters) into `stats` (nullable), and
 *  - every trace slice that is complete and was not copied by an earlier collect call on this handle into the
 *    caller's FULL-SIZE arrays (layout of htlr_cuda_fetch; any pointer may be NULL; pass the same arrays to
 *    every call). After a collect with upto_iters = iters_h + iters_rmc the arrays equal htlr_cuda_fetch's. */
int htlr_cuda_collect(htlr_handle* h, int32_t upto_iters, int32_t stats_from, htlr_iter_stats* stats,
                      double* mcdeltas, double* mcsigmasbt, double* mcvardeltas, double* mclogw,
                      double* mcloglike, double* mcuvar, double* mchmcrej);

/* One call = htlr_fit_helper's body: create, run everything (in chunks of 256, calling `tick` after each
 * chunk with the iterations done so far and the latest stats; a non-zero return from tick aborts — the
 * Rcpp glue puts Rprintf + R_CheckUserInterrupt there), fetch, destroy. */
typedef int (*htlr_tick_fn)(void* user, int32_t iters_done, const htlr_iter_stats* chunk, int32_t n_chunk);
int htlr_cuda_fit(const htlr_cfg* cfg, const double* X, int64_t ldx, const double* ymat,
                  const int64_t* ybase, const double* deltas0, const double* sigmasbt0,
                  const htlr_inject* inject, htlr_tick_fn tick, void* user, double* mcdeltas,
                  double* mcsigmasbt, double* mcvardeltas, double* mclogw, double* mcloglike,
                  double* mcuvar, double* mchmcrej, double* ddnloglike, char* errbuf, size_t errbuf_len);

/* Posterior summaries computed from the device-resident trace, so the (p+1) x K x H cube need not be fetched:
 * mean of mcdeltas over slices first, first+thin, ... <= last (0-based; first < 0 selects htlr_sdb's default:
 * warm-up slices and a further floor(0.2*iter.rmc) dropped), `summary(fit, method = mean)` R/mccoef.r:38-62;
 * sdb[j-1] = sqrt(comp_vardeltas(mean deltas[j,])/C) for features j = 1..p, `htlr_sdb` R/mccoef.r:119-131 and
 * `comp_sdb` R/util.r:91-105; nzero[j-1] = sdb[j-1] > cut * max(sdb), `nzero_idx` R/mccoef.r:151-155.
 * mean_deltas (p+1) x K, sdb p, nzero p; any may be NULL. */
int htlr_cuda_summary(htlr_handle* h, int32_t first, int32_t last, int32_t thin, double cut, double* mean_deltas,
                      double* sdb, int32_t* nzero);

/* Predictive class probabilities of n_ts new cases, averaged over trace slices first, first+thin, ... <= last
 * (0-based; first < 0 selects htlr_predict's default range), computed from the device-resident trace:
 * the arithmetic of `htlr_predict` (R/predict.R:28-92). X_ts: n_ts x (p+1) column-major with leading dimension
 * ldx, already feature-selected / standardised / with the intercept column, exactly like X at create.
 * probs: n_ts x (K+1) column-major, column 0 = pmax(0, 1 - sum of the others). */
int htlr_cuda_predict(htlr_handle* h, const double* X_ts, int64_t ldx, int32_t n_ts, int32_t first, int32_t last,
                      int32_t thin, double* probs);

/* `std_helper` (src/utils.cpp:36-48) on the GPU: column medians, column sample standard deviations (n - 1) and
 * A_std = (A - median) / sd for column-major n x p data. on_device = 1: A and A_std are device pointers (A_std may
 * equal A or be NULL); median / sd (length p, host) may be NULL. Exact medians (radix selection). */
int htlr_cuda_std(int32_t device, int32_t n, int32_t p, const double* A, int64_t lda, int32_t on_device,
                  double* A_std, int64_t ldo, double* median, double* sd);

/* ---- deterministic pieces, for the parity gate (level 1 of north_star) ---- */

/* lv = X[:,iup] * deltas[iup,:] (nothing fixed), row-softmax, log-likelihood and gradient over iup:
 * UpdatePredProb + UpdateLogLike + UpdateDNlogLike (gibbs.cpp:176-191, 236-261). Chain state untouched.
 * lv, pred: n x (K+1); grad: u x K (row q <-> iup[q]). Any output may be NULL. */
int htlr_cuda_eval(htlr_handle* h, const int32_t* iup, int32_t u, const double* deltas, double* lv,
                   double* pred, double* loglike, double* grad);

/* One HMC proposal from the current chain state with the given momenta ((p+1) x K, rows outside the
 * restricted set ignored) and L leapfrog steps — gibbs.cpp:67-87 without the accept step; chain state
 * untouched. Outputs: proposal deltas and momenta (p+1) x K, lv n x (K+1), var_deltas p+1, log-likelihood,
 * the two negative energies of gibbs.cpp:76 and :87, and the restricted-set size. */
int htlr_cuda_trajectory(htlr_handle* h, const double* momt, int32_t L, double* deltas_out,
                         double* momt_out, double* lv_out, double* loglike_out, double* var_deltas_out,
                         double* nenergy_old, double* nenergy_new, int32_t* nuvar);

/* Current chain state (any pointer may be NULL): lv n x (K+1), deltas (p+1) x K, sigmasbt, var_deltas p+1. */
int htlr_cuda_get_state(htlr_handle* h, double* lv, double* deltas, double* sigmasbt, double* var_deltas,
                        double* loglike, double* logw);

/* Warp-per-feature adaptive rejection sampling of x = log sigma^2_j, one draw per feature, standalone
 * (UpdateSigmasSgm, gibbs.cpp:351-372 + sampler.cpp:18-49 + ars.cpp). kind: HTLR_PRIOR_GHS / _NEG.
 * uniforms: m x width host array (feature-major) or NULL to use the device Philox stream (seed, step).
 * n_unif / n_eval (nullable): uniforms consumed / target evaluations per feature. */
int htlr_cuda_ars(int32_t device, int32_t kind, int32_t m, const double* var_deltas, double K, double alpha,
                  double log_aw, const double* uniforms, int32_t width, uint64_t seed, uint64_t step,
                  double* x_out, int32_t* n_unif, int32_t* n_eval, int32_t* status);

/* Device Philox variates, for cross-checking against the oracle's twin: kind 0 uniform, 1 normal,
 * 2 gamma(shape). out[a*n_b + b] for a < n_a, b < n_b (gamma: a = feature, b ignored -> n_b must be 1). */
int htlr_cuda_rng_dump(int32_t device, uint64_t seed, uint32_t purpose, uint64_t step, int32_t n_a,
                       int32_t n_b, int32_t kind, double shape, double* out);

/* ---- measurement ---- */
void* htlr_cuda_stream(htlr_handle* h);                 /* cudaStream_t all work is issued on */
int htlr_cuda_profile_enable(htlr_handle* h, int32_t on); /* on: eager launches + per-kernel CUDA events */
int htlr_cuda_profile_get(htlr_handle* h, htlr_profile* out, int32_t reset);
int64_t htlr_cuda_launch_count(const htlr_handle* h);   /* kernels launched by this handle so far */
/* Device time (CUDA events around the kernels, copies excluded) and FP64 flops (2 n_ts (p+1) K S) of the last
 * htlr_cuda_predict call of this handle. */
int htlr_cuda_predict_stats(const htlr_handle* h, double* kernel_ms, double* flops);
/* Measured FP64 peaks of a device in TFLOP/s: independent DFMA chains (vector pipe) and independent m8n8k4 DMMA chains
 * (the path htlr_cuda_predict's contraction runs on); the roofline denominators of the prediction block of bench.py. */
int htlr_cuda_fp64_peak(int32_t device, double* dfma_tflops, double* dmma_tflops);
/* Adaptive rejection sampling (ghs / neg variances, logw): the reference's hull table holds max_nhull = 1000 tangents
 * and silently stops tightening the envelope once it is full (src/ars.cpp:26, src/ars.h:59-62); the device engine does
 * the same with 16 tangents per feature (32 for logw). The draws stay exact either way. This returns how many draws of
 * this handle filled their table (waits for the enqueued iterations; -1 on error): 0 on every workload seen so far -
 * the reference's targets need at most 13. The step-out phase running out of tangents is an error, as in the
 * reference (ars.cpp:155-160, 187-193). */
int64_t htlr_cuda_ars_table_full(htlr_handle* h);
int htlr_cuda_sync(htlr_handle* h);
/* enqueue n_iters iterations without synchronising or copying stats (bench timed region) */
int htlr_cuda_run_async(htlr_handle* h, int32_t n_iters);

/* ---- synthetic data on the device: the laws of the reference's generators, written straight into the
 * layout htlr_cuda_create takes with x_on_device. All array arguments but muj / Ablock / deltas_true are DEVICE
 * pointers: X n x (p+1) column-major with leading dimension ldx (column 0 = ones), ymat n x (NC-1) column-major
 * (leading dimension n), ybase n. Variates are keyed by (seed, global row row0 + i, column): any row block of one
 * data set can be generated on any GPU. 2 <= NC <= 17.
 * gendata_MLR (R/gendata.r:26-46): X ~ N(0,1), sigmasbt = w / Gamma(nu/2, rate nu/2), betas = N(0,1) * c(0, sqrt(sigmasbt)),
 *   y ~ categorical(softmax(cbind(1, X) betas)); deltas_true (host, (p+1) x (NC-1), nullable) = betas[, -1] - betas[, 1].
 * gendata_FAM (R/gendata.r:105-113, src/utils.cpp:72-105): y = rep(1:NC, len = n), X = A z + muj[, y] + sd_g e with
 *   A = I_p whose leading pb x kb block is Ablock (host, pb x kb column-major; the vignette's factor pattern), muj host
 *   p x NC column-major; stdx != 0 standardises with the model's means and variances as the reference does. */
int htlr_cuda_gendata_mlr(int32_t device, int32_t n, int32_t p, int32_t NC, double nu, double w, uint64_t seed,
                          int64_t row0, double* X, int64_t ldx, double* ymat, int64_t* ybase, double* deltas_true);
int htlr_cuda_gendata_fam(int32_t device, int32_t n, int32_t p, int32_t NC, const double* muj, int32_t pb, int32_t kb,
                          const double* Ablock, double sd_g, int32_t stdx, uint64_t seed, int64_t row0, double* X,
                          int64_t ldx, double* ymat, int64_t* ybase);

/* ---- multi-GPU plumbing (row-sharded X) ---- */
int htlr_cuda_nccl_unique_id(void* id128 /* 128 bytes */);
int htlr_cuda_nccl_init(const void* id128, int32_t rank, int32_t world, int32_t device, void** comm);
void htlr_cuda_nccl_destroy(void* comm);

const char* htlr_cuda_last_error(const htlr_handle* h); /* h == NULL: last error of a failed create / static call */
int32_t htlr_cuda_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HTLR_CUDA_H */
