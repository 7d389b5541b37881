// C ABI of the CUDA sampler (include/htlr_cuda.h): handle life-cycle, the per-thin-step launch
// sequence (eager or one captured CUDA graph per trajectory length), result fetch, parity hooks and
// the NCCL plumbing for row-sharded X. No compute happens on the host and there is no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: the library itself is bound at run time (see NcclApi below)

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/htlr_cuda.h"
#include "common.cuh"
#include "launch.h"

using namespace htlr;

namespace {

thread_local std::string g_err;

struct CudaFail {
  int code;
  std::string msg;
};

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      throw CudaFail{HTLR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)};             \
  } while (0)
#define NC(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess)                                                                         \
      throw CudaFail{HTLR_E_NCCL, std::string(#call) + ": " + nccl_api().GetErrorString(r_)};             \
  } while (0)
#define FAIL(code, text) throw CudaFail{code, text}

int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// The few device attributes the library needs, read once per device with cudaDeviceGetAttribute:
// cudaGetDeviceProperties costs 2.5 ms per call on this platform and now and then 15-180 ms, which showed up as
// outliers of a whole fit's end-to-end time.
struct DevInfo {
  int major = 0, sm_count = 0, smem_optin = 0, coop = 0;
  bool ok = false;
};
const DevInfo& dev_info(int device) {
  static std::mutex mu;
  static DevInfo table[64];
  std::lock_guard<std::mutex> lk(mu);
  DevInfo& d = table[device & 63];
  if (!d.ok) {
    CU(cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, device));
    CU(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
    CU(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CU(cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, device));
    d.ok = true;
  }
  return d;
}

// HTLR_TRACE_CREATE=1: wall-clock stamps of the phases of htlr_cuda_create on stderr (where does a slow create go?)
struct CreateTrace {
  bool on;
  std::chrono::steady_clock::time_point t0, last;
  std::string line;
  CreateTrace() {
    const char* e = std::getenv("HTLR_TRACE_CREATE");
    on = e && std::atoi(e) > 0;
    t0 = last = std::chrono::steady_clock::now();
  }
  void mark(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof(buf), " %s=%.2f", what, std::chrono::duration<double, std::milli>(now - last).count());
    line += buf;
    last = now;
  }
  ~CreateTrace() {
    if (on) fprintf(stderr, "[htlr create ms]%s total=%.2f\n", line.c_str(),
                    std::chrono::duration<double, std::milli>(last - t0).count());
  }
};

// NCCL is bound with dlopen on first use instead of being a link-time dependency: only the row-sharded path
// needs it, and a process that also hosts another NCCL user (a PyTorch that bundles its own, newer libnccl.so.2)
// must end up with ONE copy of the library, whichever was loaded first, instead of a version clash at load time.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool ok = false;
};
NcclApi load_nccl() {
  NcclApi api;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
    return api;
  }
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  if (!api.ok) api.error = "libnccl.so.2 lacks an expected symbol";
  return api;
}
NcclApi& nccl_api() {
  static NcclApi api = load_nccl();   // thread-safe one-time initialisation
  return api;
}
NcclApi& nccl_need() {
  NcclApi& a = nccl_api();
  if (!a.ok) throw CudaFail{HTLR_E_NCCL, a.error};
  return a;
}

// Device chunks of destroyed handles, kept for the next create on the same device: a fit per CV fold or per chain
// then pays for cudaMalloc / cudaFree of its arena once per process, not once per fit (a cudaFree of a few hundred
// megabytes was seen to take > 100 ms now and then). Bounded (HTLR_CACHE_MB, default 4096; 0 disables);
// htlr_cuda_release_cache() empties it.
struct ChunkCache {
  struct Item { int device; void* p; size_t size; };
  std::mutex mu;
  std::vector<Item> items;
  size_t bytes = 0;
  static size_t limit() {
    static const size_t lim = [] {
      const char* e = std::getenv("HTLR_CACHE_MB");
      return (size_t)(e ? std::max(0L, std::atol(e)) : 4096L) << 20;
    }();
    return lim;
  }
  void* take(int device, size_t want, size_t* got) {
    std::lock_guard<std::mutex> lk(mu);
    int best = -1;
    for (int i = 0; i < (int)items.size(); ++i)
      if (items[i].device == device && items[i].size >= want && items[i].size <= 2 * want + ((size_t)8 << 20) &&
          (best < 0 || items[i].size < items[best].size))
        best = i;
    if (best < 0) return nullptr;
    void* p = items[best].p;
    *got = items[best].size;
    bytes -= items[best].size;
    items.erase(items.begin() + best);
    return p;
  }
  bool give(int device, void* p, size_t size) {
    std::lock_guard<std::mutex> lk(mu);
    if (bytes + size > limit()) return false;
    items.push_back({device, p, size});
    bytes += size;
    return true;
  }
  void clear() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& it : items) {
      int cur = 0;
      cudaGetDevice(&cur);
      cudaSetDevice(it.device);
      cudaFree(it.p);
      cudaSetDevice(cur);
    }
    items.clear();
    bytes = 0;
  }
};
ChunkCache& chunk_cache() {
  static ChunkCache c;
  return c;
}

struct EventPair {
  cudaEvent_t a, b;
  int kind;  // 0 sweep, 1 step total, 2 sigma, 3 all-reduce
};

}  // namespace

struct htlr_handle {
  htlr_cfg cfg{};
  Geom g{};
  Bufs b{};
  cudaStream_t st = nullptr;
  std::string err;
  int iters_done = 0;  // outer iterations enqueued so far
  std::vector<void*> allocs;
  std::map<int, cudaGraphExec_t> graphs;   // key: 2 * L + use_cluster
  ncclComm_t comm = nullptr;
  bool sharded = false;
  bool fused = false;       // grid-mode fused trajectory kernel applies
  FusedGeom fg{};
  bool has_cluster = false; // cluster-mode flavour applies (fgc)
  FusedGeom fgc{};
  bool use_cluster = false; // launch the cluster-mode flavour ahead of the grid-mode one
  bool use_small = false;   // launch the single-block kernel ahead of both
  int rows_cl = 0;          // > 0: launch the row-partitioned cluster kernel (this cluster size) ahead of all of them
  int rows_uk_max = 0;      //      largest u * K it takes
  bool skip_grid = false;   // the cluster-mode kernel covers every possible restricted set: no grid-mode launch
  int smem_optin = 0;
  unsigned long long small_base = 0;
  unsigned long long tl_base[12] = {};
  bool profiling = false;
  std::vector<EventPair> ev_pool;
  size_t ev_used = 0;
  htlr_profile prof{};
  int64_t launches = 0;
  int kernels_per_step_cache = 0;
  // device scratch for the hooks
  double *t_lv = nullptr, *t_R = nullptr, *t_pred = nullptr, *t_coef = nullptr, *t_g = nullptr, *t_vd = nullptr,
         *t_momt = nullptr, *t_scalar = nullptr, *t_lvsave = nullptr;
  int *t_idx = nullptr, *t_cnt = nullptr;
  double *s_mean = nullptr, *s_sdb = nullptr;   // posterior-summary scratch
  int* s_flag = nullptr;
  Ctl* t_ctlsave = nullptr;
  // inject device buffers (addresses are baked into captured graphs)
  double *inj[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t inj_cap[4] = {0, 0, 0, 0};
  unsigned long long cols_base = 0;  // Ctl::cols_swept at the last profile reset
  // pipelined download (htlr_cuda_collect): an event after every enqueue call, a copy stream of its own
  struct Mark { int iters; cudaEvent_t ev; };
  std::deque<Mark> marks;
  std::vector<cudaEvent_t> mark_free;
  cudaStream_t st_copy = nullptr;
  int slices_collected = 0;
  unsigned long long ph_base[4] = {0, 0, 0, 0};
  double last_predict_ms = 0, last_predict_flops = 0;   // kernels of the last htlr_cuda_predict call (device time)

  // Device memory comes from a few large chunks (one cudaMalloc / cudaFree each): creating and destroying a
  // handle then costs a handful of driver calls instead of ~50, which is what the end-to-end time of a short
  // fit is made of.
  struct Chunk { char* base; size_t size, used; };
  std::vector<Chunk> chunks;
  void arena_reserve(size_t bytes) {
    size_t got = bytes;
    void* p = chunk_cache().take(cfg.device, bytes, &got);
    if (!p) CU(cudaMalloc(&p, bytes));
    allocs.push_back(p);
    chunks.push_back(Chunk{static_cast<char*>(p), got, 0});
  }
  template <class T>
  T* dalloc(size_t count, bool zero = true) {
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
    if (chunks.empty() || chunks.back().used + bytes > chunks.back().size)
      arena_reserve(std::max<size_t>(bytes, (size_t)4 << 20));
    Chunk& c = chunks.back();
    char* p = c.base + c.used;
    c.used += bytes;
    if (zero) CU(cudaMemsetAsync(p, 0, bytes, st));
    return reinterpret_cast<T*>(p);
  }
  int* ctl_int(size_t off) { return reinterpret_cast<int*>(reinterpret_cast<char*>(b.ctl) + off); }
  double* ctl_dbl(size_t off) { return reinterpret_cast<double*>(reinterpret_cast<char*>(b.ctl) + off); }
};

namespace {

void drop_graphs(htlr_handle* h) {
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  h->graphs.clear();
}

int pick_L(const htlr_cfg& c, int i_mc) {
  // Fit::Traject, gibbs.cpp:394-408
  return (i_mc < c.iters_h) ? c.leap_L_h : c.leap_L;
}

// ---- profiling events --------------------------------------------------------------------------------
void prof_drain(htlr_handle* h) {
  if (!h->ev_used) return;
  CU(cudaStreamSynchronize(h->st));
  for (size_t i = 0; i < h->ev_used; ++i) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev_pool[i].a, h->ev_pool[i].b));
    if (h->ev_pool[i].kind == 0) { h->prof.sweep_ms += ms; h->prof.sweep_launches += 1; }
    else if (h->ev_pool[i].kind == 1) h->prof.total_ms += ms;
    else if (h->ev_pool[i].kind == 3) h->prof.allreduce_ms += ms;
    else h->prof.sigma_ms += ms;
  }
  h->ev_used = 0;
}
struct ProfScope {
  htlr_handle* h;
  long idx = -1;  // index into ev_pool (the vector may grow while an outer scope is open)
  ProfScope(htlr_handle* hh, int kind) : h(hh) {
    if (!h->profiling) return;
    if (h->ev_used == h->ev_pool.size()) {
      EventPair p{};
      CU(cudaEventCreate(&p.a));
      CU(cudaEventCreate(&p.b));
      h->ev_pool.push_back(p);
    }
    idx = (long)h->ev_used++;
    h->ev_pool[idx].kind = kind;
    CU(cudaEventRecord(h->ev_pool[idx].a, h->st));
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->ev_pool[idx].b, h->st);
  }
};

// ---- one thin-step -------------------------------------------------------------------------------------
// Order of Fit::StartSampling's loop body, gibbs.cpp:62-98.
int enqueue_step(htlr_handle* h, int L, bool until_energy_only = false) {
  const Geom& g = h->g;
  const Bufs& b = h->b;
  cudaStream_t st = h->st;
  int nk = 0;
  int* d_u = h->ctl_int(offsetof(Ctl, u));
  double* gtarget = (g.g_rt == 1 && !h->sharded) ? b.g : b.gpart;
  ProfScope whole(h, 1);
  auto gradient = [&](int s) {
    {
      ProfScope ps(h, 0);
      launch_grad_sweep(g, b, b.iup, d_u, b.R, gtarget, st);
      ++nk;
    }
    if (h->sharded) {
      launch_gsum(g, b, d_u, b.g, st);
      ++nk;
      ProfScope pa(h, 3);
      NC(nccl_need().AllReduce(b.g, b.g, (size_t)g.nvar * g.K + 1, ncclDouble, ncclSum, h->comm, st));
    }
    launch_post(g, b, s, L, h->sharded ? 1 : 0, st);
    ++nk;
  };
  launch_select(g, b, st); ++nk;
  launch_begin(g, b, st); ++nk;
  if (!h->fused) {
    // the reference's DetachFixlv (gibbs.cpp:197-229) and the softmax before the first gradient; the fused
    // trajectory kernels need neither: they carry lv and the residual across iterations and add only
    // X[:, iup] (delta_new - delta_old) per step
    {
      ProfScope ps(h, 0);
      launch_lv_sweep(g, b, 0, nullptr, nullptr, nullptr, st);
      ++nk;
    }
    launch_softmax(g, b, 0, nullptr, nullptr, nullptr, nullptr, nullptr, st); ++nk;
  }
  if (h->fused) {
    ProfScope ps(h, 0);
    h->fg.timing = h->fgc.timing = h->profiling ? 1 : 0;
    if (h->rows_cl > 0) {
      // few restricted columns (u K small): every block owns a slab of rows and only the gradient is exchanged
      CU(launch_traj_rows(g, b, L, h->rows_cl, h->rows_uk_max, h->smem_optin, st));
      ++nk;
    }
    if (h->use_small) {
      // a restricted set that fits in one SM's shared memory: no inter-block exchange at all
      // measured (tools/small_sweep.py, after the cluster kernel's exchange lost its barriers): it beats the cluster
      // kernel only up to u ~ 12 at n = 62 (108 vs 117 us per iteration at u = 8) and is level with it at u = 8, n = 200 -
      // one SM's FP64 rate bounds the softmax of the rows
      CU(launch_traj_small(g, b, L, h->smem_optin,
                           h->cfg.algo == HTLR_ALGO_SMALL_FIRST ? (1 << 20) : std::max(4, 46000 / (g.n * g.n)), st));
      ++nk;
    }
    if (h->use_cluster) {
      // one cluster runs the whole trajectory out of shared memory when the restricted set is small enough
      // (cl_umax); otherwise it leaves traj_done = 0 and the grid-mode launch does the work
      CU(launch_traj_fused(g, b, h->fgc, L, st));
      ++nk;
    }
    if (!h->skip_grid) {
      CU(launch_traj_fused(g, b, h->fg, L, st));
      ++nk;
    }
  } else {
    gradient(0);
  }
  for (int t = 1; t <= L && !h->fused; ++t) {
    {
      ProfScope ps(h, 0);
      launch_lv_sweep(g, b, 1, nullptr, nullptr, nullptr, st);
      ++nk;
    }
    launch_softmax(g, b, 1, nullptr, nullptr, nullptr, nullptr, nullptr, st); ++nk;
    gradient(t);
  }
  launch_energy(g, b, h->sharded ? 1 : 0, st); ++nk;
  if (until_energy_only) return nk;
  if (h->cfg.ptype == HTLR_PRIOR_T && !(h->cfg.eta > 1e-10)) {
    // the default prior: commit, the gamma draws and the trace recording are one per-feature kernel
    ProfScope ps(h, 2);
    launch_tail_t(g, b, st); ++nk;
    CU(cudaGetLastError());
    return nk;
  }
  launch_commit(g, b, st); ++nk;
  {
    ProfScope ps(h, 2);
    if (h->cfg.ptype == HTLR_PRIOR_T) {
      launch_sigma(g, b, st); ++nk;
      if (h->cfg.eta > 1e-10) { launch_logw(g, b, st); ++nk; }
    } else {
      launch_sigma_ars(g, b, st); ++nk;
    }
  }
  launch_record(g, b, st); ++nk;
  CU(cudaGetLastError());
  return nk;
}

// One graph = `reps` consecutive thin-steps with the same trajectory length (reps > 1 saves the launch gap between
// consecutive graphs: one cudaGraphLaunch per unit of iterations).
constexpr int kGraphReps = 8;
cudaGraphExec_t get_graph(htlr_handle* h, int L, int reps = 1) {
  const int key = (2 * L + (h->use_cluster ? 1 : 0)) * 16 + reps;
  auto it = h->graphs.find(key);
  if (it != h->graphs.end()) return it->second;
  cudaGraph_t graph;
  CU(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
  int nk = 0;
  try {
    for (int r = 0; r < reps; ++r) nk = enqueue_step(h, L);
  } catch (...) {
    cudaGraph_t junk = nullptr;
    cudaStreamEndCapture(h->st, &junk);
    if (junk) cudaGraphDestroy(junk);
    throw;
  }
  CU(cudaStreamEndCapture(h->st, &graph));
  cudaGraphExec_t exec;
  CU(cudaGraphInstantiate(&exec, graph, 0));
  CU(cudaGraphDestroy(graph));
  h->graphs[key] = exec;
  h->kernels_per_step_cache = nk;
  return exec;
}

int kernels_per_step(const htlr_handle* h, int L) {
  int per_grad = h->sharded ? 3 : 2;
  int n = 2 + 2 + per_grad + L * (2 + per_grad) + 1 + 1 + 1 + 1;
  if (h->fused) n = 2 + 1 + 1 + 1 + 1 + (h->skip_grid ? 0 : 1) + (h->use_cluster ? 1 : 0) +
                     (h->use_small ? 1 : 0) + (h->rows_cl > 0 ? 1 : 0);
  if (h->cfg.ptype == HTLR_PRIOR_T && h->cfg.eta > 1e-10) n += 1;
  if (h->cfg.ptype == HTLR_PRIOR_T && !(h->cfg.eta > 1e-10)) n -= 2;
  return n;
}

void check_device_error(htlr_handle* h) {
  Ctl c;
  CU(cudaMemcpy(&c, h->b.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
  if (c.err == kErrNone) return;
  // clear so that the handle stays usable for inspection
  int zero = 0;
  CU(cudaMemcpy(h->ctl_int(offsetof(Ctl, err)), &zero, sizeof(int), cudaMemcpyHostToDevice));
  switch (c.err) {
    case kErrInjectNormals: FAIL(HTLR_E_SAMPLER, "injected normals exhausted");
    case kErrInjectUniforms: FAIL(HTLR_E_SAMPLER, "injected uniforms exhausted");
    case kErrInjectGammas: FAIL(HTLR_E_SAMPLER, "injected gammas exhausted");
    case kErrInjectArs: FAIL(HTLR_E_SAMPLER, "injected ARS uniforms exhausted");
    case kErrClusterCap:
      FAIL(HTLR_E_SAMPLER, "restricted set (" + std::to_string(c.err_detail) +
                               " features) exceeds the capacity of the cluster-only trajectory kernel");
    case kErrArs:
      switch (c.err_detail) {
        case 1: FAIL(HTLR_E_SAMPLER, "Error in adaptive rejection sampling:\nthe first tangent point doesn't have positive probability.\n");
        case 2: FAIL(HTLR_E_SAMPLER, "Error in Rejection Sampling:\n'max_nhull' is set too small, or your log-PDF is NOT concave.\n"
                                     "(device engine: 16 hulls per feature, 32 for logw; the reference's default is 1000)\n");
        case 3: FAIL(HTLR_E_SAMPLER, "Error: in C function 'sample_elin':\nthe exp linear function integrates to NAN/INFINITY\n");
        default: FAIL(HTLR_E_SAMPLER, "adaptive rejection sampling did not terminate");
      }
    default: FAIL(HTLR_E_SAMPLER, "unknown device error " + std::to_string(c.err));
  }
}

void upload_inject(htlr_handle* h, const htlr_inject* inj) {
  const void* src[4] = {inj ? inj->normals : nullptr, inj ? inj->uniforms : nullptr, inj ? inj->gammas : nullptr,
                        inj ? inj->ars_uniforms : nullptr};
  int64_t cnt[4] = {inj ? inj->n_normals : 0, inj ? inj->n_uniforms : 0, inj ? inj->n_gammas : 0,
                    inj ? inj->n_ars_blocks * (int64_t)h->cfg.ars_inject_width : 0};
  for (int q = 0; q < 4; ++q) {
    if (!src[q]) cnt[q] = 0;
    if (cnt[q] > h->inj_cap[q]) {
      CU(cudaStreamSynchronize(h->st));
      drop_graphs(h);
      if (h->inj[q]) CU(cudaFree(h->inj[q]));
      int64_t cap = std::max<int64_t>(cnt[q], 1024);
      CU(cudaMalloc((void**)&h->inj[q], cap * sizeof(double)));
      h->inj_cap[q] = cap;
    }
    if (cnt[q]) CU(cudaMemcpyAsync(h->inj[q], src[q], cnt[q] * sizeof(double), cudaMemcpyHostToDevice, h->st));
  }
  h->b.inj_norm = h->inj[0];
  h->b.inj_unif = h->inj[1];
  h->b.inj_gam = h->inj[2];
  h->b.inj_ars = h->inj[3];
  // cursors restart, sizes updated: cur_norm, cur_unif, cur_gam, n_norm, n_unif, n_gam, n_ars_blocks
  long long hdr[7] = {0, 0, 0, cnt[0], cnt[1], cnt[2], (inj && inj->ars_uniforms) ? inj->n_ars_blocks : 0};
  CU(cudaMemcpyAsync(reinterpret_cast<char*>(h->b.ctl) + offsetof(Ctl, cur_norm), hdr, sizeof(hdr),
                     cudaMemcpyHostToDevice, h->st));
  long long zero = 0;
  CU(cudaMemcpyAsync(reinterpret_cast<char*>(h->b.ctl) + offsetof(Ctl, step_in_call), &zero, sizeof(zero),
                     cudaMemcpyHostToDevice, h->st));
  CU(cudaStreamSynchronize(h->st));  // hdr / zero are stack variables
}

void do_create(const htlr_cfg* cfg, const double* X, int64_t ldx, const double* ymat, const int64_t* ybase,
               const double* deltas0, const double* sigmasbt0, htlr_handle* h) {
  if (!cfg || cfg->struct_size != sizeof(htlr_cfg)) FAIL(HTLR_E_ARG, "htlr_cfg.struct_size mismatch");
  if (cfg->abi_version != HTLR_CUDA_ABI_VERSION) FAIL(HTLR_E_ARG, "htlr_cfg.abi_version mismatch");
  const htlr_cfg& c = *cfg;
  if (c.p < 1 || c.K < 1 || c.n < 1) FAIL(HTLR_E_ARG, "p, K, n must be positive");
  if (c.K > 4096) FAIL(HTLR_E_ARG, "more than 4097 classes are not supported");
  if (c.ptype < 0 || c.ptype > 2) FAIL(HTLR_E_ARG, "Unsupported prior type");
  if (!(c.iters_rmc > 0 && c.iters_h >= 0 && c.thin > 0 && c.leap_L > 0 && c.leap_L_h > 0))
    FAIL(HTLR_E_ARG, "iters_rmc > 0, iters_h >= 0, thin > 0, leap_L > 0, leap_L_h > 0 required");
  if (!(c.alpha > 0) || !(c.eta >= 0)) FAIL(HTLR_E_ARG, "alpha > 0 and eta >= 0 required");
  if (!X || !ymat || !ybase || !deltas0 || !sigmasbt0) FAIL(HTLR_E_ARG, "null input array");
  if (ldx < c.n) FAIL(HTLR_E_ARG, "ldx < n");
  if (c.rng_mode != HTLR_RNG_DEVICE && c.rng_mode != HTLR_RNG_INJECT) FAIL(HTLR_E_ARG, "bad rng_mode");
  CreateTrace tr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    FAIL(HTLR_E_CUDA, "no CUDA device visible: this library has no CPU fallback");
  if (c.device < 0 || c.device >= ndev) FAIL(HTLR_E_ARG, "device ordinal out of range");
  CU(cudaSetDevice(c.device));
  const DevInfo& dev = dev_info(c.device);
  if (dev.major != 10) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c.device));
    FAIL(HTLR_E_CUDA, std::string("built for sm_100a only; device is ") + prop.name);
  }
  tr.mark("props");

  h->cfg = c;
  h->sharded = c.nccl_comm != nullptr;
  h->comm = static_cast<ncclComm_t>(c.nccl_comm);
  CU(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));

  const int n = c.n, K = c.K, nvar = c.p + 1;
  Geom& g = h->g;
  g.n = n; g.nvar = nvar; g.K = K;
  g.n_s = (int)round_up(n, 16);
  g.sm_count = dev.sm_count;
  // Sweep grids: ONE full wave of resident blocks; a block walks its kernel's (row tile, column split) items with the
  // grid's stride, and the number of splits is chosen so that every block gets the same number of items. (Grids of
  // 1.06 / 1.32 waves - a second wave of a few blocks that cannot keep HBM busy - cost 30 % at 200000 rows:
  // profiles/r02_ncu_tall_sweeps.txt.)
  auto pick_split = [](int tiles, int wave, int max_split) {
    int best = 1;
    double best_eff = 0;
    for (int sp = 1; sp <= max_split; ++sp) {
      const int64_t items = (int64_t)tiles * sp;
      const double eff = (double)items / ((double)((items + wave - 1) / wave) * wave);
      if (eff > best_eff + 0.02) { best_eff = eff; best = sp; }   // a larger split only for a real gain
    }
    return best;
  };
  g.lv_rt = (n + 511) / 512;
  const int lv_wave = g.sm_count * lv_sweep_blocks_per_sm(K);
  g.lv_cs = pick_split(g.lv_rt, lv_wave, 64);
  g.lv_grid = (int)std::min<int64_t>((int64_t)g.lv_rt * g.lv_cs, lv_wave);
  g.g_tr = grad_sweep_tile_rows(K);
  g.g_rt = (n + g.g_tr - 1) / g.g_tr;
  const int g_wave = g.sm_count * grad_sweep_blocks_per_sm(K);
  g.g_gx = pick_split(g.g_rt, g_wave, 1024);
  g.g_grid = (int)std::min<int64_t>((int64_t)g.g_rt * g.g_gx, g_wave);
  g.e_grid = std::max(1, std::min((nvar + 255) / 256, 2 * g.sm_count));

  Bufs& b = h->b;
  {
    // everything create allocates, in one chunk (upper bound of the dalloc calls below + alignment slack)
    const size_t Hh = (size_t)c.iters_rmc + 1 + (c.keep_warmup_hist ? c.iters_h : 0);
    const size_t nkk = (size_t)g.n_s * K, nv = (size_t)nvar + 64;   // (+64: one more int array than doubles counted)
    size_t dbl = (c.x_on_device ? 0 : (size_t)round_up(n, 16) * nv) + nkk + nv + nv * K + 2 * nv + 5 * nkk +
                 2 * nv * K + 3 * nv + nv * K + 1 + (size_t)g.g_rt * nv * K +
                 (size_t)std::max(g.lv_cs, g.sm_count) * K * g.n_s + Hh * (nv * (K + 2) + 4);
    size_t bytes = dbl * sizeof(double) + ((size_t)n + 3 * nv) * sizeof(int) + nv +
                   ((size_t)std::max(std::max(g.e_grid, (n + 255) / 256), g.sm_count) + g.e_grid + 8 + c.iters_h + c.iters_rmc) * sizeof(Red4) +
                   sizeof(Ctl) + 64 * 256;
    h->arena_reserve(bytes);
  }
  tr.mark("stream+arena");
  // ---- data
  if (c.x_on_device && (reinterpret_cast<uintptr_t>(X) & 15))
    FAIL(HTLR_E_ARG, "device-resident X must be 16-byte aligned (whole columns are fetched with bulk copies)");
  if (c.x_on_device) {
    if (ldx % 2) FAIL(HTLR_E_ARG, "device-resident X needs an even leading dimension");
    g.ld = ldx;
    // the fused kernels copy whole row pairs, so row n of an odd n must be finite - make it zero
    if ((n & 1) && ldx > n)
      CU(cudaMemset2DAsync(const_cast<double*>(X) + n, ldx * sizeof(double), 0, sizeof(double), nvar, h->st));
    b.X = X;
  } else {
    g.ld = round_up(n, 16);
    double* dX = h->dalloc<double>((size_t)g.ld * nvar, g.ld != n);
    CU(cudaMemcpy2DAsync(dX, g.ld * sizeof(double), X, ldx * sizeof(double), (size_t)n * sizeof(double), nvar,
                         cudaMemcpyHostToDevice, h->st));
    b.X = dX;
  }
  tr.mark("X_enqueue");
  std::vector<int64_t> yb_host(n);
  if (c.x_on_device) CU(cudaMemcpy(yb_host.data(), ybase, n * sizeof(int64_t), cudaMemcpyDeviceToHost));
  else std::memcpy(yb_host.data(), ybase, n * sizeof(int64_t));
  std::vector<int> yb32(n);
  for (int i = 0; i < n; ++i) {
    if (yb_host[i] < 0 || yb_host[i] > K) FAIL(HTLR_E_ARG, "ybase out of range [0, C)");
    yb32[i] = (int)yb_host[i];
  }
  int* d_yb = h->dalloc<int>(n);
  CU(cudaMemcpyAsync(d_yb, yb32.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->st));
  b.ybase = d_yb;
  double* d_ymat = h->dalloc<double>((size_t)g.n_s * K);
  CU(cudaMemcpy2DAsync(d_ymat, g.n_s * sizeof(double), ymat, (size_t)n * sizeof(double), (size_t)n * sizeof(double),
                       K, c.x_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->st));
  b.ymat = d_ymat;
  double* d_ddn = h->dalloc<double>(nvar);
  b.ddn = d_ddn;

  // ---- state
  b.deltas = h->dalloc<double>((size_t)nvar * K);
  b.sigmasbt = h->dalloc<double>(nvar);
  b.var_deltas = h->dalloc<double>(nvar);
  CU(cudaMemcpyAsync(b.deltas, deltas0, (size_t)nvar * K * sizeof(double), cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(b.sigmasbt, sigmasbt0, (size_t)nvar * sizeof(double), cudaMemcpyHostToDevice, h->st));
  const size_t nk = (size_t)g.n_s * K;
  b.lv = h->dalloc<double>(nk);
  b.lv_old = h->dalloc<double>(nk);
  b.lv_fix = h->dalloc<double>(nk);
  b.R_old = h->dalloc<double>(nk);
  b.R = h->dalloc<double>(nk);
  b.iup = h->dalloc<int>(nvar);
  b.ifix = h->dalloc<int>(nvar);
  b.posof = h->dalloc<int>(nvar);
  b.dc = h->dalloc<double>((size_t)nvar * K);
  b.momt = h->dalloc<double>((size_t)nvar * K);
  b.stepc = h->dalloc<double>(nvar);
  b.sigc = h->dalloc<double>(nvar);
  b.vdc = h->dalloc<double>(nvar);
  b.g = h->dalloc<double>((size_t)nvar * K + 1);
  b.gpart = h->dalloc<double>((size_t)g.g_rt * nvar * K);
  tr.mark("state_enqueue");
  // ---- trajectory algorithm
  {
    FusedGeom fgm{};
    const bool can = dev.coop && !h->sharded && fused_plan(g, dev.smem_optin, 0, &fgm);
    if ((c.algo == HTLR_ALGO_FUSED || c.algo == HTLR_ALGO_FUSED_CLUSTER || c.algo == HTLR_ALGO_CLUSTER_ONLY) && !can)
      FAIL(HTLR_E_ARG, "fused trajectory not applicable (needs K <= 8, a column tile in shared memory, one GPU)");
    h->fused = can && c.algo != HTLR_ALGO_TWO_SWEEP;
    h->fg = fgm;
    if (h->fused && c.algo != HTLR_ALGO_FUSED) {   // AUTO, FUSED_CLUSTER, CLUSTER_ONLY, FUSED_NO_SMALL, SMALL_FIRST
      const int cl = fused_cluster_size(g, dev.smem_optin);
      h->has_cluster = cl > 0 && fused_plan(g, dev.smem_optin, cl, &h->fgc);
      if (h->has_cluster) {
        // Which flavour runs is decided on the device, per iteration, from the restricted-set size: both are
        // launched and the one that is not wanted returns at once (~2.5 us). Grid mode (148 blocks, exchange
        // through L2) is bandwidth efficient; cluster mode (one cluster, exchange through shared memory,
        // hardware barriers) has lower latency per leapfrog step but a fraction of the bandwidth: measured
        // crossover at about 1.5 MB of X[:, iup] (tools/mode_sweep.py).
        const int cap = h->fgc.cl * h->fgc.ncmax;
        // (re-measured after the cluster-mode exchange lost its barriers: crossover at 1.5 MB for n = 500 / K = 3,
        // 2.0 MB for n = 200 / K = 1, 2.6 - 2.8 MB for n = 1000 and 2000 / K = 2)
        const double cross = K >= 3 ? 1.5e6 : (n >= 1000 ? 2.5e6 : 1.9e6);
        const int by_bytes = (int)std::min<double>(cross / (8.0 * n), 2.0e9);
        const bool forced = c.algo == HTLR_ALGO_FUSED_CLUSTER || c.algo == HTLR_ALGO_CLUSTER_ONLY;
        h->fgc.cl_umax = forced ? cap : std::min(cap, by_bytes);
        h->fgc.cl_only = c.algo == HTLR_ALGO_CLUSTER_ONLY ? 1 : 0;
      }
    }
    if ((c.algo == HTLR_ALGO_FUSED_CLUSTER || c.algo == HTLR_ALGO_CLUSTER_ONLY) && !h->has_cluster)
      FAIL(HTLR_E_ARG, "cluster-mode trajectory kernel not available on this device");
    h->use_cluster = h->fused && h->has_cluster;
    // small problems (config A: 2001 features): even u = nvar stays below the cluster kernel's limit, so the grid-mode
    // launch could only ever return at once (1.5 us per iteration)
    h->skip_grid = c.algo == HTLR_ALGO_CLUSTER_ONLY ||
                   (c.algo == HTLR_ALGO_AUTO && h->use_cluster && h->fgc.cl_umax >= nvar);
    if (h->skip_grid) h->fgc.cl_only = 1;   // nothing follows the cluster kernel: exceeding its limit is an error
    h->smem_optin = dev.smem_optin;
    h->use_small = h->fused && small_applies(g) &&
                   (((c.algo == HTLR_ALGO_AUTO || c.algo == HTLR_ALGO_CLUSTER_ONLY) && g.n <= 256) ||
                    c.algo == HTLR_ALGO_SMALL_FIRST);
    // Row-partitioned kernel: largest restricted set (columns) it is handed, from tools/rows_sweep.py (us per iteration
    // against k_traj's cluster mode over n = 300..2000, K = 1..4, u = 4..128, profiles/r02_rules/rows_sweep.jsonl;
    // e.g. n = 1000 / K = 4 / u = 8: 185 vs 338, n = 500 / K = 3 / u = 12: 164 vs 208, n = 1000 / K = 2 / u = 32: 187 vs 201,
    // n = 500 / K = 1: never). Rows: K = 1..4; columns: n < 800, n < 1500, n >= 1500.
    static const int kRowsUMax[4][3] = {{0, 16, 32}, {32, 32, 32}, {24, 32, 64}, {32, 64, 64}};
    const int rows_u_max = (K <= 4 && n >= 300) ? kRowsUMax[K - 1][n >= 1500 ? 2 : (n >= 800 ? 1 : 0)] : 0;
    if (h->fused && (((c.algo == HTLR_ALGO_AUTO || c.algo == HTLR_ALGO_CLUSTER_ONLY) && rows_u_max > 0) ||
                     c.algo == HTLR_ALGO_ROWS_FIRST)) {
      h->rows_cl = rows_cluster_size(g, dev.smem_optin);
      h->rows_uk_max = c.algo == HTLR_ALGO_ROWS_FIRST ? (1 << 20) : rows_u_max * K;
      if (const char* e = std::getenv("HTLR_ROWS_UK_MAX")) h->rows_uk_max = std::atoi(e);   // sweeps only
      if (h->rows_uk_max <= 0) h->rows_cl = 0;
    }
  }
  b.lvpart = h->dalloc<double>((size_t)std::max(g.lv_cs, g.sm_count) * K * g.n_s);
  b.red = h->dalloc<Red4>((size_t)std::max(std::max(g.e_grid, (n + 255) / 256), g.sm_count));
  b.red_begin = h->dalloc<Red4>((size_t)g.e_grid);
  b.ctl = h->dalloc<Ctl>(1);

  // ---- history
  const int H = c.iters_rmc + 1 + (c.keep_warmup_hist ? c.iters_h : 0);
  b.mcdeltas = h->dalloc<double>((size_t)nvar * K * H);
  b.mcsigmasbt = h->dalloc<double>((size_t)nvar * H);
  b.mcvardeltas = h->dalloc<double>((size_t)nvar * H);
  b.mclogw = h->dalloc<double>(H);
  b.mcloglike = h->dalloc<double>(H);
  b.mcuvar = h->dalloc<double>(H);
  b.mchmcrej = h->dalloc<double>(H);
  b.stats = h->dalloc<Red4>(c.iters_h + c.iters_rmc);

  // ---- control block
  Ctl ctl;
  std::memset(&ctl, 0, sizeof(ctl));
  ctl.nvar = nvar; ctl.K = K; ctl.C = K + 1; ctl.n = n; ctl.p = c.p;
  ctl.iters_h = c.iters_h; ctl.iters_rmc = c.iters_rmc; ctl.thin = c.thin;
  ctl.keep_warmup_hist = c.keep_warmup_hist ? 1 : 0;
  ctl.ptype = c.ptype; ctl.rng_mode = c.rng_mode;
  {
    const char* e = getenv("HTLR_TIMELINE");
    ctl.timeline = e && atoi(e) > 0;
  }
  ctl.alpha = c.alpha; ctl.s = c.s; ctl.eta = c.eta; ctl.leap_step = c.leap_step;
  ctl.sgmsq_cut = c.hmc_sgmcut > 0 ? c.hmc_sgmcut * c.hmc_sgmcut : c.hmc_sgmcut;  // gibbs.cpp:14
  ctl.seed = c.seed;
  ctl.logw = c.s;
  ctl.init_all = 1;
  ctl.ars_width = c.ars_inject_width;
  CU(cudaMemcpyAsync(b.ctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, h->st));
  tr.mark("plan");
  CU(cudaStreamSynchronize(h->st));  // ctl / yb32 are host temporaries
  tr.mark("sync_uploads");

  // ---- Fit::Fit tail + Fit::Initialize (gibbs.cpp:15, 519-530)
  launch_colsumsq(g, b.X, d_ddn, h->st);
  if (h->sharded) NC(nccl_need().AllReduce(d_ddn, d_ddn, nvar, ncclDouble, ncclSum, h->comm, h->st));
  launch_select(g, b, h->st);
  int* d_u = h->ctl_int(offsetof(Ctl, u));
  launch_lv_sweep(g, b, 2, b.iup, d_u, b.deltas, h->st);
  double* d_ll = h->ctl_dbl(offsetof(Ctl, loglike));
  launch_softmax(g, b, 2, d_u, b.lv, b.R, nullptr, d_ll, h->st);
  if (h->sharded) NC(nccl_need().AllReduce(d_ll, d_ll, 1, ncclDouble, ncclSum, h->comm, h->st));
  launch_vardeltas_all(g, b, h->st);
  h->launches += 5;
  CU(cudaMemcpyAsync(b.mcdeltas, b.deltas, (size_t)nvar * K * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  CU(cudaMemcpyAsync(b.mcsigmasbt, b.sigmasbt, (size_t)nvar * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  CU(cudaMemcpyAsync(b.mcvardeltas, b.var_deltas, (size_t)nvar * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  CU(cudaMemcpyAsync(b.mcloglike, d_ll, sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  CU(cudaMemcpyAsync(b.mclogw, h->ctl_dbl(offsetof(Ctl, logw)), sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  int zero = 0;
  CU(cudaMemcpyAsync(h->ctl_int(offsetof(Ctl, init_all)), &zero, sizeof(int), cudaMemcpyHostToDevice, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  tr.mark("initialize");
}

void ensure_scratch(htlr_handle* h) {
  if (h->t_lv) return;
  const Geom& g = h->g;
  const size_t nk = (size_t)g.n_s * g.K;
  h->t_lv = h->dalloc<double>(nk);
  h->t_R = h->dalloc<double>(nk);
  h->t_lvsave = h->dalloc<double>(nk);
  h->t_pred = h->dalloc<double>((size_t)g.n * (g.K + 1));
  h->t_coef = h->dalloc<double>((size_t)g.nvar * g.K);
  h->t_momt = h->dalloc<double>((size_t)g.nvar * g.K);
  h->t_g = h->dalloc<double>((size_t)g.g_rt * g.nvar * g.K + 1);
  h->t_vd = h->dalloc<double>(g.nvar);
  h->t_scalar = h->dalloc<double>(4);
  h->t_idx = h->dalloc<int>(g.nvar);
  h->t_cnt = h->dalloc<int>(1);
  h->t_ctlsave = h->dalloc<Ctl>(1);
}

// [k][i] stride n_s  ->  n x C column-major with a zero first column
void lv_to_host(htlr_handle* h, const double* d_lv, double* out) {
  const Geom& g = h->g;
  std::memset(out, 0, sizeof(double) * g.n);
  CU(cudaMemcpy2D(out + g.n, (size_t)g.n * sizeof(double), d_lv, (size_t)g.n_s * sizeof(double),
                  (size_t)g.n * sizeof(double), g.K, cudaMemcpyDeviceToHost));
}

template <class F>
int guarded(htlr_handle* h, F&& f) {
  try {
    f();
    return HTLR_OK;
  } catch (const CudaFail& e) {
    if (h) h->err = e.msg; else g_err = e.msg;
    g_err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    if (h) h->err = e.what();
    g_err = e.what();
    return HTLR_E_STATE;
  }
}

}  // namespace

extern "C" {

int32_t htlr_cuda_abi_version(void) { return HTLR_CUDA_ABI_VERSION; }

const char* htlr_cuda_last_error(const htlr_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int htlr_cuda_create(const htlr_cfg* cfg, const double* X, int64_t ldx, const double* ymat, const int64_t* ybase,
                     const double* deltas0, const double* sigmasbt0, htlr_handle** out) {
  if (!out) { g_err = "null out pointer"; return HTLR_E_ARG; }
  *out = nullptr;
  htlr_handle* h = new htlr_handle();
  int rc = guarded(nullptr, [&] { do_create(cfg, X, ldx, ymat, ybase, deltas0, sigmasbt0, h); });
  if (rc != HTLR_OK) {
    htlr_cuda_destroy(h);
    return rc;
  }
  *out = h;
  return HTLR_OK;
}

void htlr_cuda_destroy(htlr_handle* h) {
  if (!h) return;
  if (h->st) cudaStreamSynchronize(h->st);
  drop_graphs(h);
  for (auto& p : h->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto& m : h->marks) cudaEventDestroy(m.ev);
  for (cudaEvent_t e : h->mark_free) cudaEventDestroy(e);
  if (h->st_copy) { cudaStreamSynchronize(h->st_copy); cudaStreamDestroy(h->st_copy); }
  for (auto& c : h->chunks)
    if (!chunk_cache().give(h->cfg.device, c.base, c.size)) cudaFree(c.base);
  for (int q = 0; q < 4; ++q) if (h->inj[q]) cudaFree(h->inj[q]);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
}

void htlr_cuda_release_cache(void) { chunk_cache().clear(); }

int32_t htlr_cuda_hist_len(const htlr_handle* h) {
  return h->cfg.iters_rmc + 1 + (h->cfg.keep_warmup_hist ? h->cfg.iters_h : 0);
}

static void enqueue_iters(htlr_handle* h, int32_t n_iters) {
  const htlr_cfg& c = h->cfg;
  if (n_iters < 0 || h->iters_done + n_iters > c.iters_h + c.iters_rmc)
    FAIL(HTLR_E_STATE, "run past the configured number of iterations");
  CU(cudaSetDevice(c.device));
  const bool graph = c.use_graph && !h->profiling;
  static const bool multi = [] { const char* e = std::getenv("HTLR_GRAPH_REPS"); return !e || std::atoi(e) != 1; }();
  for (int it = 0; it < n_iters; ++it) {
    const int L = pick_L(c, h->iters_done);
    if (graph && multi && c.thin == 1 && it + kGraphReps <= n_iters &&
        pick_L(c, h->iters_done + kGraphReps - 1) == L) {
      // kGraphReps iterations of the same phase in one launch (pick_L is monotone in the iteration index)
      CU(cudaGraphLaunch(get_graph(h, L, kGraphReps), h->st));
      h->launches += (int64_t)kGraphReps * kernels_per_step(h, L);
      h->iters_done += kGraphReps;
      it += kGraphReps - 1;
      continue;
    }
    for (int t = 0; t < c.thin; ++t) {
      if (graph) {
        cudaGraphExec_t ge = get_graph(h, L);
        CU(cudaGraphLaunch(ge, h->st));
        h->launches += kernels_per_step(h, L);
      } else {
        if (h->ev_used > 8192) prof_drain(h);
        h->launches += enqueue_step(h, L);
      }
    }
    h->iters_done += 1;
  }
  if (n_iters > 0) {
    // completion mark of this call (htlr_cuda_collect waits on it instead of on the whole stream)
    cudaEvent_t ev;
    if (!h->mark_free.empty()) { ev = h->mark_free.back(); h->mark_free.pop_back(); }
    else CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ev, h->st));
    h->marks.push_back({h->iters_done, ev});
    while (h->marks.size() > 256) {   // nobody is collecting: a later mark covers the earlier ones
      h->mark_free.push_back(h->marks.front().ev);
      h->marks.pop_front();
    }
  }
}

int htlr_cuda_run_async(htlr_handle* h, int32_t n_iters) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    if (h->cfg.rng_mode == HTLR_RNG_INJECT) FAIL(HTLR_E_STATE, "run_async needs the device RNG");
    enqueue_iters(h, n_iters);
  });
}

int htlr_cuda_sync(htlr_handle* h) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    CU(cudaStreamSynchronize(h->st));
    check_device_error(h);
  });
}

int htlr_cuda_run(htlr_handle* h, int32_t n_iters, const htlr_inject* inject, htlr_iter_stats* stats) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    CU(cudaSetDevice(h->cfg.device));
    if (h->cfg.rng_mode == HTLR_RNG_INJECT) {
      if (!inject) FAIL(HTLR_E_ARG, "rng_mode INJECT needs an htlr_inject");
      upload_inject(h, inject);
    }
    const int first = h->iters_done;
    enqueue_iters(h, n_iters);
    CU(cudaStreamSynchronize(h->st));
    check_device_error(h);
    if (stats && n_iters > 0) {
      static_assert(sizeof(htlr_iter_stats) == sizeof(Red4), "stats layout");
      CU(cudaMemcpy(stats, h->b.stats + first, sizeof(Red4) * n_iters, cudaMemcpyDeviceToHost));
    }
  });
}

int htlr_cuda_fetch(htlr_handle* h, double* mcdeltas, double* mcsigmasbt, double* mcvardeltas, double* mclogw,
                    double* mcloglike, double* mcuvar, double* mchmcrej, double* ddnloglike) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->st));
    const size_t H = htlr_cuda_hist_len(h), nvar = h->g.nvar, K = h->g.K;
    auto get = [&](double* dst, const double* src, size_t cnt) {
      if (dst) CU(cudaMemcpy(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    };
    get(mcdeltas, h->b.mcdeltas, nvar * K * H);
    get(mcsigmasbt, h->b.mcsigmasbt, nvar * H);
    get(mcvardeltas, h->b.mcvardeltas, nvar * H);
    get(mclogw, h->b.mclogw, H);
    get(mcloglike, h->b.mcloglike, H);
    get(mcuvar, h->b.mcuvar, H);
    get(mchmcrej, h->b.mchmcrej, H);
    get(ddnloglike, h->b.ddn, nvar);
  });
}

int htlr_cuda_collect(htlr_handle* h, int32_t upto_iters, int32_t stats_from, htlr_iter_stats* stats, double* mcdeltas,
                      double* mcsigmasbt, double* mcvardeltas, double* mclogw, double* mcloglike, double* mcuvar,
                      double* mchmcrej) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    const htlr_cfg& c = h->cfg;
    if (upto_iters < 0 || upto_iters > h->iters_done) FAIL(HTLR_E_STATE, "collect past the enqueued iterations");
    if (stats && (stats_from < 0 || stats_from > upto_iters)) FAIL(HTLR_E_ARG, "bad stats range");
    CU(cudaSetDevice(c.device));
    if (!h->st_copy) CU(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking));
    // the earliest completion mark that covers `upto_iters`; later work on the compute stream is not waited for
    bool covered = upto_iters == 0;
    while (!h->marks.empty() && !covered) {
      htlr_handle::Mark m = h->marks.front();
      if (m.iters >= upto_iters) {
        CU(cudaStreamWaitEvent(h->st_copy, m.ev, 0));
        covered = true;
        if (m.iters > upto_iters) break;   // keep it for the next call
      }
      h->mark_free.push_back(m.ev);
      h->marks.pop_front();
    }
    if (!covered) CU(cudaStreamSynchronize(h->st));   // marks were recycled: fall back to the whole stream
    const size_t nvar = h->g.nvar, K = h->g.K;
    const int avail = 1 + (c.keep_warmup_hist ? upto_iters : std::max(0, upto_iters - c.iters_h));
    const int s0 = h->slices_collected, ns = avail - s0;
    auto get = [&](double* dst, const double* src, size_t per_slice) {
      if (dst && ns > 0)
        CU(cudaMemcpyAsync(dst + per_slice * s0, src + per_slice * s0, per_slice * ns * sizeof(double),
                           cudaMemcpyDeviceToHost, h->st_copy));
    };
    get(mcdeltas, h->b.mcdeltas, nvar * K);
    get(mcsigmasbt, h->b.mcsigmasbt, nvar);
    get(mcvardeltas, h->b.mcvardeltas, nvar);
    get(mclogw, h->b.mclogw, 1);
    get(mcloglike, h->b.mcloglike, 1);
    get(mcuvar, h->b.mcuvar, 1);
    get(mchmcrej, h->b.mchmcrej, 1);
    if (stats && upto_iters > stats_from) {
      static_assert(sizeof(htlr_iter_stats) == sizeof(Red4), "stats layout");
      CU(cudaMemcpyAsync(stats, h->b.stats + stats_from, sizeof(Red4) * (upto_iters - stats_from),
                         cudaMemcpyDeviceToHost, h->st_copy));
    }
    CU(cudaStreamSynchronize(h->st_copy));
    if (ns > 0) h->slices_collected = avail;
    check_device_error(h);
  });
}

int htlr_cuda_get_state(htlr_handle* h, double* lv, double* deltas, double* sigmasbt, double* var_deltas,
                        double* loglike, double* logw) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->st));
    const size_t nvar = h->g.nvar, K = h->g.K;
    if (lv) lv_to_host(h, h->b.lv, lv);
    if (deltas) CU(cudaMemcpy(deltas, h->b.deltas, nvar * K * sizeof(double), cudaMemcpyDeviceToHost));
    if (sigmasbt) CU(cudaMemcpy(sigmasbt, h->b.sigmasbt, nvar * sizeof(double), cudaMemcpyDeviceToHost));
    if (var_deltas) CU(cudaMemcpy(var_deltas, h->b.var_deltas, nvar * sizeof(double), cudaMemcpyDeviceToHost));
    Ctl c;
    CU(cudaMemcpy(&c, h->b.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    if (loglike) *loglike = c.loglike;
    if (logw) *logw = c.logw;
  });
}

int htlr_cuda_eval(htlr_handle* h, const int32_t* iup, int32_t u, const double* deltas, double* lv, double* pred,
                   double* loglike, double* grad) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    if (!iup || !deltas || u < 0 || u > h->g.nvar) FAIL(HTLR_E_ARG, "bad restricted set");
    if (h->sharded) FAIL(HTLR_E_STATE, "eval hook is single-GPU only");
    CU(cudaSetDevice(h->cfg.device));
    ensure_scratch(h);
    const Geom& g = h->g;
    const size_t nvar = g.nvar, K = g.K;
    for (int q = 0; q < u; ++q)
      if (iup[q] < 0 || iup[q] >= g.nvar) FAIL(HTLR_E_ARG, "restricted-set index out of range");
    CU(cudaMemcpyAsync(h->t_idx, iup, sizeof(int) * u, cudaMemcpyHostToDevice, h->st));
    CU(cudaMemcpyAsync(h->t_cnt, &u, sizeof(int), cudaMemcpyHostToDevice, h->st));
    CU(cudaMemcpyAsync(h->t_coef, deltas, sizeof(double) * nvar * K, cudaMemcpyHostToDevice, h->st));
    launch_lv_sweep(g, h->b, 2, h->t_idx, h->t_cnt, h->t_coef, h->st);
    launch_softmax(g, h->b, 2, h->t_cnt, h->t_lv, h->t_R, h->t_pred, h->t_scalar, h->st);
    launch_grad_sweep(g, h->b, h->t_idx, h->t_cnt, h->t_R, h->t_g, h->st);
    h->launches += 3;
    CU(cudaStreamSynchronize(h->st));
    CU(cudaGetLastError());
    if (lv) lv_to_host(h, h->t_lv, lv);
    if (pred) CU(cudaMemcpy(pred, h->t_pred, sizeof(double) * g.n * (K + 1), cudaMemcpyDeviceToHost));
    if (loglike) CU(cudaMemcpy(loglike, h->t_scalar, sizeof(double), cudaMemcpyDeviceToHost));
    if (grad) {
      std::vector<double> part((size_t)g.g_rt * nvar * K);
      CU(cudaMemcpy(part.data(), h->t_g, part.size() * sizeof(double), cudaMemcpyDeviceToHost));
      // compact [r*K + k] (summed over row tiles in tile order) -> u x K column-major
      for (int q = 0; q < u; ++q)
        for (size_t k = 0; k < K; ++k) {
          double a = 0;
          for (int rt = 0; rt < g.g_rt; ++rt) a += part[(size_t)rt * nvar * K + (size_t)q * K + k];
          grad[k * u + q] = a;
        }
    }
  });
}

int htlr_cuda_trajectory(htlr_handle* h, const double* momt, int32_t L, double* deltas_out, double* momt_out,
                         double* lv_out, double* loglike_out, double* var_deltas_out, double* nenergy_old,
                         double* nenergy_new, int32_t* nuvar) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    if (!momt || L < 1) FAIL(HTLR_E_ARG, "bad trajectory arguments");
    CU(cudaSetDevice(h->cfg.device));
    ensure_scratch(h);
    const Geom& g = h->g;
    const size_t nvar = g.nvar, K = g.K, nk = (size_t)g.n_s * K;
    CU(cudaStreamSynchronize(h->st));
    // snapshot what a proposal touches: control block and lv (deltas / var_deltas change only on commit)
    CU(cudaMemcpyAsync(h->t_ctlsave, h->b.ctl, sizeof(Ctl), cudaMemcpyDeviceToDevice, h->st));
    CU(cudaMemcpyAsync(h->t_lvsave, h->b.lv, nk * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CU(cudaMemcpyAsync(h->t_momt, momt, nvar * K * sizeof(double), cudaMemcpyHostToDevice, h->st));
    int dev_rng = 0;
    CU(cudaMemcpyAsync(h->ctl_int(offsetof(Ctl, rng_mode)), &dev_rng, sizeof(int), cudaMemcpyHostToDevice, h->st));
    const bool prof = h->profiling;
    h->profiling = false;
    h->b.user_momt = h->t_momt;
    try {
      h->launches += enqueue_step(h, L, /*until_energy_only=*/true);
    } catch (...) {
      h->b.user_momt = nullptr;
      h->profiling = prof;
      throw;
    }
    h->b.user_momt = nullptr;
    h->profiling = prof;
    // proposal in full layout
    CU(cudaMemcpyAsync(h->t_coef, h->b.deltas, nvar * K * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CU(cudaMemsetAsync(h->t_momt, 0, nvar * K * sizeof(double), h->st));
    CU(cudaMemcpyAsync(h->t_vd, h->b.var_deltas, nvar * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    launch_scatter_proposal(g, h->b, h->t_coef, h->t_momt, h->t_vd, h->st);
    h->launches += 1;
    CU(cudaStreamSynchronize(h->st));
    CU(cudaGetLastError());
    Ctl c;
    CU(cudaMemcpy(&c, h->b.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    if (deltas_out) CU(cudaMemcpy(deltas_out, h->t_coef, nvar * K * sizeof(double), cudaMemcpyDeviceToHost));
    if (momt_out) CU(cudaMemcpy(momt_out, h->t_momt, nvar * K * sizeof(double), cudaMemcpyDeviceToHost));
    if (var_deltas_out) CU(cudaMemcpy(var_deltas_out, h->t_vd, nvar * sizeof(double), cudaMemcpyDeviceToHost));
    if (lv_out) lv_to_host(h, h->b.lv, lv_out);
    if (loglike_out) *loglike_out = c.loglike_new;
    if (nenergy_old) *nenergy_old = c.nenergy_old;
    if (nenergy_new) *nenergy_new = c.nenergy_new;
    if (nuvar) *nuvar = c.u;
    // restore
    CU(cudaMemcpyAsync(h->b.ctl, h->t_ctlsave, sizeof(Ctl), cudaMemcpyDeviceToDevice, h->st));
    CU(cudaMemcpyAsync(h->b.lv, h->t_lvsave, nk * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    CU(cudaMemcpyAsync(h->b.R, h->b.R_old, nk * sizeof(double), cudaMemcpyDeviceToDevice, h->st));   // k_begin's copy
    CU(cudaStreamSynchronize(h->st));
  });
}

int htlr_cuda_ars(int32_t device, int32_t kind, int32_t m, const double* var_deltas, double K, double alpha,
                  double log_aw, const double* uniforms, int32_t width, uint64_t seed, uint64_t step, double* x_out,
                  int32_t* n_unif, int32_t* n_eval, int32_t* status) {
  return guarded(nullptr, [&] {
    if (m < 1 || !var_deltas || !x_out) FAIL(HTLR_E_ARG, "bad ARS arguments");
    if (kind != HTLR_PRIOR_GHS && kind != HTLR_PRIOR_NEG) FAIL(HTLR_E_ARG, "kind must be ghs or neg");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(HTLR_E_CUDA, "no CUDA device visible");
    CU(cudaSetDevice(device));
    double *d_v = nullptr, *d_u = nullptr, *d_x = nullptr;
    int *d_nu = nullptr, *d_ne = nullptr, *d_st = nullptr;
    CU(cudaMalloc((void**)&d_v, sizeof(double) * m));
    CU(cudaMalloc((void**)&d_x, sizeof(double) * m));
    CU(cudaMalloc((void**)&d_nu, sizeof(int) * m));
    CU(cudaMalloc((void**)&d_ne, sizeof(int) * m));
    CU(cudaMalloc((void**)&d_st, sizeof(int) * m));
    CU(cudaMemcpy(d_v, var_deltas, sizeof(double) * m, cudaMemcpyHostToDevice));
    if (uniforms) {
      CU(cudaMalloc((void**)&d_u, sizeof(double) * (size_t)m * width));
      CU(cudaMemcpy(d_u, uniforms, sizeof(double) * (size_t)m * width, cudaMemcpyHostToDevice));
    }
    launch_ars_standalone(kind, m, d_v, K, alpha, log_aw, d_u, width, seed, step, d_x, d_nu, d_ne, d_st, 0);
    CU(cudaDeviceSynchronize());
    CU(cudaGetLastError());
    CU(cudaMemcpy(x_out, d_x, sizeof(double) * m, cudaMemcpyDeviceToHost));
    if (n_unif) CU(cudaMemcpy(n_unif, d_nu, sizeof(int) * m, cudaMemcpyDeviceToHost));
    if (n_eval) CU(cudaMemcpy(n_eval, d_ne, sizeof(int) * m, cudaMemcpyDeviceToHost));
    if (status) CU(cudaMemcpy(status, d_st, sizeof(int) * m, cudaMemcpyDeviceToHost));
    cudaFree(d_v); cudaFree(d_x); cudaFree(d_nu); cudaFree(d_ne); cudaFree(d_st);
    if (d_u) cudaFree(d_u);
  });
}

int htlr_cuda_rng_dump(int32_t device, uint64_t seed, uint32_t purpose, uint64_t step, int32_t n_a, int32_t n_b,
                       int32_t kind, double shape, double* out) {
  return guarded(nullptr, [&] {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(HTLR_E_CUDA, "no CUDA device visible");
    CU(cudaSetDevice(device));
    double* d = nullptr;
    const size_t tot = (size_t)n_a * n_b;
    CU(cudaMalloc((void**)&d, sizeof(double) * tot));
    launch_rng_dump(seed, purpose, step, n_a, n_b, kind, shape, d, 0);
    CU(cudaDeviceSynchronize());
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, d, sizeof(double) * tot, cudaMemcpyDeviceToHost));
    cudaFree(d);
  });
}

int htlr_cuda_summary(htlr_handle* h, int32_t first, int32_t last, int32_t thin, double cut, double* mean_deltas,
                      double* sdb, int32_t* nzero) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    CU(cudaSetDevice(h->cfg.device));
    const int H = htlr_cuda_hist_len(h);
    if (first < 0) {
      // htlr_sdb's default (R/mccoef.r:104-131): drop the warm-up slices and a further floor(0.2 * iter.rmc)
      first = H - h->cfg.iters_rmc + (int)std::floor(0.2 * h->cfg.iters_rmc);
      last = H - 1;
    }
    if (thin < 1 || first > last || last >= H) FAIL(HTLR_E_ARG, "bad slice range for the posterior summary");
    const Geom& g = h->g;
    const size_t nvar = g.nvar, K = g.K;
    if (!h->s_mean) {
      h->s_mean = h->dalloc<double>(nvar * K, false);
      h->s_sdb = h->dalloc<double>(nvar, false);
      h->s_flag = h->dalloc<int>(nvar, false);
    }
    double* d_mean = h->s_mean;
    double* d_sdb = h->s_sdb;
    int* d_flag = h->s_flag;
    launch_trace_summary(g, h->b, first, last, thin, cut, d_mean, d_sdb, d_flag, h->st);
    h->launches += 3;
    CU(cudaStreamSynchronize(h->st));
    CU(cudaGetLastError());
    if (mean_deltas) CU(cudaMemcpy(mean_deltas, d_mean, nvar * K * sizeof(double), cudaMemcpyDeviceToHost));
    if (sdb) CU(cudaMemcpy(sdb, d_sdb, (nvar - 1) * sizeof(double), cudaMemcpyDeviceToHost));
    if (nzero) CU(cudaMemcpy(nzero, d_flag, (nvar - 1) * sizeof(int), cudaMemcpyDeviceToHost));
  });
}

int htlr_cuda_predict(htlr_handle* h, const double* X_ts, int64_t ldx, int32_t n_ts, int32_t first, int32_t last,
                      int32_t thin, double* probs) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    CU(cudaSetDevice(h->cfg.device));
    const int H = htlr_cuda_hist_len(h);
    if (first < 0) {
      // htlr_predict's default (R/predict.R:43-49 with rep.legacy = TRUE): warm-up slices and a further
      // floor(0.1 * iter.rmc) dropped, the first kept slice being the last dropped one (ignore.first = FALSE)
      first = std::max(0, H - h->cfg.iters_rmc + (int)std::floor(0.1 * h->cfg.iters_rmc) - 1);
      last = H - 1;
    }
    if (!X_ts || !probs || n_ts < 1 || ldx < n_ts) FAIL(HTLR_E_ARG, "bad test matrix");
    if (thin < 1 || first > last || last >= H) FAIL(HTLR_E_ARG, "bad slice range for prediction");
    if (h->sharded) FAIL(HTLR_E_STATE, "prediction is per GPU: call it on one rank (the trace is replicated)");
    const Geom& g = h->g;
    const size_t nvar = g.nvar, K = g.K;
    const int S = (last - first) / thin + 1;
    // scratch lives outside the handle's arena: test sets come and go
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)S, ((size_t)256 << 20) / (sizeof(double) * n_ts * K)));
    double *dA = nullptr, *dLV = nullptr, *dAcc = nullptr, *dP = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&] {
      cudaFree(dA); cudaFree(dLV); cudaFree(dAcc); cudaFree(dP);
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
    };
    const int64_t lda = round_up(n_ts, 16);   // columns 128-byte aligned, pad rows zero (16-byte async copies)
    try {
      CU(cudaMalloc((void**)&dA, sizeof(double) * (size_t)lda * nvar));
      if (lda != n_ts) CU(cudaMemsetAsync(dA, 0, sizeof(double) * (size_t)lda * nvar, h->st));
      CU(cudaMalloc((void**)&dLV, sizeof(double) * (size_t)n_ts * K * chunk));
      CU(cudaMalloc((void**)&dAcc, sizeof(double) * (size_t)n_ts * K));
      CU(cudaMalloc((void**)&dP, sizeof(double) * (size_t)n_ts * (K + 1)));
      CU(cudaMemcpy2DAsync(dA, (size_t)lda * sizeof(double), X_ts, (size_t)ldx * sizeof(double),
                           (size_t)n_ts * sizeof(double), nvar, cudaMemcpyHostToDevice, h->st));
      CU(cudaEventCreate(&e0));
      CU(cudaEventCreate(&e1));
      CU(cudaEventRecord(e0, h->st));
      h->launches += launch_predict(g, h->b, n_ts, dA, lda, first, thin, S, dLV, chunk, dAcc, dP, h->st);
      CU(cudaEventRecord(e1, h->st));
      CU(cudaStreamSynchronize(h->st));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, e0, e1));
      h->last_predict_ms = ms;
      h->last_predict_flops = 2.0 * n_ts * (double)nvar * (double)K * S;
      CU(cudaGetLastError());
      CU(cudaMemcpy(probs, dP, sizeof(double) * (size_t)n_ts * (K + 1), cudaMemcpyDeviceToHost));
    } catch (...) {
      cleanup();
      throw;
    }
    cleanup();
  });
}

int htlr_cuda_std(int32_t device, int32_t n, int32_t p, const double* A, int64_t lda, int32_t on_device,
                  double* A_std, int64_t ldo, double* median, double* sd) {
  return guarded(nullptr, [&] {
    if (n < 2 || p < 1 || !A || lda < n || (A_std && ldo < n)) FAIL(HTLR_E_ARG, "bad matrix for std");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      FAIL(HTLR_E_CUDA, "no CUDA device visible: this library has no CPU fallback");
    CU(cudaSetDevice(device));
    const DevInfo& dev = dev_info(device);
    double *dA = nullptr, *dO = nullptr, *dm = nullptr, *ds = nullptr;
    auto cleanup = [&] { if (!on_device) { cudaFree(dA); cudaFree(dO); } cudaFree(dm); cudaFree(ds); };
    try {
      CU(cudaMalloc((void**)&dm, sizeof(double) * p));
      CU(cudaMalloc((void**)&ds, sizeof(double) * p));
      const double* src = A;
      double* dst = A_std;
      int64_t ls = lda, ld = ldo;
      if (!on_device) {
        CU(cudaMalloc((void**)&dA, sizeof(double) * (size_t)n * p));
        CU(cudaMemcpy2D(dA, (size_t)n * sizeof(double), A, (size_t)lda * sizeof(double), (size_t)n * sizeof(double), p,
                        cudaMemcpyHostToDevice));
        src = dA; ls = n;
        if (A_std) { CU(cudaMalloc((void**)&dO, sizeof(double) * (size_t)n * p)); dst = dO; ld = n; }
      }
      launch_std(n, p, src, ls, dst, ld, dm, ds, dev.sm_count, 0);
      CU(cudaDeviceSynchronize());
      CU(cudaGetLastError());
      if (!on_device && A_std)
        CU(cudaMemcpy2D(A_std, (size_t)ldo * sizeof(double), dO, (size_t)n * sizeof(double), (size_t)n * sizeof(double), p,
                        cudaMemcpyDeviceToHost));
      if (median) CU(cudaMemcpy(median, dm, sizeof(double) * p, cudaMemcpyDeviceToHost));
      if (sd) CU(cudaMemcpy(sd, ds, sizeof(double) * p, cudaMemcpyDeviceToHost));
    } catch (...) {
      cleanup();
      throw;
    }
    cleanup();
  });
}

int htlr_cuda_gendata_mlr(int32_t device, int32_t n, int32_t p, int32_t NC, double nu, double w, uint64_t seed,
                          int64_t row0, double* X, int64_t ldx, double* ymat, int64_t* ybase, double* deltas_true) {
  return guarded(nullptr, [&] {
    if (n < 1 || p < 1 || NC < 2 || NC > kMaxK + 1 || !X || !ymat || !ybase || ldx < n || !(nu > 0) || !(w > 0))
      FAIL(HTLR_E_ARG, "bad arguments for gendata_mlr (2 <= NC <= 17, ldx >= n, device pointers)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      FAIL(HTLR_E_CUDA, "no CUDA device visible: this library has no CPU fallback");
    CU(cudaSetDevice(device));
    double *betas = nullptr, *part = nullptr;
    auto cleanup = [&] { cudaFree(betas); cudaFree(part); };
    try {
      CU(cudaMalloc((void**)&betas, sizeof(double) * (size_t)(p + 1) * NC));
      CU(cudaMalloc((void**)&part, sizeof(double) * (size_t)gen_mlr_chunks(p) * NC * n));
      launch_gen_mlr(n, p, NC, nu, w, seed, row0, X, ldx, ymat, reinterpret_cast<long long*>(ybase), betas, part, 0);
      CU(cudaDeviceSynchronize());
      CU(cudaGetLastError());
      if (deltas_true) {   // deltas = betas[, -1] - betas[, 1]
        std::vector<double> bh((size_t)(p + 1) * NC);
        CU(cudaMemcpy(bh.data(), betas, bh.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int k = 0; k < NC - 1; ++k)
          for (int j = 0; j <= p; ++j)
            deltas_true[(size_t)k * (p + 1) + j] = bh[(size_t)(k + 1) * (p + 1) + j] - bh[j];
      }
    } catch (...) {
      cleanup();
      throw;
    }
    cleanup();
  });
}

int htlr_cuda_gendata_fam(int32_t device, int32_t n, int32_t p, int32_t NC, const double* muj, int32_t pb, int32_t kb,
                          const double* Ablock, double sd_g, int32_t stdx, uint64_t seed, int64_t row0, double* X,
                          int64_t ldx, double* ymat, int64_t* ybase) {
  return guarded(nullptr, [&] {
    if (n < 1 || p < 1 || NC < 2 || NC > kMaxK + 1 || !muj || !X || !ymat || !ybase || ldx < n || pb < 0 || pb > p ||
        kb < 0 || kb > pb || (pb > 0 && !Ablock) || !(sd_g >= 0))
      FAIL(HTLR_E_ARG, "bad arguments for gendata_fam (2 <= NC <= 17, 0 <= kb <= pb <= p, ldx >= n, device pointers)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      FAIL(HTLR_E_CUDA, "no CUDA device visible: this library has no CPU fallback");
    CU(cudaSetDevice(device));
    double *dmu = nullptr, *dA = nullptr;
    auto cleanup = [&] { cudaFree(dmu); cudaFree(dA); };
    try {
      CU(cudaMalloc((void**)&dmu, sizeof(double) * (size_t)p * NC));
      CU(cudaMemcpy(dmu, muj, sizeof(double) * (size_t)p * NC, cudaMemcpyHostToDevice));
      CU(cudaMalloc((void**)&dA, sizeof(double) * (size_t)std::max(1, pb * kb)));
      if (pb * kb > 0) CU(cudaMemcpy(dA, Ablock, sizeof(double) * (size_t)pb * kb, cudaMemcpyHostToDevice));
      launch_gen_fam(n, p, NC, pb, kb, dmu, dA, sd_g, stdx, seed, row0, X, ldx, ymat, reinterpret_cast<long long*>(ybase), 0);
      CU(cudaDeviceSynchronize());
      CU(cudaGetLastError());
    } catch (...) {
      cleanup();
      throw;
    }
    cleanup();
  });
}

void* htlr_cuda_stream(htlr_handle* h) { return h ? (void*)h->st : nullptr; }

int htlr_cuda_profile_enable(htlr_handle* h, int32_t on) {
  if (!h) return HTLR_E_ARG;
  return guarded(h, [&] {
    CU(cudaSetDevice(h->cfg.device));
    prof_drain(h);
    h->profiling = on != 0;
  });
}

int htlr_cuda_profile_get(htlr_handle* h, htlr_profile* out, int32_t reset) {
  if (!h || !out) return HTLR_E_ARG;
  return guarded(h, [&] {
    CU(cudaSetDevice(h->cfg.device));
    prof_drain(h);
    CU(cudaStreamSynchronize(h->st));
    Ctl c;
    CU(cudaMemcpy(&c, h->b.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    h->prof.kernel_launches = h->launches;
    h->prof.sweep_bytes = 8.0 * h->g.n * (double)(c.cols_swept - h->cols_base);
    int khz = 0;
    CU(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->cfg.device));
    const double per_ms = khz > 0 ? (double)khz : 1.0;  // cycles per millisecond at the boost clock
    h->prof.fused_sweep_ms = (double)(c.ph_cycles[0] - h->ph_base[0]) / per_ms;
    h->prof.fused_barrier_ms = (double)(c.ph_cycles[1] - h->ph_base[1] + c.ph_cycles[3] - h->ph_base[3]) / per_ms;
    h->prof.fused_rowphase_ms = (double)(c.ph_cycles[2] - h->ph_base[2]) / per_ms;
    h->prof.small_kernel_trajectories = (double)(c.small_runs - h->small_base);
    for (int q = 0; q < 12; ++q) h->prof.timeline_us[q] = (double)(c.tl_acc[q] - h->tl_base[q]) * 1e-3;
    *out = h->prof;
    if (reset) {
      h->prof = htlr_profile{};
      h->cols_base = c.cols_swept;
      for (int q = 0; q < 4; ++q) h->ph_base[q] = c.ph_cycles[q];
      h->small_base = c.small_runs;
      for (int q = 0; q < 12; ++q) h->tl_base[q] = c.tl_acc[q];
    }
  });
}

int64_t htlr_cuda_launch_count(const htlr_handle* h) { return h ? h->launches : 0; }

int htlr_cuda_predict_stats(const htlr_handle* h, double* kernel_ms, double* flops) {
  if (!h) return HTLR_E_ARG;
  if (kernel_ms) *kernel_ms = h->last_predict_ms;
  if (flops) *flops = h->last_predict_flops;
  return HTLR_OK;
}

int htlr_cuda_fp64_peak(int32_t device, double* dfma_tflops, double* dmma_tflops) {
  return guarded(nullptr, [&] {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      FAIL(HTLR_E_CUDA, "no CUDA device visible: this library has no CPU fallback");
    CU(cudaSetDevice(device));
    const DevInfo& dev = dev_info(device);
    double* sink = nullptr;
    CU(cudaMalloc((void**)&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
      launch_fp64_peak(dev.sm_count, 200, mode, sink, 0);   // warm-up
      CU(cudaEventRecord(e0, 0));
      launch_fp64_peak(dev.sm_count, iters, mode, sink, 0);
      CU(cudaEventRecord(e1, 0));
      CU(cudaEventSynchronize(e1));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, e0, e1));
      const double threads = (double)dev.sm_count * 8 * 256;
      // per thread and round: 16 FMAs (2 flops each) / 8 warp-wide DMMAs (8x8x4 FMAs per warp = 16 flops per lane)
      const double flops = threads * iters * (mode ? 8 * 16.0 : 16 * 2.0);
      double* out = mode ? dmma_tflops : dfma_tflops;
      if (out) *out = flops / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    CU(cudaGetLastError());
  });
}

int64_t htlr_cuda_ars_table_full(htlr_handle* h) {
  if (!h) return -1;
  int64_t out = -1;
  guarded(h, [&] {
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->st));
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, reinterpret_cast<char*>(h->b.ctl) + offsetof(Ctl, ars_full), sizeof(v), cudaMemcpyDeviceToHost));
    out = (int64_t)v;
  });
  return out;
}

int htlr_cuda_fit(const htlr_cfg* cfg, const double* X, int64_t ldx, const double* ymat, const int64_t* ybase,
                  const double* deltas0, const double* sigmasbt0, const htlr_inject* inject, htlr_tick_fn tick,
                  void* user, double* mcdeltas, double* mcsigmasbt, double* mcvardeltas, double* mclogw,
                  double* mcloglike, double* mcuvar, double* mchmcrej, double* ddnloglike, char* errbuf,
                  size_t errbuf_len) {
  htlr_handle* h = nullptr;
  auto fail = [&](int rc, const char* msg) {
    if (errbuf && errbuf_len) snprintf(errbuf, errbuf_len, "%s", msg);
    if (h) htlr_cuda_destroy(h);
    return rc;
  };
  int rc = htlr_cuda_create(cfg, X, ldx, ymat, ybase, deltas0, sigmasbt0, &h);
  if (rc) return fail(rc, htlr_cuda_last_error(nullptr));
  const int total = cfg->iters_h + cfg->iters_rmc;
  if (cfg->rng_mode == HTLR_RNG_INJECT) {
    // injected streams are positional: run everything in one call
    std::vector<htlr_iter_stats> st(total);
    rc = htlr_cuda_run(h, total, inject, st.data());
    if (rc) return fail(rc, htlr_cuda_last_error(h));
    if (tick && tick(user, total, st.data(), total)) return fail(HTLR_E_STATE, "interrupted");
  } else {
    // Pipelined: kUnit iterations per enqueue call, up to kDepth iterations in flight; the trace slices and stats of
    // a unit are copied out while the following units run, so only the last unit's download is exposed. The tick
    // (progress line + interrupt poll; the reference polls every 256 iterations, gibbs.cpp:124) runs per unit.
    constexpr int kUnit = 8, kDepth = 64;
    std::vector<htlr_iter_stats> st(kUnit);
    int enq = 0, done = 0;
    while (done < total) {
      while (enq < total && enq - done < kDepth) {
        const int nq = std::min(kUnit, total - enq);
        rc = htlr_cuda_run_async(h, nq);
        if (rc) return fail(rc, htlr_cuda_last_error(h));
        enq += nq;
      }
      const int upto = std::min(done + kUnit, total);
      rc = htlr_cuda_collect(h, upto, done, st.data(), mcdeltas, mcsigmasbt, mcvardeltas, mclogw, mcloglike, mcuvar,
                             mchmcrej);
      if (rc) return fail(rc, htlr_cuda_last_error(h));
      const int nd = upto - done;
      done = upto;
      if (tick && tick(user, done, st.data(), nd)) return fail(HTLR_E_STATE, "interrupted");
    }
    mcdeltas = mcsigmasbt = mcvardeltas = mclogw = mcloglike = mcuvar = mchmcrej = nullptr;   // already collected
  }
  rc = htlr_cuda_fetch(h, mcdeltas, mcsigmasbt, mcvardeltas, mclogw, mcloglike, mcuvar, mchmcrej, ddnloglike);
  if (rc) return fail(rc, htlr_cuda_last_error(h));
  htlr_cuda_destroy(h);
  return HTLR_OK;
}

int htlr_cuda_nccl_unique_id(void* id128) {
  return guarded(nullptr, [&] {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    NC(nccl_need().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
  });
}

int htlr_cuda_nccl_init(const void* id128, int32_t rank, int32_t world, int32_t device, void** comm) {
  return guarded(nullptr, [&] {
    CU(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t c;
    NC(nccl_need().CommInitRank(&c, world, id, rank));
    *comm = c;
  });
}

void htlr_cuda_nccl_destroy(void* comm) {
  if (comm && nccl_api().ok) nccl_api().CommDestroy(static_cast<ncclComm_t>(comm));
}

}  // extern "C"
