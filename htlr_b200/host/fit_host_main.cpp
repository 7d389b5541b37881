// Stand-alone C++ host for the CUDA sampler: what the Rcpp glue does, without R.
//   fit_host <problem.bin> <result.bin>
// Reads the 22 htlr_fit_helper arguments (+ device options, optional injected variates) from a flat
// binary file, runs htlr::htlr_fit_helper (cuda_fit.hpp) and writes the OutputR arrays. The GPU tests
// drive it to check that the C++ host layer and the ctypes mirror give identical chains.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cuda_fit.hpp"

namespace {

struct Reader {
  FILE* f;
  template <class T>
  T one() {
    T v;
    if (std::fread(&v, sizeof(T), 1, f) != 1) { std::fprintf(stderr, "short problem file\n"); std::exit(2); }
    return v;
  }
  template <class T>
  std::vector<T> many(size_t cnt) {
    std::vector<T> v(cnt);
    if (cnt && std::fread(v.data(), sizeof(T), cnt, f) != cnt) { std::fprintf(stderr, "short problem file\n"); std::exit(2); }
    return v;
  }
};

template <class T>
void put(FILE* f, const std::vector<T>& v) { std::fwrite(v.data(), sizeof(T), v.size(), f); }

}  // namespace

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: fit_host problem.bin result.bin\n"); return 2; }
  FILE* fi = std::fopen(argv[1], "rb");
  if (!fi) { std::perror(argv[1]); return 2; }
  Reader r{fi};
  if (r.one<uint32_t>() != 0x48544c52u) { std::fprintf(stderr, "bad magic\n"); return 2; }
  const int p = r.one<int32_t>(), K = r.one<int32_t>(), n = r.one<int32_t>();
  const int ptype_len = r.one<int32_t>();
  std::vector<char> pt = r.many<char>(ptype_len);
  const std::string ptype(pt.begin(), pt.end());
  const double alpha = r.one<double>(), s = r.one<double>(), eta = r.one<double>();
  const int iters_rmc = r.one<int32_t>(), iters_h = r.one<int32_t>(), thin = r.one<int32_t>();
  const int leap_L = r.one<int32_t>(), leap_L_h = r.one<int32_t>();
  const double leap_step = r.one<double>(), hmc_sgmcut = r.one<double>();
  const int keep = r.one<int32_t>(), silence = r.one<int32_t>(), legacy = r.one<int32_t>();
  htlr::DeviceOptions opt;
  opt.device = r.one<int32_t>();
  opt.algo = r.one<int32_t>();
  opt.seed = r.one<uint64_t>();
  opt.ars_inject_width = r.one<int32_t>();
  const int has_inject = r.one<int32_t>();
  const size_t nvar = (size_t)p + 1;
  auto X = r.many<double>((size_t)n * nvar);
  auto ymat = r.many<double>((size_t)n * K);
  auto ybase = r.many<int64_t>(n);
  auto deltas = r.many<double>(nvar * K);
  auto sigmasbt = r.many<double>(nvar);
  std::vector<double> inorm, iunif, igam, iars;
  htlr_inject inj{};
  if (has_inject) {
    const int64_t c0 = r.one<int64_t>(), c1 = r.one<int64_t>(), c2 = r.one<int64_t>(), c3 = r.one<int64_t>();
    inorm = r.many<double>(c0); iunif = r.many<double>(c1); igam = r.many<double>(c2);
    iars = r.many<double>((size_t)c3 * (size_t)(opt.ars_inject_width > 0 ? opt.ars_inject_width : 0));
    inj.normals = inorm.data(); inj.n_normals = c0;
    inj.uniforms = iunif.data(); inj.n_uniforms = c1;
    inj.gammas = igam.data(); inj.n_gammas = c2;
    inj.ars_uniforms = c3 ? iars.data() : nullptr; inj.n_ars_blocks = c3;
  }
  std::fclose(fi);
  try {
    htlr::FitOutput o = htlr::htlr_fit_helper(p, K, n, X.data(), ymat.data(), ybase.data(), ptype, alpha, s, eta,
                                              iters_rmc, iters_h, thin, leap_L, leap_L_h, leap_step, hmc_sgmcut,
                                              deltas.data(), sigmasbt.data(), keep != 0, silence, legacy != 0, opt,
                                              has_inject ? &inj : nullptr);
    FILE* fo = std::fopen(argv[2], "wb");
    if (!fo) { std::perror(argv[2]); return 2; }
    const int32_t H = o.hist_len;
    std::fwrite(&H, sizeof(H), 1, fo);
    std::fwrite(&o.sgmsq_cut, sizeof(double), 1, fo);
    put(fo, o.DDNloglike); put(fo, o.mcdeltas); put(fo, o.mclogw); put(fo, o.mcsigmasbt); put(fo, o.mcvardeltas);
    put(fo, o.mcloglike); put(fo, o.mcuvar); put(fo, o.mchmcrej);
    std::fclose(fo);
  } catch (const htlr::CudaFitError& e) {
    std::fprintf(stderr, "htlr_fit_helper failed (%d): %s\n", e.code, e.what());
    return 3;
  }
  return 0;
}
