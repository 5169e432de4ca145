// microbench.cu -- fast A/B of engine libraries on the GPU box without Python: loads one or more libb200nb*.so with
// dlopen, builds a config-2-like workload on the host (makeExampleDESeqDataSet law, ~condition design), keeps it
// resident on the device in the engine's gene-major layout and times the device entry points with CUDA events.
//
//   nvcc -O2 -std=c++17 -o scripts/microbench scripts/microbench.cu -ldl
//   scripts/microbench [--genes 50000] [--samples 100] [--reps 20] deseq2_b200/libb200nb.so deseq2_b200/libb200nb_exp_*.so
//
// One line per library: fitDisp (MLE start, no prior) and fitBeta kernel times (median of --reps after 3 warm-ups) and the
// largest difference of its log-dispersions / coefficients from the FIRST library on the command line (the experiments
// change rounding, not results: differences should be ~1e-12).  Written at the end of round 1 without GPU access
// (compiled, not run); scripts/ab_experiments.sh remains the reference procedure (bench.py contract numbers).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <random>
#include <string>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

typedef int (*fit_disp_dev_t)(const void*, int, const double*, const double*, const double*, const double*, double, double,
                              double, double, int, int, const double*, int, double, int, int, int, int, long long, double*,
                              int32_t*, int32_t*, double*, double*, double*, double*, double*, double*, void*);
typedef int (*fit_beta_dev_t)(const void*, int, const double*, const double*, int, const double*, const double*,
                              const double*, const double*, const double*, int, double, int, int, double, int, int, int,
                              long long, double*, double*, double*, double*, double*, double*, double*, double*, void*);
typedef const char* (*last_error_t)(void);
typedef int (*fit_disp_host_t)(const void*, int, const double*, const double*, const double*, const double*, double, double,
                               double, double, int, int, const double*, int, double, int, int, int, int, double*, int32_t*,
                               int32_t*, double*, double*, double*, double*, double*, double*);
typedef int (*fit_beta_host_t)(const void*, int, const double*, const double*, const double*, const double*, const double*,
                               const double*, const double*, int, double, int, int, double, int, int, int, double*, double*,
                               double*, double*, double*, double*, double*, double*);

template <typename T>
static T* to_device(const std::vector<T>& h) {
  T* d = nullptr;
  CK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(h.size(), 1)));
  CK(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return d;
}

static double median(std::vector<float> v) {
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}

int main(int argc, char** argv) {
  int n = 50000, m = 100, reps = 20;
  bool host_mode = false;   // --host: time the HOST entry points (R-layout pageable buffers, fresh result buffers each call)
  std::vector<std::string> libs;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--host")) { host_mode = true; continue; }
    if (!strcmp(argv[i], "--genes") && i + 1 < argc) n = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--samples") && i + 1 < argc) m = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--reps") && i + 1 < argc) reps = atoi(argv[++i]);
    else libs.push_back(argv[i]);
  }
  if (libs.empty()) {
    fprintf(stderr, "usage: %s [--genes N] [--samples M] [--reps R] lib1.so [lib2.so ...]\n", argv[0]);
    return 2;
  }
  const int p = 2;
  const long long ld = (m + 3) & ~3;

  // ---- workload: beta0 ~ N(4, 2) (log2), beta1 ~ N(0, 1), alpha = 4 / 2^beta0 + 0.1, size factors ~ exp(N(0, 0.25))
  std::mt19937_64 rng(20260923);
  std::normal_distribution<double> gauss(0.0, 1.0);
  std::vector<double> sf(m), x((size_t)m * p), xt((size_t)p * m);
  double lsum = 0.0;
  for (int j = 0; j < m; j++) { sf[j] = exp(0.25 * gauss(rng)); lsum += log(sf[j]); }
  for (int j = 0; j < m; j++) sf[j] /= exp(lsum / m);
  for (int j = 0; j < m; j++) {
    x[j] = 1.0;
    x[(size_t)m + j] = (j >= m / 2) ? 1.0 : 0.0;       // column-major m x p
  }
  std::vector<int32_t> y((size_t)n * ld, 0);
  std::vector<double> mu((size_t)n * ld, 1.0), la0(n), alpha(n), beta0((size_t)p * n, 0.0);
  for (int i = 0; i < n; i++) {
    const double b0 = 4.0 + 2.0 * gauss(rng), b1 = gauss(rng);
    const double a = 4.0 / pow(2.0, b0) + 0.1;
    double sum[2] = {0.0, 0.0};
    int cnt[2] = {0, 0};
    for (int j = 0; j < m; j++) {
      const int g = j >= m / 2;
      const double mean = sf[j] * pow(2.0, b0 + b1 * g);
      std::gamma_distribution<double> gam(1.0 / a, a * mean);
      std::poisson_distribution<long long> pois(std::max(gam(rng), 1e-300));
      const long long v = std::min<long long>(pois(rng), 2000000000LL);
      y[(size_t)i * ld + j] = (int32_t)v;
      sum[g] += (double)v / sf[j];
      cnt[g]++;
    }
    double all = 0.0, var = 0.0;
    for (int j = 0; j < m; j++) all += y[(size_t)i * ld + j] / sf[j];
    all /= m;
    for (int j = 0; j < m; j++) { const double d = y[(size_t)i * ld + j] / sf[j] - all; var += d * d; }
    var /= (m - 1);
    for (int j = 0; j < m; j++) {
      const int g = j >= m / 2;
      mu[(size_t)i * ld + j] = std::max(sf[j] * sum[g] / cnt[g], 0.5);     // linear-model mu of the two-group design
    }
    const double mom = all > 0 ? (var - all) / (all * all) : 1e-8;        // moments estimate (unit size-factor mean)
    alpha[i] = a;
    la0[i] = log(std::min(std::max(mom, 1e-8), std::max(10.0, (double)m)));
    beta0[i] = log(std::max(all, 0.1));                                     // (p, n) layout = column-major n x p
  }
  // keep only genes with a non-zero sum? all-zero rows are legal input for the kernels and rare here: keep them
  void* d_y = to_device(y);
  double* d_mu = to_device(mu);
  double* d_x = to_device(x);
  double* d_la0 = to_device(la0);
  double* d_alpha = to_device(alpha);
  double* d_beta0 = to_device(beta0);
  double* d_sf = to_device(sf);
  std::vector<double> contrast = {1.0, 0.0}, lam = {1e-6 / (M_LN2 * M_LN2), 1e-6 / (M_LN2 * M_LN2)};
  double* d_contrast = to_device(contrast);
  double* d_lam = to_device(lam);
  double *d_out[9], *d_bout, *d_bvar, *d_bsc[4], *d_h, *d_muo;
  int32_t *d_it, *d_ita;
  for (auto& q : d_out) CK(cudaMalloc(&q, sizeof(double) * n));
  for (auto& q : d_bsc) CK(cudaMalloc(&q, sizeof(double) * n));
  CK(cudaMalloc(&d_it, sizeof(int32_t) * n));
  CK(cudaMalloc(&d_ita, sizeof(int32_t) * n));
  CK(cudaMalloc(&d_bout, sizeof(double) * n * p));
  CK(cudaMalloc(&d_bvar, sizeof(double) * n * p));
  CK(cudaMalloc(&d_h, sizeof(double) * n * ld));
  CK(cudaMalloc(&d_muo, sizeof(double) * n * ld));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));

  if (host_mode) {
    // R layout: column-major n x m; the host entry points of ONE library (knobs are read from the environment once per
    // process, so run the binary once per setting); a step = fitDisp + fitDisp + fitBeta like bench.py's e2e
    std::vector<int32_t> yc((size_t)n * m);
    std::vector<double> muc((size_t)n * m), nfc((size_t)n * m), b0c((size_t)n * p, 0.0);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < m; j++) {
        yc[(size_t)j * n + i] = y[(size_t)i * ld + j];
        muc[(size_t)j * n + i] = mu[(size_t)i * ld + j];
        nfc[(size_t)j * n + i] = sf[j];
      }
    for (int i = 0; i < n; i++) b0c[i] = beta0[i];
    void* h = dlopen(libs[0].c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) { printf("%s: dlopen failed: %s\n", libs[0].c_str(), dlerror()); return 1; }
    auto fdh = (fit_disp_host_t)dlsym(h, "b200nb_fit_disp");
    auto fbh = (fit_beta_host_t)dlsym(h, "b200nb_fit_beta");
    auto le = (last_error_t)dlsym(h, "b200nb_last_error");
    auto cache_clear = (void (*)(void))dlsym(h, "b200nb_cache_clear");   // a step = a new DESeq() run: nothing may be reused
    std::vector<double> tstep, t1, t3;
    for (int r = 0; r < reps + 2; r++) {
      // fresh result buffers every call, like R's allocVector: their pages are first touched by the library
      double* o[14];
      for (auto& q : o) q = (double*)malloc(sizeof(double) * n);
      int32_t* it = (int32_t*)malloc(sizeof(int32_t) * n);
      int32_t* ita = (int32_t*)malloc(sizeof(int32_t) * n);
      double* H = (double*)malloc(sizeof(double) * (size_t)n * m);
      double* bo = (double*)malloc(sizeof(double) * (size_t)n * p);
      double* bv = (double*)malloc(sizeof(double) * (size_t)n * p);
      double* sc4[4];
      for (auto& q : sc4) q = (double*)malloc(sizeof(double) * n);
      if (cache_clear) cache_clear();
      const auto c0 = std::chrono::steady_clock::now();
      int rc = fdh(yc.data(), 0, x.data(), muc.data(), la0.data(), la0.data(), 1.0, log(1e-8 / 10), 1.0, 1e-6, 100, 0, nullptr, 0,
                   1e-2, 1, n, m, p, o[0], it, ita, o[1], o[2], o[3], o[4], o[5], o[6]);
      const auto c1 = std::chrono::steady_clock::now();
      rc |= fdh(yc.data(), 0, x.data(), muc.data(), o[0], la0.data(), 1.0, log(1e-8 / 10), 1.0, 1e-6, 100, 1, nullptr, 0, 1e-2, 1,
                n, m, p, o[7], it, ita, o[8], o[9], o[10], o[11], o[12], o[13]);
      const auto c2 = std::chrono::steady_clock::now();
      rc |= fbh(yc.data(), 0, x.data(), nfc.data(), alpha.data(), contrast.data(), b0c.data(), lam.data(), nullptr, 0, 1e-8, 100, 1,
                0.5, n, m, p, bo, bv, sc4[0], H, sc4[1], sc4[2], sc4[3], nullptr);
      const auto c3 = std::chrono::steady_clock::now();
      if (rc) { printf("host call failed: %s\n", le()); return 1; }
      if (r >= 2) {
        t1.push_back(std::chrono::duration<double, std::milli>(c1 - c0).count());
        t3.push_back(std::chrono::duration<double, std::milli>(c3 - c2).count());
        tstep.push_back(std::chrono::duration<double, std::milli>(c3 - c0).count());
      }
      for (auto q : o) free(q);
      for (auto q : sc4) free(q);
      free(it); free(ita); free(H); free(bo); free(bv);
    }
    std::sort(tstep.begin(), tstep.end()); std::sort(t1.begin(), t1.end()); std::sort(t3.begin(), t3.end());
    printf("host path %s: step %.2f ms (fitDisp %.2f, fitBeta %.2f; medians of %d) = %.2f M genes/s end to end\n", libs[0].c_str(),
           tstep[tstep.size() / 2], t1[t1.size() / 2], t3[t3.size() / 2], reps, n / tstep[tstep.size() / 2] / 1e3);
    return 0;
  }

  std::vector<double> ref_la, ref_beta;
  printf("workload: %d genes x %d samples, p = %d (gene-major, resident); %d timed repetitions per kernel\n", n, m, p, reps);
  for (const std::string& path : libs) {
    void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) { printf("%s: dlopen failed: %s\n", path.c_str(), dlerror()); continue; }
    auto fd = (fit_disp_dev_t)dlsym(h, "b200nb_fit_disp_dev");
    auto fb = (fit_beta_dev_t)dlsym(h, "b200nb_fit_beta_dev");
    auto le = (last_error_t)dlsym(h, "b200nb_last_error");
    if (!fd || !fb || !le) { printf("%s: missing symbols\n", path.c_str()); continue; }
    auto run_disp = [&]() {
      return fd(d_y, 0, d_x, d_mu, d_la0, d_la0, 1.0, log(1e-8 / 10), 1.0, 1e-6, 100, 0, nullptr, 0, 1e-2, 1, n, m, p, ld,
                d_out[0], d_it, d_ita, d_out[1], d_out[2], d_out[3], d_out[4], d_out[5], d_out[6], st);
    };
    auto run_beta = [&]() {
      return fb(d_y, 0, d_x, d_sf, 1, d_alpha, d_contrast, d_beta0, d_lam, nullptr, 0, 1e-8, 100, 1, 0.5, n, m, p, ld, d_bout,
                d_bvar, d_bsc[0], d_h, d_bsc[1], d_bsc[2], d_bsc[3], d_muo, st);
    };
    bool ok = true;
    for (int w = 0; w < 3 && ok; w++) ok = !run_disp() && !run_beta();
    if (!ok) { printf("%s: call failed: %s\n", path.c_str(), le()); continue; }
    CK(cudaStreamSynchronize(st));
    std::vector<float> td, tb;
    for (int r = 0; r < reps; r++) {
      float ms;
      CK(cudaEventRecord(e0, st));
      run_disp();
      CK(cudaEventRecord(e1, st));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      td.push_back(ms);
      CK(cudaEventRecord(e0, st));
      run_beta();
      CK(cudaEventRecord(e1, st));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      tb.push_back(ms);
    }
    std::vector<double> la(n), bt((size_t)n * p);
    CK(cudaMemcpy(la.data(), d_out[0], sizeof(double) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bt.data(), d_bout, sizeof(double) * n * p, cudaMemcpyDeviceToHost));
    double dla = 0.0, dbt = 0.0;
    if (ref_la.empty()) { ref_la = la; ref_beta = bt; }
    for (int i = 0; i < n; i++)
      if (isfinite(la[i]) && isfinite(ref_la[i])) dla = std::max(dla, fabs(la[i] - ref_la[i]));
    for (size_t i = 0; i < bt.size(); i++)
      if (isfinite(bt[i]) && isfinite(ref_beta[i]) && fabs(ref_beta[i]) < 25) dbt = std::max(dbt, fabs(bt[i] - ref_beta[i]));
    const double md = median(td), mb = median(tb);
    printf("%-58s fitDisp %.3f ms  fitBeta %.3f ms  (%.1f M genes/s for 2 x fitDisp + fitBeta)  max|dlog_alpha| %.2e  max|dbeta| %.2e\n",
           path.c_str(), md, mb, n / (2 * md + mb) / 1e3, dla, dbt);
    // the library keeps device arenas alive; leave it loaded (dlclose would not free them anyway)
  }
  return 0;
}
