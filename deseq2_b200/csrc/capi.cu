// capi.cu -- the C ABI of libb200nb.so (declared in include/b200nb.h).
// Host entry points: H2D copy of R-layout buffers -> device transpose to gene-major -> kernels ->
// (transpose back) -> D2H.  Device entry points: enqueue only.  No CPU fallback anywhere.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <math.h>
#include <stdlib.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/b200nb.h"
#include "engine.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// grow-only device workspace of the host entry points (one set per process; calls are serialised)
enum Slot {
  S_RAW0, S_RAW1, S_RAW2, S_Y, S_MU, S_W, S_NF, S_X, S_V0, S_V1, S_V2, S_OUTD, S_OUTI, S_H, S_MUO, S_HC, S_MUC,
  S_BETA_IN, S_BETA_OUT, S_BETA_VAR, S_NSLOTS
};
struct Workspace {
  void* p[S_NSLOTS] = {};
  size_t bytes[S_NSLOTS] = {};
  cudaStream_t stream = nullptr;
};

// ---------------------------------------------------------------- pinned staging for pageable host buffers
// R hands over ordinary (pageable) memory.  cudaMemcpy from pageable memory runs at ~9 GB/s on the GPU box;
// staging through a ring of pinned buffers filled by a few host threads while the previous chunk is in flight
// reaches PCIe speed.  B200NB_STAGE_THREADS (default 8; 4..16 measure the same) OpenMP workers do the host-side memcpy.
// bytes per pinned staging buffer: 16 MB as measured in round 1; B200NB_STAGE_CHUNK_MB (1..256) is a knob for the next
// GPU session (smaller chunks start the first DMA sooner, larger ones pay fewer event round trips)
size_t stage_chunk_bytes() {
  static const size_t v = [] {
    const char* e = getenv("B200NB_STAGE_CHUNK_MB");
    long mb = e ? atol(e) : 16;
    if (mb < 1 || mb > 256) mb = 16;
    return (size_t)mb << 20;
  }();
  return v;
}
#define kStageChunk (stage_chunk_bytes())   // kept as a name: the copy routines below read like the fixed-size originals
constexpr int kStageRing = 3;
struct Staging {
  void* buf[kStageRing] = {};
  cudaEvent_t ev[kStageRing] = {};
  bool ready = false;
};

// Host-side context of the host entry points: device workspace, stream and pinned staging ring.  Context 0 belongs
// to the calling thread; contexts 1.. are used by the extra worker threads of the gene-chunked path (opt-in, see
// run_chunked below), each with its own stream so that one worker's copies overlap another worker's kernels.
struct HostCtx {
  Workspace ws;
  Staging stage;
};
constexpr int kMaxWorkers = 4;
HostCtx g_ctx[kMaxWorkers];
thread_local HostCtx* t_ctx = &g_ctx[0];
std::mutex g_call_mu;    // host entry points are serialised (one set of contexts per process)
std::mutex g_arena_mu;   // guards the (re)allocation of the shared device arenas below
inline Workspace& cur_ws() { return t_ctx->ws; }        // the calling thread's context
inline Staging& cur_stage() { return t_ctx->stage; }

int ws_get(Slot s, size_t bytes, void** out) {
  if (bytes == 0) bytes = 16;
  if (cur_ws().bytes[s] < bytes) {
    if (cur_ws().p[s]) cudaFree(cur_ws().p[s]);
    cur_ws().p[s] = nullptr;
    cur_ws().bytes[s] = 0;
    size_t want = bytes + bytes / 8;
    CU(cudaMalloc(&cur_ws().p[s], want));
    cur_ws().bytes[s] = want;
  }
  *out = cur_ws().p[s];
  return 0;
}

int ws_stream(cudaStream_t* st) {
  if (!cur_ws().stream) CU(cudaStreamCreateWithFlags(&cur_ws().stream, cudaStreamNonBlocking));
  *st = cur_ws().stream;
  return 0;
}

// device-side work-queue counters: a ring of slots, one per launch, so launches in flight on different
// streams never share a counter (a slot is reused only after kRing further launches).
constexpr int kRing = 1024;
unsigned int* g_counters = nullptr;
std::atomic<unsigned int> g_counter_next{0};
int next_counter(unsigned int** out) {
  std::lock_guard<std::mutex> lk(g_arena_mu);
  if (!g_counters) CU(cudaMalloc(&g_counters, sizeof(unsigned int) * kRing));
  *out = g_counters + (g_counter_next.fetch_add(1) % kRing);
  return 0;
}

int stage_threads() {
  static int t = [] {
    const char* e = getenv("B200NB_STAGE_THREADS");
    int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 64 ? 64 : v);
  }();
  return t;
}
int stage_init() {
  if (cur_stage().ready) return 0;
  for (int i = 0; i < kStageRing; i++) {
    CU(cudaHostAlloc(&cur_stage().buf[i], kStageChunk, cudaHostAllocDefault));
    CU(cudaEventCreateWithFlags(&cur_stage().ev[i], cudaEventDisableTiming));
  }
  cur_stage().ready = true;
  return 0;
}

// Large result matrices (hat_diagonals, mu) land in memory the caller has just allocated: every 4 KB page of it is
// first touched by the D2H scatter, which was measured to run at ~6 GB/s for that reason.  Two opt-in knobs for the
// next GPU session (both off by default: not timed yet): B200NB_D2H_HUGEPAGE=1 asks the kernel for transparent huge
// pages on the destination range before it is touched (512x fewer faults); B200NB_D2H_THREADS=<T> uses a different
// number of host threads for the device-to-host scatter than for the host-to-device gather.
int d2h_threads() {
  static int t = [] {
    const char* e = getenv("B200NB_D2H_THREADS");
    int v = e ? atoi(e) : 0;
    return v < 1 ? stage_threads() : (v > 64 ? 64 : v);
  }();
  return t;
}
void advise_hugepages(void* p, size_t bytes) {
  static const bool on = [] {
    const char* e = getenv("B200NB_D2H_HUGEPAGE");
    return e && atoi(e) != 0;
  }();
  if (!on || bytes < (8u << 20)) return;
  const uintptr_t huge = (uintptr_t)2 << 20;
  const uintptr_t lo = ((uintptr_t)p + huge - 1) & ~(huge - 1), hi = ((uintptr_t)p + bytes) & ~(huge - 1);
  if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);   // a hint: failure is harmless
}

// B200NB_D2H_POPULATE=1 (opt-in, not yet timed on the GPU box): each copying thread first asks the kernel to populate
// the page tables of its destination slice in one call (MADV_POPULATE_WRITE, Linux >= 5.14) instead of taking one page
// fault per 4 KB while it copies; on the build container that took a 40 MB first-touch scatter from 3.7 to 2.7 ms.
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
bool d2h_populate() {
  static const bool on = [] {
    const char* e = getenv("B200NB_D2H_POPULATE");
    return e && atoi(e) != 0;
  }();
  return on;
}

void par_memcpy(void* dst, const void* src, size_t bytes, int T = 0, bool populate_dst = false) {
  if (bytes < (1u << 20)) {
    memcpy(dst, src, bytes);
    return;
  }
  if (T <= 0) T = stage_threads();
#pragma omp parallel for num_threads(T) schedule(static)
  for (int t = 0; t < T; t++) {
    const size_t lo = bytes * t / T, hi = bytes * (t + 1) / T;
    if (populate_dst) {
      const uintptr_t a = ((uintptr_t)dst + lo + 4095) & ~(uintptr_t)4095, b = ((uintptr_t)dst + hi) & ~(uintptr_t)4095;
      if (b > a) madvise(reinterpret_cast<void*>(a), b - a, MADV_POPULATE_WRITE);   // a hint: failure is harmless
    }
    memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, hi - lo);
  }
}

int h2d_staged(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes <= (256u << 10)) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  if (stage_init()) return 1;
  size_t off = 0;
  for (int k = 0; off < bytes; k++) {
    const int b = k % kStageRing;
    const size_t len = (bytes - off < kStageChunk) ? bytes - off : kStageChunk;
    if (k >= kStageRing) CU(cudaEventSynchronize(cur_stage().ev[b]));
    par_memcpy(cur_stage().buf[b], static_cast<const char*>(src) + off, len);
    CU(cudaMemcpyAsync(static_cast<char*>(dst) + off, cur_stage().buf[b], len, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(cur_stage().ev[b], st));
    off += len;
  }
  // the ring is reused by the next transfer: drain it
  for (int b = 0; b < kStageRing; b++) CU(cudaEventSynchronize(cur_stage().ev[b]));
  return 0;
}

// device -> pageable host through the pinned ring; synchronous on return
int d2h_staged(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes <= (256u << 10)) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
  }
  if (stage_init()) return 1;
  const size_t nchunk = (bytes + kStageChunk - 1) / kStageChunk;
  auto issue = [&](size_t k) -> int {
    const int b = (int)(k % kStageRing);
    const size_t off = k * kStageChunk;
    const size_t len = (bytes - off < kStageChunk) ? bytes - off : kStageChunk;
    CU(cudaMemcpyAsync(cur_stage().buf[b], static_cast<const char*>(src) + off, len, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(cur_stage().ev[b], st));
    return 0;
  };
  for (size_t k = 0; k < nchunk && k < (size_t)kStageRing - 1; k++)
    if (issue(k)) return 1;
  for (size_t k = 0; k < nchunk; k++) {
    if (k + kStageRing - 1 < nchunk && issue(k + kStageRing - 1)) return 1;
    const int b = (int)(k % kStageRing);
    const size_t off = k * kStageChunk;
    const size_t len = (bytes - off < kStageChunk) ? bytes - off : kStageChunk;
    CU(cudaEventSynchronize(cur_stage().ev[b]));
    par_memcpy(static_cast<char*>(dst) + off, cur_stage().buf[b], len, d2h_threads(), d2h_populate());
  }
  return 0;
}

// ---- row blocks of column-major host matrices (the gene-chunked path) ----------------------------------------
// `host` is a column-major matrix with n_total rows (R layout); the block is rows [g0, g0 + gc) of its m columns and
// lives on the device as a contiguous column-major gc x m matrix.  gc == n_total is the whole matrix: the contiguous
// routines above.  Otherwise whole column segments are gathered into / scattered from the pinned ring.
int h2d_block(void* dst, const void* host, size_t n_total, size_t g0, size_t gc, int m, int elem, cudaStream_t st) {
  if (gc == n_total) return h2d_staged(dst, host, gc * m * elem, st);
  const size_t col = gc * elem;
  const char* src = static_cast<const char*>(host);
  if (col * m <= (256u << 10) || col > kStageChunk) {
    for (int j = 0; j < m; j++) {
      const void* seg = src + ((size_t)j * n_total + g0) * elem;
      if (col > kStageChunk) {
        if (h2d_staged(static_cast<char*>(dst) + j * col, seg, col, st)) return 1;
      } else {
        CU(cudaMemcpyAsync(static_cast<char*>(dst) + j * col, seg, col, cudaMemcpyHostToDevice, st));
      }
    }
    return 0;
  }
  if (stage_init()) return 1;
  const int cpc = (int)(kStageChunk / col);   // whole columns per staging buffer (>= 1)
  const int T = stage_threads();
  int j = 0;
  for (int k = 0; j < m; k++) {
    const int b = k % kStageRing;
    const int nc = (m - j < cpc) ? m - j : cpc;
    if (k >= kStageRing) CU(cudaEventSynchronize(cur_stage().ev[b]));
    char* buf = static_cast<char*>(cur_stage().buf[b]);
#pragma omp parallel for num_threads(T) schedule(static) if (nc * col >= (1u << 20))
    for (int c = 0; c < nc; c++) memcpy(buf + (size_t)c * col, src + ((size_t)(j + c) * n_total + g0) * elem, col);
    CU(cudaMemcpyAsync(static_cast<char*>(dst) + (size_t)j * col, buf, (size_t)nc * col, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(cur_stage().ev[b], st));
    j += nc;
  }
  for (int b = 0; b < kStageRing; b++) CU(cudaEventSynchronize(cur_stage().ev[b]));
  return 0;
}

// device (contiguous column-major gc x m) -> rows [g0, g0 + gc) of the column-major host matrix; synchronous on return
int d2h_block(void* host, const void* src, size_t n_total, size_t g0, size_t gc, int m, int elem, cudaStream_t st) {
  if (gc == n_total) return d2h_staged(host, src, gc * m * elem, st);
  const size_t col = gc * elem;
  char* dst = static_cast<char*>(host);
  if (col * m <= (256u << 10) || col > kStageChunk) {
    for (int j = 0; j < m; j++) {
      void* seg = dst + ((size_t)j * n_total + g0) * elem;
      if (col > kStageChunk) {
        if (d2h_staged(seg, static_cast<const char*>(src) + j * col, col, st)) return 1;
      } else {
        CU(cudaMemcpyAsync(seg, static_cast<const char*>(src) + j * col, col, cudaMemcpyDeviceToHost, st));
      }
    }
    CU(cudaStreamSynchronize(st));
    return 0;
  }
  if (stage_init()) return 1;
  const int cpc = (int)(kStageChunk / col);
  const int T = d2h_threads();
  const int nbatch = (m + cpc - 1) / cpc;
  auto issue = [&](int k) -> int {
    const int b = k % kStageRing, j = k * cpc, nc = (m - j < cpc) ? m - j : cpc;
    CU(cudaMemcpyAsync(cur_stage().buf[b], static_cast<const char*>(src) + (size_t)j * col, (size_t)nc * col,
                       cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(cur_stage().ev[b], st));
    return 0;
  };
  for (int k = 0; k < nbatch && k < kStageRing - 1; k++)
    if (issue(k)) return 1;
  for (int k = 0; k < nbatch; k++) {
    if (k + kStageRing - 1 < nbatch && issue(k + kStageRing - 1)) return 1;
    const int b = k % kStageRing, j = k * cpc, nc = (m - j < cpc) ? m - j : cpc;
    CU(cudaEventSynchronize(cur_stage().ev[b]));
    const char* buf = static_cast<const char*>(cur_stage().buf[b]);
#pragma omp parallel for num_threads(T) schedule(static) if (nc * col >= (1u << 20))
    for (int c = 0; c < nc; c++) memcpy(dst + ((size_t)(j + c) * n_total + g0) * elem, buf + (size_t)c * col, col);
  }
  return 0;
}

// scratch buffers for fit_disp launches (work queue + per-mode gene lists): ONE device arena cut into a ring of
// equal slots, one slot per launch in flight (slot reuse after kScratchRing further launches; launches on one stream
// are ordered).  The arena is (re)allocated only when a call needs a larger slot than any before -- a per-slot
// cudaFree/cudaMalloc was measured to stall the pipeline by up to 80 ms whenever the gene count changed, and a
// lazily allocated ring stalled the first 8 steps of every run.
constexpr int kScratchRing = 16;
void* g_scratch_arena = nullptr;
size_t g_scratch_slot = 0;
std::atomic<unsigned int> g_scratch_next{0};
int next_scratch(size_t bytes, unsigned int** out) {
  std::lock_guard<std::mutex> lk(g_arena_mu);
  bytes = (bytes + 255) & ~(size_t)255;
  if (g_scratch_slot < bytes) {
    if (g_scratch_arena) {
      CU(cudaDeviceSynchronize());
      CU(cudaFree(g_scratch_arena));
    }
    g_scratch_arena = nullptr;
    g_scratch_slot = 0;
    const size_t slot = (bytes + bytes / 2 + 4095) & ~(size_t)4095;
    CU(cudaMalloc(&g_scratch_arena, slot * kScratchRing));
    g_scratch_slot = slot;
  }
  const unsigned int s = g_scratch_next.fetch_add(1) % kScratchRing;
  *out = reinterpret_cast<unsigned int*>(static_cast<char*>(g_scratch_arena) + (size_t)s * g_scratch_slot);
  return 0;
}

// ---------------------------------------------------------------- design analysis for the general-p kernels
// Distinct rows of the design matrix (what R/core.R:2450 modelMatrixGroups computes) and a row id per sample.
// <= 32 distinct rows: GROUPED (xg = G x ps); otherwise SAMPLEWISE (xg = all m rows, gid[j] = j).
bool use_generic(int p) {
  static const bool force = getenv("B200NB_FORCE_GENERIC") != nullptr;
  return force || p > nb::kMaxSmallP;
}

struct DesignDev {
  const double* xg;
  const int* gid;
  int G, grouped;
  int saturated;       // G == p and the distinct rows form an invertible p x p matrix
  double sat_logdet;   // 2 log|det X_g|
};
constexpr int kDesignRing = 16;
void* g_design_arena = nullptr;
size_t g_design_slot = 0;
std::atomic<unsigned int> g_design_next{0};
int next_design_slot(size_t need, char** out) {
  std::lock_guard<std::mutex> lk(g_arena_mu);
  need = (need + 255) & ~(size_t)255;
  if (g_design_slot < need) {
    if (g_design_arena) {
      CU(cudaDeviceSynchronize());
      CU(cudaFree(g_design_arena));
    }
    g_design_arena = nullptr;
    g_design_slot = 0;
    const size_t slot = (need + need / 2 + 4095) & ~(size_t)4095;
    CU(cudaMalloc(&g_design_arena, slot * kDesignRing));
    g_design_slot = slot;
  }
  const unsigned int s = g_design_next.fetch_add(1) % kDesignRing;
  *out = static_cast<char*>(g_design_arena) + (size_t)s * g_design_slot;
  return 0;
}

// x_host: m x p column-major.  Uploads xg / gid on `st`.
int prepare_design(const double* x_host, int m, int p, cudaStream_t st, DesignDev* out) {
  const int ps = p | 1;
  std::vector<int> gid(m), rep;
  for (int j = 0; j < m; j++) {
    int found = -1;
    for (size_t g = 0; g < rep.size() && found < 0; g++) {
      bool same = true;
      for (int k = 0; k < p && same; k++) same = (x_host[j + (size_t)m * k] == x_host[rep[g] + (size_t)m * k]);
      if (same) found = (int)g;
    }
    if (found < 0) {
      if (rep.size() > 32) break;   // too many distinct rows: samplewise
      found = (int)rep.size();
      rep.push_back(j);
    }
    gid[j] = found;
  }
  const bool grouped = rep.size() <= 32;
  const int rows = grouped ? (int)rep.size() : m;
  std::vector<double> xg((size_t)rows * ps, 0.0);
  for (int r = 0; r < rows; r++) {
    const int j = grouped ? rep[r] : r;
    for (int k = 0; k < p; k++) xg[(size_t)r * ps + k] = x_host[j + (size_t)m * k];
  }
  if (!grouped)
    for (int j = 0; j < m; j++) gid[j] = j;
  const size_t xbytes = xg.size() * sizeof(double), gbytes = (size_t)m * sizeof(int);
  const size_t need = ((xbytes + 15) & ~(size_t)15) + gbytes;
  char* base = nullptr;
  if (next_design_slot(need, &base)) return 1;
  // pageable H2D copies are staged before cudaMemcpyAsync returns, so the vectors may die at scope exit
  CU(cudaMemcpyAsync(base, xg.data(), xbytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(base + ((xbytes + 15) & ~(size_t)15), gid.data(), gbytes, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  out->xg = reinterpret_cast<const double*>(base);
  out->gid = reinterpret_cast<const int*>(base + ((xbytes + 15) & ~(size_t)15));
  out->G = grouped ? (int)rep.size() : m;
  out->grouped = grouped ? 1 : 0;
  out->saturated = 0;
  out->sat_logdet = 0.0;
  if (grouped && (int)rep.size() == p) {
    // log|det| of the p x p matrix of distinct rows (partial-pivot LU)
    std::vector<double> a((size_t)p * p);
    for (int r = 0; r < p; r++)
      for (int k = 0; k < p; k++) a[(size_t)r * p + k] = xg[(size_t)r * ps + k];
    double logdet = 0.0;
    bool ok = true;
    for (int c = 0; c < p && ok; c++) {
      int pr = c;
      for (int r = c + 1; r < p; r++)
        if (fabs(a[(size_t)r * p + c]) > fabs(a[(size_t)pr * p + c])) pr = r;
      if (fabs(a[(size_t)pr * p + c]) < 1e-12) { ok = false; break; }
      if (pr != c)
        for (int k = 0; k < p; k++) { double t = a[(size_t)c * p + k]; a[(size_t)c * p + k] = a[(size_t)pr * p + k]; a[(size_t)pr * p + k] = t; }
      logdet += log(fabs(a[(size_t)c * p + c]));
      for (int r = c + 1; r < p; r++) {
        const double f = a[(size_t)r * p + c] / a[(size_t)c * p + c];
        for (int k = c; k < p; k++) a[(size_t)r * p + k] -= f * a[(size_t)c * p + k];
      }
    }
    if (ok) {
      out->saturated = 1;
      out->sat_logdet = 2.0 * logdet;
    }
  }
  return 0;
}

// design matrix given as a DEVICE pointer (the *_dev entry points): fetch it (m*p doubles) and analyse.
// This synchronises `st`; it only happens on the general-p path.
int prepare_design_from_device(const double* x_dev, int m, int p, cudaStream_t st, DesignDev* out) {
  std::vector<double> xh((size_t)m * p);
  CU(cudaMemcpyAsync(xh.data(), x_dev, xh.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return prepare_design(xh.data(), m, p, st, out);
}

// The library keeps per-process device state (work-queue arenas, workspace, pinned staging): one process drives one
// GPU, as in the one-process-per-GPU deployment the engine is built for.  A second device in the same process is
// refused loudly instead of silently using buffers that live on the first one.
int g_bound_device = -1;
int check_device() {
  int dev = -1;
  CU(cudaGetDevice(&dev));
  if (g_bound_device < 0) g_bound_device = dev;
  if (dev != g_bound_device)
    return fail("libb200nb is bound to CUDA device %d in this process (current device is %d): use one process per GPU",
                g_bound_device, dev);
  return 0;
}

int check_dims(int n, int m, int p) {
  if (n < 0 || m < 1 || p < 1) return fail("bad dimensions n=%d m=%d p=%d", n, m, p);
  if (p > nb::kMaxP) return fail("p=%d not supported (max %d design columns)", p, nb::kMaxP);
  if (n > 0 && check_device()) return 1;
  return 0;
}

long long ld_for(int m) { return ((long long)m + 3) & ~3LL; }

// Rows [g0, g0 + n) of a column-major host matrix with n_total rows: copy to the device, convert to gene-major.
int upload_matrix(const void* host, int n_total, int g0, int n, int m, int elem, Slot raw, Slot dst, cudaStream_t st,
                  void** out) {
  void *d_raw, *d_dst;
  const long long ld = ld_for(m);
  if (ws_get(raw, (size_t)n * m * elem, &d_raw)) return 1;
  if (ws_get(dst, (size_t)n * ld * elem + 64, &d_dst)) return 1;
  if (h2d_block(d_raw, host, (size_t)n_total, (size_t)g0, (size_t)n, m, elem, st)) return 1;
  CU(nb::launch_to_gene_major(d_raw, d_dst, n, m, ld, elem, st));
  g_launches++;
  *out = d_dst;
  return 0;
}

int upload_vec(const void* host, size_t bytes, Slot s, cudaStream_t st, void** out) {
  void* d;
  if (ws_get(s, bytes, &d)) return 1;
  CU(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, st));
  *out = d;
  return 0;
}

// B200NB_HOST_TIMING=1: the host entry points print where their wall time goes (upload incl. layout conversion /
// kernels / download) to stderr; it adds a stream synchronisation between the phases, so it is a diagnosis aid, not
// something to leave on while measuring.
bool host_timing() {
  static const bool on = [] {
    const char* e = getenv("B200NB_HOST_TIMING");
    return e && atoi(e) != 0;
  }();
  return on;
}
struct PhaseClock {
  std::chrono::steady_clock::time_point t0;
  double ms[3] = {0.0, 0.0, 0.0};
  int phase = 0;
  bool on;
  cudaStream_t st;
  explicit PhaseClock(cudaStream_t s) : on(host_timing()), st(s) {
    if (on) t0 = std::chrono::steady_clock::now();
  }
  void next() {   // close the current phase
    if (!on || phase > 2) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    ms[phase++] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    t0 = t1;
  }
  void report(const char* what, int g0, int n) {
    if (on) fprintf(stderr, "b200nb timing %s genes [%d, %d): upload %.3f ms, kernels %.3f ms, download %.3f ms\n", what, g0,
                    g0 + n, ms[0], ms[1], ms[2]);
  }
};

// ---------------------------------------------------------------- gene-chunked host path (opt-in)
// B200NB_CHUNK_GENES=<genes per chunk> splits one host call into gene chunks that B200NB_CHUNK_WORKERS (default 2,
// max 4) host threads process independently -- each with its own stream, device workspace and pinned ring -- so that
// one chunk's H2D staging overlaps another chunk's kernels and D2H.  Genes are independent, so the results are those
// of the unchunked call.  Off by default (0): it has not been timed on the GPU yet.
int chunk_genes() {
  static int v = [] {
    const char* e = getenv("B200NB_CHUNK_GENES");
    const int g = e ? atoi(e) : 0;
    return g < 0 ? 0 : g;
  }();
  return v;
}
int chunk_workers() {
  static int v = [] {
    const char* e = getenv("B200NB_CHUNK_WORKERS");
    const int w = e ? atoi(e) : 2;
    return w < 1 ? 1 : (w > kMaxWorkers ? kMaxWorkers : w);
  }();
  return v;
}

// body(g0, count) handles genes [g0, g0 + count) through the current thread's context; returns non-zero on failure
// with the message in the thread-local g_err.
template <typename F>
int run_chunked(int n, F&& body) {
  const int want = chunk_genes();
  if (want <= 0 || n < 2 * want) return body(0, n);
  const int K = (n + want - 1) / want;
  const int gc = (n + K - 1) / K;                       // balanced chunks
  const int W = chunk_workers() < K ? chunk_workers() : K;
  std::atomic<int> next{0}, failed{0};
  char errs[kMaxWorkers][sizeof(g_err)];
  for (int w = 0; w < kMaxWorkers; w++) errs[w][0] = 0;
  auto work = [&](int w) {
    t_ctx = &g_ctx[w];
    if (w > 0 && g_bound_device >= 0) cudaSetDevice(g_bound_device);
    for (;;) {
      const int k = next.fetch_add(1);
      if (k >= K || failed.load()) break;
      const int g0 = k * gc, cnt = (n - g0 < gc) ? n - g0 : gc;
      if (cnt <= 0) break;
      if (body(g0, cnt)) {
        memcpy(errs[w], g_err, sizeof(g_err));
        failed.store(1);
        break;
      }
    }
    t_ctx = &g_ctx[0];
  };
  std::vector<std::thread> pool;
  for (int w = 1; w < W; w++) pool.emplace_back(work, w);
  work(0);
  for (auto& t : pool) t.join();
  if (failed.load()) {
    for (int w = 0; w < kMaxWorkers; w++)
      if (errs[w][0]) { memcpy(g_err, errs[w], sizeof(g_err)); break; }
    return 1;
  }
  return 0;
}

}  // namespace

extern "C" {

const char* b200nb_last_error(void) { return g_err; }
const char* b200nb_version(void) { return "b200nb 0.1 (sm_100a)"; }
long long b200nb_kernel_launches(void) { return g_launches.load(); }

int b200nb_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return c;
}

void b200nb_release_workspace(void) {
  std::lock_guard<std::mutex> lk(g_call_mu);
  for (int c = 0; c < kMaxWorkers; c++) {
    Workspace& ws = g_ctx[c].ws;
    Staging& sg = g_ctx[c].stage;
    for (int s = 0; s < S_NSLOTS; s++) {
      if (ws.p[s]) cudaFree(ws.p[s]);
      ws.p[s] = nullptr;
      ws.bytes[s] = 0;
    }
    if (sg.ready) {
      for (int i = 0; i < kStageRing; i++) {
        cudaFreeHost(sg.buf[i]);
        cudaEventDestroy(sg.ev[i]);
        sg.buf[i] = nullptr;
      }
      sg.ready = false;
    }
  }
}

/* ------------------------------------------------------------------ device entry points */

int b200nb_fit_disp_dev(const void* y, int y_type, const double* x, const double* mu_hat, const double* log_alpha,
                        const double* log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                        double kappa_0, double tol, int maxit, int use_prior, const double* weights,
                        int use_weights, double weight_threshold, int use_cr, int n, int m, int p, long long ld,
                        double* out_log_alpha, int32_t* out_iter, int32_t* out_iter_accept,
                        double* out_last_change, double* out_initial_lp, double* out_initial_dlp,
                        double* out_last_lp, double* out_last_dlp, double* out_last_d2lp, void* stream) {
  if (check_dims(n, m, p)) return 1;
  if (ld < m || (ld & 3)) return fail("ld=%lld must be >= m and a multiple of 4", ld);
  if (use_weights && !weights) return fail("use_weights set but weights == NULL");
  nb::DispArgs a{};
  a.y = y; a.y_is_f64 = (y_type == B200NB_Y_F64); a.mu = mu_hat; a.w = use_weights ? weights : nullptr; a.x = x;
  a.log_alpha_in = log_alpha; a.prior_mean = log_alpha_prior_mean;
  a.prior_sigmasq = log_alpha_prior_sigmasq; a.min_log_alpha = min_log_alpha; a.kappa_0 = kappa_0; a.tol = tol;
  a.maxit = maxit; a.use_prior = use_prior; a.use_weights = use_weights; a.use_cr = use_cr;
  a.weight_threshold = weight_threshold; a.n = n; a.m = m; a.p = p; a.ld = ld;
  a.log_alpha = out_log_alpha; a.iter = out_iter; a.iter_accept = out_iter_accept; a.last_change = out_last_change;
  a.initial_lp = out_initial_lp; a.initial_dlp = out_initial_dlp; a.last_lp = out_last_lp; a.last_dlp = out_last_dlp;
  a.last_d2lp = out_last_d2lp; a.grid = nullptr; a.grid_n = 0;
  if (n == 0) return 0;
  if (next_scratch(nb::disp_scratch_bytes(n), &a.scratch)) return 1;
  if (use_generic(p)) {
    DesignDev dd;
    if (prepare_design_from_device(x, m, p, (cudaStream_t)stream, &dd)) return 1;
    a.xg = dd.xg; a.gid = dd.gid; a.G = dd.G; a.grouped = dd.grouped;
    a.saturated = dd.saturated && !use_weights; a.sat_logdet = dd.sat_logdet;
    CU(nb::launch_fit_disp_generic(a, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
  }
  CU(nb::launch_fit_disp(a, (cudaStream_t)stream));
  g_launches += 2;   // classify + line search
  return 0;
}

int b200nb_fit_disp_grid_dev(const void* y, int y_type, const double* x, const double* mu_hat,
                             const double* disp_grid, int disp_grid_n, const double* log_alpha_prior_mean,
                             double log_alpha_prior_sigmasq, int use_prior, const double* weights, int use_weights,
                             double weight_threshold, int use_cr, int n, int m, int p, long long ld,
                             double* out_log_alpha, void* stream) {
  if (check_dims(n, m, p)) return 1;
  if (ld < m || (ld & 3)) return fail("ld=%lld must be >= m and a multiple of 4", ld);
  if (disp_grid_n < 2) return fail("disp_grid needs at least 2 points");
  if (use_weights && !weights) return fail("use_weights set but weights == NULL");
  nb::DispArgs a{};
  a.y = y; a.y_is_f64 = (y_type == B200NB_Y_F64); a.mu = mu_hat; a.w = use_weights ? weights : nullptr; a.x = x;
  a.log_alpha_in = nullptr; a.prior_mean = log_alpha_prior_mean; a.prior_sigmasq = log_alpha_prior_sigmasq;
  a.use_prior = use_prior; a.use_weights = use_weights; a.use_cr = use_cr; a.weight_threshold = weight_threshold;
  a.n = n; a.m = m; a.p = p; a.ld = ld; a.log_alpha = out_log_alpha; a.grid = disp_grid; a.grid_n = disp_grid_n;
  if (n == 0) return 0;
  if (next_scratch(nb::disp_scratch_bytes(n), &a.scratch)) return 1;
  if (use_generic(p)) {
    DesignDev dd;
    if (prepare_design_from_device(x, m, p, (cudaStream_t)stream, &dd)) return 1;
    a.xg = dd.xg; a.gid = dd.gid; a.G = dd.G; a.grouped = dd.grouped;
    a.saturated = dd.saturated && !use_weights; a.sat_logdet = dd.sat_logdet;
    CU(nb::launch_fit_disp_generic(a, (cudaStream_t)stream));
    g_launches++;
    return 0;
  }
  CU(nb::launch_fit_disp(a, (cudaStream_t)stream));
  g_launches++;
  return 0;
}

int b200nb_fit_beta_dev(const void* y, int y_type, const double* x, const double* nf, int nf_is_vector,
                        const double* alpha_hat, const double* contrast, const double* beta_mat,
                        const double* lambda, const double* weights, int use_weights, double tol, int maxit,
                        int use_qr, double minmu, int n, int m, int p, long long ld, double* out_beta_mat,
                        double* out_beta_var_mat, double* out_iter, double* out_hat_diagonals,
                        double* out_contrast_num, double* out_contrast_denom, double* out_deviance, double* out_mu,
                        void* stream) {
  if (check_dims(n, m, p)) return 1;
  if (ld < m || (ld & 3)) return fail("ld=%lld must be >= m and a multiple of 4", ld);
  if (use_weights && !weights) return fail("use_weights set but weights == NULL");
  if (maxit < 0) return fail("maxit must be >= 0");
  unsigned int* ctr;
  if (next_counter(&ctr)) return 1;
  nb::BetaArgs a{};
  a.y = y; a.y_is_f64 = (y_type == B200NB_Y_F64); a.nf = nf; a.nf_is_vector = nf_is_vector;
  a.w = use_weights ? weights : nullptr; a.x = x; a.alpha_hat = alpha_hat; a.contrast = contrast; a.beta_in = beta_mat;
  a.lambda = lambda; a.use_weights = use_weights; a.tol = tol; a.maxit = maxit; a.use_qr = use_qr; a.minmu = minmu;
  a.n = n; a.m = m; a.p = p; a.ld = ld; a.beta_out = out_beta_mat; a.beta_var = out_beta_var_mat; a.iter = out_iter;
  a.hat_diag = out_hat_diagonals; a.mu_out = out_mu; a.contrast_num = out_contrast_num;
  a.contrast_denom = out_contrast_denom; a.deviance = out_deviance; a.counter = ctr;
  if (use_generic(p)) {
    if (n == 0) return 0;
    DesignDev dd;
    if (prepare_design_from_device(x, m, p, (cudaStream_t)stream, &dd)) return 1;
    a.xg = dd.xg; a.gid = dd.gid; a.G = dd.G; a.grouped = dd.grouped;
    CU(nb::launch_fit_beta_generic(a, (cudaStream_t)stream));
    g_launches++;
    return 0;
  }
  CU(nb::launch_fit_beta(a, (cudaStream_t)stream));
  if (n > 0) g_launches++;
  return 0;
}

int b200nb_nb_loglik_dev(const void* y, int y_type, const double* x, const double* nf, int nf_is_vector,
                         const double* alpha_hat, const double* beta_mat, const double* weights, int use_weights, int n,
                         int m, int p, long long ld, double* out_loglik, double* out_mu, void* stream) {
  if (check_dims(n, m, p)) return 1;
  if (ld < m || (ld & 3)) return fail("ld=%lld must be >= m and a multiple of 4", ld);
  if (use_weights && !weights) return fail("use_weights set but weights == NULL");
  if (!out_loglik) return fail("out_loglik == NULL");
  nb::LogLikArgs a{};
  a.y = y; a.y_is_f64 = (y_type == B200NB_Y_F64); a.x = x; a.nf = nf; a.nf_is_vector = nf_is_vector;
  a.alpha = alpha_hat; a.beta = beta_mat; a.w = use_weights ? weights : nullptr; a.n = n; a.m = m; a.p = p; a.ld = ld;
  a.loglik = out_loglik; a.mu_out = out_mu;
  CU(nb::launch_nb_loglik(a, (cudaStream_t)stream));
  if (n > 0) g_launches++;
  return 0;
}

int b200nb_to_gene_major_dev(const void* src, void* dst, int n, int m, long long ld, int elem_size, void* stream) {
  CU(nb::launch_to_gene_major(src, dst, n, m, ld, elem_size, (cudaStream_t)stream));
  g_launches++;
  return 0;
}

int b200nb_to_col_major_dev(const double* src, double* dst, int n, int m, long long ld, void* stream) {
  CU(nb::launch_to_col_major(src, dst, n, m, ld, (cudaStream_t)stream));
  g_launches++;
  return 0;
}

int b200nb_prep_dev(const void* y, int y_type, const double* x, const double* proj, const double* size_factors,
                    double xim, double min_disp, double max_disp, double minmu, int n, int m, int p, long long ld,
                    double* base_mean, double* base_var, int32_t* all_zero, double* alpha0, double* mu_lin,
                    double* beta0, void* stream) {
  if (check_dims(n, m, p)) return 1;
  if (ld < m || (ld & 3)) return fail("ld=%lld must be >= m and a multiple of 4", ld);
  if (m <= p) return fail("prep needs m > p (m=%d, p=%d)", m, p);
  nb::PrepArgs a{};
  a.y = y; a.y_is_f64 = (y_type == B200NB_Y_F64); a.x = x; a.proj = proj; a.size_factors = size_factors; a.xim = xim;
  a.min_disp = min_disp; a.max_disp = max_disp; a.minmu = minmu; a.n = n; a.m = m; a.p = p; a.ld = ld;
  a.base_mean = base_mean; a.base_var = base_var; a.all_zero = all_zero; a.alpha0 = alpha0; a.mu_lin = mu_lin;
  a.beta0 = beta0;
  CU(nb::launch_prep(a, (cudaStream_t)stream));
  if (n > 0) g_launches++;
  return 0;
}

int b200nb_trend_fit_dev(const double* means, const double* disps, int n, double min_disp, double* out4,
                         void* stream) {
  if (n < 1) return fail("trend fit needs at least one gene");
  CU(nb::launch_trend_fit(means, disps, n, min_disp, out4, (cudaStream_t)stream));
  g_launches++;
  return 0;
}

int b200nb_cooks_dev(const void* y, int y_type, const double* mu, const double* hat, const double* size_factors,
                     const int32_t* cell_ptr, const int32_t* cell_samples, int ncell, int n, int m, int p,
                     long long ld, double* cooks, double* max_cooks, double* robust_disp, void* stream) {
  if (check_dims(n, m, p)) return 1;
  if (ld < m || (ld & 3)) return fail("ld=%lld must be >= m and a multiple of 4", ld);
  if (ncell < 1 || ncell > m) return fail("bad number of design cells %d", ncell);
  nb::CooksArgs a{};
  a.y = y; a.y_is_f64 = (y_type == B200NB_Y_F64); a.mu = mu; a.hat = hat; a.size_factors = size_factors;
  a.cell_ptr = cell_ptr; a.cell_samples = cell_samples; a.ncell = ncell; a.n = n; a.m = m; a.p = p; a.ld = ld;
  a.cooks = cooks; a.max_cooks = max_cooks; a.robust_disp = robust_disp;
  CU(nb::launch_cooks(a, (cudaStream_t)stream));
  if (n > 0) g_launches++;
  return 0;
}

int b200nb_size_factors_dev(const void* y, int y_type, int poscounts, int n, int m, long long ld,
                            double* loggeomeans, double* scratch_gm, double* scratch_cm, double* size_factors,
                            int32_t* n_finite, void* stream) {
  if (check_dims(n, m, 1)) return 1;
  if (ld < m || (ld & 3)) return fail("ld=%lld must be >= m and a multiple of 4", ld);
  if (n < 1) return fail("size factors need at least one gene");
  nb::SizeFactorArgs a{};
  a.y = y; a.y_is_f64 = (y_type == B200NB_Y_F64); a.poscounts = poscounts ? 1 : 0; a.n = n; a.m = m; a.ld = ld;
  a.loggeomeans = loggeomeans; a.ratios = scratch_gm; a.ratios_colmajor = scratch_cm; a.size_factors = size_factors;
  a.n_finite = n_finite;
  CU(nb::launch_size_factors(a, (cudaStream_t)stream));
  g_launches += 3;
  return 0;
}

/* ------------------------------------------------------------------ host entry points */
/* Each entry point validates, takes the call lock and hands gene blocks [g0, g0 + n) of the caller's R-layout arrays
 * (n_total rows) to a *_block routine; without B200NB_CHUNK_GENES there is one block, the whole call. */

static int fit_disp_block(const void* y, int y_type, const double* x, const double* mu_hat, const double* log_alpha,
                          const double* log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                          double kappa_0, double tol, int maxit, int use_prior, const double* weights,
                          int use_weights, double weight_threshold, int use_cr, int n_total, int g0, int n, int m,
                          int p, double* out_log_alpha, int32_t* out_iter, int32_t* out_iter_accept,
                          double* out_last_change, double* out_initial_lp, double* out_initial_dlp,
                          double* out_last_lp, double* out_last_dlp, double* out_last_d2lp) {
  cudaStream_t st;
  if (ws_stream(&st)) return 1;
  PhaseClock clk(st);
  const int ye = (y_type == B200NB_Y_F64) ? 8 : 4;
  const long long ld = ld_for(m);
  void *d_y, *d_mu, *d_w = nullptr, *d_x, *d_la, *d_pm, *d_outd, *d_outi;
  if (upload_matrix(y, n_total, g0, n, m, ye, S_RAW0, S_Y, st, &d_y)) return 1;
  if (upload_matrix(mu_hat, n_total, g0, n, m, 8, S_RAW1, S_MU, st, &d_mu)) return 1;
  if (use_weights && upload_matrix(weights, n_total, g0, n, m, 8, S_RAW2, S_W, st, &d_w)) return 1;
  if (upload_vec(x, sizeof(double) * m * p, S_X, st, &d_x)) return 1;
  if (upload_vec(log_alpha + g0, sizeof(double) * n, S_V0, st, &d_la)) return 1;
  if (upload_vec(log_alpha_prior_mean + g0, sizeof(double) * n, S_V1, st, &d_pm)) return 1;
  if (ws_get(S_OUTD, sizeof(double) * 7 * n, &d_outd)) return 1;
  if (ws_get(S_OUTI, sizeof(int32_t) * 2 * n, &d_outi)) return 1;
  double* od = (double*)d_outd;
  int32_t* oi = (int32_t*)d_outi;
  clk.next();
  if (b200nb_fit_disp_dev(d_y, y_type, (const double*)d_x, (const double*)d_mu, (const double*)d_la,
                          (const double*)d_pm, log_alpha_prior_sigmasq, min_log_alpha, kappa_0, tol, maxit, use_prior,
                          (const double*)d_w, use_weights, weight_threshold, use_cr, n, m, p, ld, od, oi, oi + n,
                          od + n, od + 2 * (size_t)n, od + 3 * (size_t)n, od + 4 * (size_t)n, od + 5 * (size_t)n,
                          od + 6 * (size_t)n, st))
    return 1;
  clk.next();
  double* outs[7] = {out_log_alpha, out_last_change, out_initial_lp, out_initial_dlp, out_last_lp, out_last_dlp,
                     out_last_d2lp};
  for (int k = 0; k < 7; k++)
    CU(cudaMemcpyAsync(outs[k] + g0, od + (size_t)k * n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_iter + g0, oi, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_iter_accept + g0, oi + n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  clk.next();
  clk.report("fitDisp", g0, n);
  return 0;
}

int b200nb_fit_disp(const void* y, int y_type, const double* x, const double* mu_hat, const double* log_alpha,
                    const double* log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                    double kappa_0, double tol, int maxit, int use_prior, const double* weights, int use_weights,
                    double weight_threshold, int use_cr, int n, int m, int p,
                    double* out_log_alpha, int32_t* out_iter, int32_t* out_iter_accept, double* out_last_change,
                    double* out_initial_lp, double* out_initial_dlp, double* out_last_lp, double* out_last_dlp,
                    double* out_last_d2lp) {
  if (check_dims(n, m, p)) return 1;
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(g_call_mu);
  auto block = [&](int g0, int cnt) {
    return fit_disp_block(y, y_type, x, mu_hat, log_alpha, log_alpha_prior_mean, log_alpha_prior_sigmasq, min_log_alpha,
                          kappa_0, tol, maxit, use_prior, weights, use_weights, weight_threshold, use_cr, n, g0, cnt, m,
                          p, out_log_alpha, out_iter, out_iter_accept, out_last_change, out_initial_lp,
                          out_initial_dlp, out_last_lp, out_last_dlp, out_last_d2lp);
  };
  if (use_generic(p)) return block(0, n);   // the general-p path analyses the design with a stream sync per launch
  unsigned int* presize;
  if (next_scratch(nb::disp_scratch_bytes(n), &presize)) return 1;   // size the shared arena before workers start
  return run_chunked(n, block);
}

static int fit_disp_grid_block(const void* y, int y_type, const double* x, const double* mu_hat,
                               const double* disp_grid, int disp_grid_n, const double* log_alpha_prior_mean,
                               double log_alpha_prior_sigmasq, int use_prior, const double* weights, int use_weights,
                               double weight_threshold, int use_cr, int n_total, int g0, int n, int m, int p,
                               double* out_log_alpha) {
  cudaStream_t st;
  if (ws_stream(&st)) return 1;
  const int ye = (y_type == B200NB_Y_F64) ? 8 : 4;
  const long long ld = ld_for(m);
  void *d_y, *d_mu, *d_w = nullptr, *d_x, *d_grid, *d_pm, *d_out;
  if (upload_matrix(y, n_total, g0, n, m, ye, S_RAW0, S_Y, st, &d_y)) return 1;
  if (upload_matrix(mu_hat, n_total, g0, n, m, 8, S_RAW1, S_MU, st, &d_mu)) return 1;
  if (use_weights && upload_matrix(weights, n_total, g0, n, m, 8, S_RAW2, S_W, st, &d_w)) return 1;
  if (upload_vec(x, sizeof(double) * m * p, S_X, st, &d_x)) return 1;
  if (upload_vec(disp_grid, sizeof(double) * disp_grid_n, S_V0, st, &d_grid)) return 1;
  if (upload_vec(log_alpha_prior_mean + g0, sizeof(double) * n, S_V1, st, &d_pm)) return 1;
  if (ws_get(S_OUTD, sizeof(double) * n, &d_out)) return 1;
  if (b200nb_fit_disp_grid_dev(d_y, y_type, (const double*)d_x, (const double*)d_mu, (const double*)d_grid,
                               disp_grid_n, (const double*)d_pm, log_alpha_prior_sigmasq, use_prior,
                               (const double*)d_w, use_weights, weight_threshold, use_cr, n, m, p, ld, (double*)d_out,
                               st))
    return 1;
  CU(cudaMemcpyAsync(out_log_alpha + g0, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

int b200nb_fit_disp_grid(const void* y, int y_type, const double* x, const double* mu_hat, const double* disp_grid,
                         int disp_grid_n, const double* log_alpha_prior_mean, double log_alpha_prior_sigmasq,
                         int use_prior, const double* weights, int use_weights, double weight_threshold, int use_cr,
                         int n, int m, int p, double* out_log_alpha) {
  if (check_dims(n, m, p)) return 1;
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(g_call_mu);
  auto block = [&](int g0, int cnt) {
    return fit_disp_grid_block(y, y_type, x, mu_hat, disp_grid, disp_grid_n, log_alpha_prior_mean,
                               log_alpha_prior_sigmasq, use_prior, weights, use_weights, weight_threshold, use_cr, n,
                               g0, cnt, m, p, out_log_alpha);
  };
  if (use_generic(p)) return block(0, n);
  unsigned int* presize;
  if (next_scratch(nb::disp_scratch_bytes(n), &presize)) return 1;
  return run_chunked(n, block);
}

// B200NB_DETECT_SF=1 (opt-in, not yet timed): R always hands fitBeta an n x m matrix of normalisation factors, which for
// the usual size-factor analysis is the same row n times (R/core.R:2221-2228).  One parallel read of the matrix tells;
// if so, only the m factors cross PCIe and the kernel takes its size-factor-vector path (log nf once per CTA instead
// of once per gene and sample).  Same values in, same results out; any NaN or differing entry keeps the matrix path.
static bool detect_sf() {
  static const bool on = [] {
    const char* e = getenv("B200NB_DETECT_SF");
    return e && atoi(e) != 0;
  }();
  return on;
}
static bool rows_identical(const double* a, size_t n, int m) {
  std::atomic<int> differs{0};
  const int T = stage_threads();
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
  for (int j = 0; j < m; j++) {
    if (differs.load(std::memory_order_relaxed)) continue;
    const double* c = a + (size_t)j * n;
    const double v = c[0];
    bool same = (v == v);
    for (size_t i0 = 0; i0 < n && same; i0 += 4096) {
      const size_t i1 = (i0 + 4096 < n) ? i0 + 4096 : n;
      int bad = 0;
      for (size_t i = i0; i < i1; i++) bad |= (c[i] != v);
      same = !bad;
    }
    if (!same) differs.store(1, std::memory_order_relaxed);
  }
  return differs.load() == 0;
}

static int fit_beta_block(const void* y, int y_type, const double* x, const double* nf, const double* sf_vector,
                          const double* alpha_hat,
                          const double* contrast, const double* beta_mat, const double* lambda,
                          const double* weights, int use_weights, double tol, int maxit, int use_qr, double minmu,
                          int n_total, int g0, int n, int m, int p, double* out_beta_mat, double* out_beta_var_mat,
                          double* out_iter, double* out_hat_diagonals, double* out_contrast_num,
                          double* out_contrast_denom, double* out_deviance, double* out_mu) {
  cudaStream_t st;
  if (ws_stream(&st)) return 1;
  PhaseClock clk(st);
  const int ye = (y_type == B200NB_Y_F64) ? 8 : 4;
  const long long ld = ld_for(m);
  void *d_y, *d_nf, *d_w = nullptr, *d_x, *d_alpha, *d_contrast, *d_lambda, *d_bin, *d_bout, *d_bvar, *d_outd;
  void *d_h = nullptr, *d_mu = nullptr, *d_hc = nullptr, *d_muc = nullptr;
  if (upload_matrix(y, n_total, g0, n, m, ye, S_RAW0, S_Y, st, &d_y)) return 1;
  if (sf_vector) {
    if (upload_vec(sf_vector, sizeof(double) * m, S_NF, st, &d_nf)) return 1;
  } else {
    if (upload_matrix(nf, n_total, g0, n, m, 8, S_RAW1, S_NF, st, &d_nf)) return 1;
  }
  if (use_weights && upload_matrix(weights, n_total, g0, n, m, 8, S_RAW2, S_W, st, &d_w)) return 1;
  if (upload_vec(x, sizeof(double) * m * p, S_X, st, &d_x)) return 1;
  if (upload_vec(alpha_hat + g0, sizeof(double) * n, S_V0, st, &d_alpha)) return 1;
  if (upload_vec(contrast, sizeof(double) * p, S_V1, st, &d_contrast)) return 1;
  if (upload_vec(lambda, sizeof(double) * p, S_V2, st, &d_lambda)) return 1;
  if (ws_get(S_BETA_IN, sizeof(double) * n * p, &d_bin)) return 1;
  if (h2d_block(d_bin, beta_mat, (size_t)n_total, (size_t)g0, (size_t)n, p, 8, st)) return 1;
  if (ws_get(S_BETA_OUT, sizeof(double) * n * p, &d_bout)) return 1;
  if (ws_get(S_BETA_VAR, sizeof(double) * n * p, &d_bvar)) return 1;
  if (ws_get(S_OUTD, sizeof(double) * 4 * n, &d_outd)) return 1;
  if (out_hat_diagonals) {
    if (ws_get(S_H, sizeof(double) * n * ld, &d_h)) return 1;
    if (ws_get(S_HC, sizeof(double) * n * m, &d_hc)) return 1;
  }
  if (out_mu) {
    if (ws_get(S_MUO, sizeof(double) * n * ld, &d_mu)) return 1;
    if (ws_get(S_MUC, sizeof(double) * n * m, &d_muc)) return 1;
  }
  double* od = (double*)d_outd;
  clk.next();
  if (b200nb_fit_beta_dev(d_y, y_type, (const double*)d_x, (const double*)d_nf, sf_vector ? 1 : 0, (const double*)d_alpha,
                          (const double*)d_contrast, (const double*)d_bin, (const double*)d_lambda,
                          (const double*)d_w, use_weights, tol, maxit, use_qr, minmu, n, m, p, ld, (double*)d_bout,
                          (double*)d_bvar, od, (double*)d_h, od + n, od + 2 * (size_t)n, od + 3 * (size_t)n,
                          (double*)d_mu, st))
    return 1;
  clk.next();
  if (out_hat_diagonals) {
    if (b200nb_to_col_major_dev((const double*)d_h, (double*)d_hc, n, m, ld, st)) return 1;
    if (d2h_block(out_hat_diagonals, d_hc, (size_t)n_total, (size_t)g0, (size_t)n, m, 8, st)) return 1;
  }
  if (out_mu) {
    if (b200nb_to_col_major_dev((const double*)d_mu, (double*)d_muc, n, m, ld, st)) return 1;
    if (d2h_block(out_mu, d_muc, (size_t)n_total, (size_t)g0, (size_t)n, m, 8, st)) return 1;
  }
  if (n == n_total) {
    CU(cudaMemcpyAsync(out_beta_mat, d_bout, sizeof(double) * n * p, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_beta_var_mat, d_bvar, sizeof(double) * n * p, cudaMemcpyDeviceToHost, st));
  } else {
    if (d2h_block(out_beta_mat, d_bout, (size_t)n_total, (size_t)g0, (size_t)n, p, 8, st)) return 1;
    if (d2h_block(out_beta_var_mat, d_bvar, (size_t)n_total, (size_t)g0, (size_t)n, p, 8, st)) return 1;
  }
  CU(cudaMemcpyAsync(out_iter + g0, od, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_contrast_num + g0, od + n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_contrast_denom + g0, od + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_deviance + g0, od + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  clk.next();
  clk.report("fitBeta", g0, n);
  return 0;
}

int b200nb_fit_beta(const void* y, int y_type, const double* x, const double* nf, const double* alpha_hat,
                    const double* contrast, const double* beta_mat, const double* lambda, const double* weights,
                    int use_weights, double tol, int maxit, int use_qr, double minmu, int n, int m, int p,
                    double* out_beta_mat, double* out_beta_var_mat, double* out_iter, double* out_hat_diagonals,
                    double* out_contrast_num, double* out_contrast_denom, double* out_deviance, double* out_mu) {
  if (check_dims(n, m, p)) return 1;
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(g_call_mu);
  if (out_hat_diagonals) advise_hugepages(out_hat_diagonals, sizeof(double) * (size_t)n * m);
  if (out_mu) advise_hugepages(out_mu, sizeof(double) * (size_t)n * m);
  std::vector<double> sfv;
  if (detect_sf() && rows_identical(nf, (size_t)n, m)) {
    sfv.resize(m);
    for (int j = 0; j < m; j++) sfv[j] = nf[(size_t)j * n];
  }
  const double* sf_vector = sfv.empty() ? nullptr : sfv.data();
  auto block = [&](int g0, int cnt) {
    return fit_beta_block(y, y_type, x, nf, sf_vector, alpha_hat, contrast, beta_mat, lambda, weights, use_weights, tol, maxit,
                          use_qr, minmu, n, g0, cnt, m, p, out_beta_mat, out_beta_var_mat, out_iter, out_hat_diagonals,
                          out_contrast_num, out_contrast_denom, out_deviance, out_mu);
  };
  if (use_generic(p)) return block(0, n);
  return run_chunked(n, block);
}

int b200nb_test_special(const double* x, int n, double* out_lgamma, double* out_digamma, double* out_trigamma) {
  if (n <= 0) return 0;
  std::lock_guard<std::mutex> lk(g_call_mu);
  cudaStream_t st;
  if (ws_stream(&st)) return 1;
  void *d_x, *d_o;
  if (upload_vec(x, sizeof(double) * n, S_V0, st, &d_x)) return 1;
  if (ws_get(S_OUTD, sizeof(double) * 3 * n, &d_o)) return 1;
  double* o = (double*)d_o;
  CU(nb::launch_special_test((const double*)d_x, n, o, o + n, o + 2 * (size_t)n, st));
  g_launches++;
  CU(cudaMemcpyAsync(out_lgamma, o, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_digamma, o + n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_trigamma, o + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
