// hostrt.h -- host-side runtime of the HOST entry points (capi.cu): what sits between R's pageable column-major
// buffers and the device.  Nothing here touches CUDA except one PCI-bus-id query; it is plain C++17 + pthreads.
//
//   Pool        persistent worker threads (no OpenMP: an R process may carry its own OpenMP runtime, and torchrun exports
//               OMP_NUM_THREADS=1), bound as a group to the CPUs of the NUMA node the GPU hangs off
//               (/sys/bus/pci/devices/<bus id>/numa_node -> /sys/devices/system/node/nodeK/cpulist, intersected with the
//               process affinity): 8 ranks on a two-socket box otherwise stage through the wrong socket's memory.
//               Recreated after fork() (BiocParallel's multicore back-end forks R workers).
//   hash_range  128-bit position-dependent content hash, commutative over blocks, so any partition of a buffer among
//               threads gives the same value; it is what makes the device-side input cache safe (capi.cu): a hit
//               requires every byte of the host buffer to hash to the cached value -- not a pointer comparison (R copies
//               the count matrix for each of the three calls of one DESeq() run, and recycles addresses across runs).
//   populate    MADV_POPULATE_WRITE of a freshly allocated result matrix, started while the GPU is busy, so that the
//               device-to-host scatter does not take one page fault per 4 KB.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif

namespace hostrt {

struct Hash128 {
  uint64_t a = 0, b = 0;
  bool operator==(const Hash128& o) const { return a == o.a && b == o.b; }
  void add(const Hash128& o) { a += o.a; b += o.b; }
};

// Hash of the 8-byte words [first_word, first_word + bytes/8) of a buffer (the last partial word is zero-padded).
// Each word is mixed with its global index, the per-word values are summed: block-order independent.
inline Hash128 hash_range(const void* p, size_t bytes, uint64_t first_word) {
  const unsigned char* s = static_cast<const unsigned char*>(p);
  const uint64_t K1 = 0x9E3779B97F4A7C15ull, K2 = 0xD6E8FEB86659FD93ull, K3 = 0xA0761D6478BD642Full;
  uint64_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  uint64_t ik = (first_word + 1) * K1;
  const size_t nw = bytes / 8;
  size_t i = 0;
  for (; i + 2 <= nw; i += 2) {
    uint64_t w0, w1;
    memcpy(&w0, s + 8 * i, 8);
    memcpy(&w1, s + 8 * i + 8, 8);
    uint64_t t0 = (w0 ^ ik) * K2;
    ik += K1;
    uint64_t t1 = (w1 ^ ik) * K2;
    ik += K1;
    t0 ^= t0 >> 29;
    t1 ^= t1 >> 29;
    a0 += t0;
    a1 += t1;
    b0 += (t0 * K3) ^ w0;
    b1 += (t1 * K3) ^ w1;
  }
  for (; i < nw; i++) {
    uint64_t w0;
    memcpy(&w0, s + 8 * i, 8);
    uint64_t t0 = (w0 ^ ik) * K2;
    ik += K1;
    t0 ^= t0 >> 29;
    a0 += t0;
    b0 += (t0 * K3) ^ w0;
  }
  if (bytes & 7) {
    uint64_t w0 = 0;
    memcpy(&w0, s + 8 * nw, bytes & 7);
    uint64_t t0 = (w0 ^ ik) * K2;
    t0 ^= t0 >> 29;
    a0 += t0;
    b0 += (t0 * K3) ^ w0;
  }
  Hash128 h;
  h.a = a0 + a1;
  h.b = b0 + b1;
  return h;
}

inline int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e || !*e) return dflt;
  const int v = atoi(e);
  return v < lo ? lo : (v > hi ? hi : v);
}

// CPUs of NUMA node `node` that the process may run on; empty when unknown
inline std::vector<int> node_cpus(int node) {
  std::vector<int> out;
  if (node < 0) return out;
  char path[128];
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  FILE* f = fopen(path, "r");
  if (!f) return out;
  char buf[4096];
  const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
  fclose(f);
  if (!ok) return out;
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return out;
  for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int lo = 0, hi = 0;
    if (sscanf(tok, "%d-%d", &lo, &hi) == 2) {
    } else if (sscanf(tok, "%d", &lo) == 1) {
      hi = lo;
    } else {
      continue;
    }
    for (int c = lo; c <= hi && c < CPU_SETSIZE; c++)
      if (CPU_ISSET(c, &allowed)) out.push_back(c);
  }
  return out;
}

inline int pci_numa_node(const char* pci_bus_id) {   // "0000:1b:00.0" (any case) -> node, or -1
  if (!pci_bus_id || !*pci_bus_id) return -1;
  std::string id(pci_bus_id);
  for (char& c : id) c = (char)tolower((unsigned char)c);
  const std::string path = "/sys/bus/pci/devices/" + id + "/numa_node";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

inline int affinity_count() {
  cpu_set_t s;
  CPU_ZERO(&s);
  if (sched_getaffinity(0, sizeof(s), &s) != 0) return (int)sysconf(_SC_NPROCESSORS_ONLN);
  return CPU_COUNT(&s);
}

class Pool {
 public:
  // `cpus`: the CPU group the workers are bound to (empty = inherit the caller's affinity)
  Pool(int nthreads, const std::vector<int>& cpus) : nworkers_(nthreads > 1 ? nthreads - 1 : 0), pid_(getpid()), cpus_(cpus) {
    for (int w = 0; w < nworkers_; w++) th_.emplace_back([this, w] { worker(w + 1); });
  }
  ~Pool() {
    if (getpid() != pid_) {   // forked child: the threads do not exist here; do not join what was never cloned
      for (auto& t : th_) t.detach();
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_.fetch_add(1);
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int threads() const { return nworkers_ + 1; }
  pid_t pid() const { return pid_; }
  const std::vector<int>& cpus() const { return cpus_; }

  // f(task) for task in [0, ntasks), dynamically distributed over the workers and the caller; returns when all are done
  template <typename F>
  void parallel_for(size_t ntasks, F&& f) {
    if (ntasks == 0) return;
    if (nworkers_ == 0 || ntasks == 1) {
      for (size_t t = 0; t < ntasks; t++) f(t);
      return;
    }
    auto tramp = [](void* ctx, size_t t) { (*static_cast<typename std::remove_reference<F>::type*>(ctx))(t); };
    {
      std::lock_guard<std::mutex> lk(mu_);
      call_ = tramp;
      ctx_ = (void*)&f;
      ntasks_ = ntasks;
      next_.store(0, std::memory_order_relaxed);
      running_.store(nworkers_, std::memory_order_relaxed);
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (size_t t; (t = next_.fetch_add(1, std::memory_order_relaxed)) < ntasks;) f(t);
    // wait for the workers to leave this generation (spin briefly: the tail is short)
    for (int spin = 0; running_.load(std::memory_order_acquire) != 0; spin++) {
      if (spin > 2000) {
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait_for(lk, std::chrono::microseconds(200), [this] { return running_.load() == 0; });
      } else {
        cpu_relax();
      }
    }
  }

 private:
  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  void bind_self() {
    if (cpus_.empty()) return;
    cpu_set_t s;
    CPU_ZERO(&s);
    for (int c : cpus_) CPU_SET(c, &s);
    pthread_setaffinity_np(pthread_self(), sizeof(s), &s);   // a group binding (the kernel balances inside the node)
  }
  void worker(int /*id*/) {
    bind_self();
    uint64_t seen = 0;
    for (;;) {
      // wait for a new generation: spin a little (back-to-back parallel_for calls), then sleep
      uint64_t g = gen_.load(std::memory_order_acquire);
      for (int spin = 0; g == seen && spin < 4000; spin++) {
        cpu_relax();
        g = gen_.load(std::memory_order_acquire);
      }
      if (g == seen) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
        g = gen_.load(std::memory_order_acquire);
      }
      seen = g;
      if (stop_) return;
      void (*call)(void*, size_t);
      void* ctx;
      size_t nt;
      {
        std::lock_guard<std::mutex> lk(mu_);   // pairs with the publisher: call_/ctx_/ntasks_ are consistent with gen_
        call = call_;
        ctx = ctx_;
        nt = ntasks_;
        if (stop_) return;
      }
      for (size_t t; (t = next_.fetch_add(1, std::memory_order_relaxed)) < nt;) call(ctx, t);
      if (running_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> lk(mu_);
        done_cv_.notify_all();
      }
    }
  }

  int nworkers_;
  pid_t pid_;
  std::vector<int> cpus_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::atomic<uint64_t> gen_{0};
  std::atomic<size_t> next_{0};
  std::atomic<int> running_{0};
  void (*call_)(void*, size_t) = nullptr;
  void* ctx_ = nullptr;
  size_t ntasks_ = 0;
  bool stop_ = false;
};

// fill the page tables of [p, p + bytes) for writing (the pages are about to be overwritten completely)
inline void populate_write(void* p, size_t bytes) {
  const uintptr_t a = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, b = ((uintptr_t)p + bytes) & ~(uintptr_t)4095;
  if (b <= a) return;
  if (madvise(reinterpret_cast<void*>(a), b - a, MADV_POPULATE_WRITE) == 0) return;
  // older kernels: touch one byte per page (zero is as good as anything: every byte is rewritten afterwards)
  for (uintptr_t q = a; q < b; q += 4096) *reinterpret_cast<volatile char*>(q) = 0;
}

}  // namespace hostrt
