// Multi-threaded host memcpy for the replay-batch staging path (pageable numpy arrays -> the pinned buffer the H2D copy
// reads).  A single thread moves ~18 GB/s, so the 10 MB ManiSkill batch costs 0.56 ms of the 1.9 ms end-to-end update;
// a few threads bring it to the PCIe copy's own 0.19 ms.  A small persistent pool (created on first use): python-level
// thread pools pay ~25 us per task for these ~1 MB pieces, which eats the gain.
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace pcrl {
namespace {

struct CopyPool {
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  char* dst = nullptr;
  const char* src = nullptr;
  int64_t nbytes = 0;
  int parts = 0;      // chunks of the current job (worker w takes chunk w + 1, the caller chunk 0)
  uint64_t gen = 0;   // job generation
  int pending = 0;
  bool stop = false;

  void run(int w) {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(m);
      cv_go.wait(lk, [&] { return stop || gen != seen; });
      if (stop) return;
      seen = gen;
      const int p = parts;
      char* d = dst;
      const char* s = src;
      const int64_t n = nbytes;
      lk.unlock();
      if (w + 1 < p) {
        const int64_t chunk = (n / p + 63) & ~int64_t(63);
        const int64_t lo = std::min<int64_t>(n, chunk * (w + 1)), hi = std::min<int64_t>(n, chunk * (w + 2));
        if (hi > lo) std::memcpy(d + lo, s + lo, (size_t)(hi - lo));
      }
      lk.lock();
      if (--pending == 0) cv_done.notify_one();
    }
  }

  void ensure(int n_workers) {
    while ((int)workers.size() < n_workers) {
      const int w = (int)workers.size();
      workers.emplace_back([this, w] { run(w); });
    }
  }

  void copy(void* d, const void* s, int64_t n, int threads) {
    threads = std::max(1, std::min(threads, 16));
    if (threads == 1 || n < (256 << 10)) {
      std::memcpy(d, s, (size_t)n);
      return;
    }
    std::unique_lock<std::mutex> lk(m);
    ensure(threads - 1);
    dst = (char*)d;
    src = (const char*)s;
    nbytes = n;
    parts = threads;
    pending = (int)workers.size();
    ++gen;
    lk.unlock();
    cv_go.notify_all();
    const int64_t chunk = (n / threads + 63) & ~int64_t(63);
    std::memcpy(d, s, (size_t)std::min<int64_t>(n, chunk));
    lk.lock();
    cv_done.wait(lk, [&] { return pending == 0; });
  }

  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv_go.notify_all();
    for (auto& t : workers) t.join();
  }
};

// One pool per PROCESS: worker threads do not survive fork() (the reference forks rollout / evaluation workers), so a
// child that stages a batch gets a fresh pool instead of waiting for threads that only exist in its parent.  The parent's
// pool object is deliberately leaked in the child (its mutex / condition variables may be in any state).
CopyPool& pool() {
  static std::mutex guard;
  static CopyPool* p = nullptr;
  static pid_t owner = 0;
  std::lock_guard<std::mutex> lk(guard);
  const pid_t me = getpid();
  if (!p || owner != me) {
    p = new CopyPool();
    owner = me;
  }
  return *p;
}

}  // namespace
}  // namespace pcrl

extern "C" int pcrl_host_memcpy_mt(void* dst_host, const void* src_host, int64_t nbytes, int threads) {
  PCRL_CHECK_ARG((dst_host && src_host) || nbytes == 0);
  PCRL_CHECK_ARG(nbytes >= 0);
  if (nbytes == 0) return PCRL_OK;
  pcrl::pool().copy(dst_host, src_host, nbytes, threads);
  return PCRL_OK;
}

// One call per replay batch: every leaf is staged into its slot of the pinned buffer (pool memcpy; leaves with a NULL
// source were already written there by the caller) and its host->device copy is enqueued on `stream` right away, so the
// DMA of leaf i runs under the memcpy of leaf i+1 and the python side pays one foreign call instead of ~5 per leaf.
extern "C" int pcrl_upload_leaves(void* pinned_host, void* landing_dev, const void* const* srcs_host,
                                  const int64_t* offsets, const int64_t* sizes, int n_leaves, int threads, void* stream) {
  PCRL_CHECK_ARG(n_leaves >= 0 && (n_leaves == 0 || (pinned_host && landing_dev && srcs_host && offsets && sizes)));
  cudaStream_t st = pcrl::as_stream(stream);
  for (int i = 0; i < n_leaves; ++i) {
    PCRL_CHECK_ARG(offsets[i] >= 0 && sizes[i] >= 0);
    if (sizes[i] == 0) continue;
    char* h = static_cast<char*>(pinned_host) + offsets[i];
    if (srcs_host[i]) pcrl::pool().copy(h, srcs_host[i], sizes[i], threads);
    PCRL_CHECK_CUDA(cudaMemcpyAsync(static_cast<char*>(landing_dev) + offsets[i], h, (size_t)sizes[i],
                                    cudaMemcpyHostToDevice, st));
  }
  return PCRL_OK;
}
