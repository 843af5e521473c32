// Host <-> device transfer pipeline of the stateless drop-in entry points (atlj_force_lj / atlj_force_le / atlj_hessian_lj:
// the force(box, r_cut, r) call of md_lj_ll_module.py:69 with NumPy arrays).
//
// A NumPy array is pageable memory: cudaMemcpyAsync from it is a synchronous, staged copy at a fraction of the PCIe
// rate, and at 2 M atoms (49 MB each way) the call was 20x slower than the device loop.  Here
//   * a pageable source is copied chunk by chunk into a persistent page-locked ring by a small pool of host threads and
//     every chunk is sent as soon as it is complete (DMA overlaps the host copy of the next chunks);
//   * a pageable destination is filled the same way in reverse (DMA chunk k while the threads copy chunk k-1 out);
//   * page-locked memory (cudaHostAlloc / cudaHostRegister: the arrays examples_b200 hands out for `f`) is recognised
//     with cudaPointerGetAttributes and transferred with ONE asynchronous copy, no staging.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "hostpipe.cuh"

namespace atlj {

namespace {

class Workers {
 public:
  explicit Workers(int n) {
    for (int i = 0; i < n; i++) th_.emplace_back([this] { run(); });
  }
  ~Workers() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  void submit(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> g(m_);
      q_.push_back(std::move(f));
    }
    cv_.notify_one();
  }
  int size() const { return (int)th_.size(); }

 private:
  void run() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return stop_ || !q_.empty(); });
        if (stop_ && q_.empty()) return;
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
    }
  }
  std::vector<std::thread> th_;
  std::deque<std::function<void()>> q_;
  std::mutex m_;
  std::condition_variable cv_;
  bool stop_ = false;
};

struct Pipe {
  Workers *workers = nullptr;
  unsigned char *ring[2] = {nullptr, nullptr};  // page-locked staging: [0] towards the device, [1] from it
  size_t ring_bytes[2] = {0, 0};
  std::vector<cudaEvent_t> ev;
};
Pipe g_pipe;

constexpr size_t CHUNK = 4u << 20;  // 4 MB: ~75 us of DMA at 55 GB/s, ~0.4 ms of one host thread's memcpy

Workers *workers() {
  if (!g_pipe.workers) {
    unsigned hc = std::thread::hardware_concurrency();
    int n = hc >= 16 ? 8 : (hc >= 8 ? 4 : (hc >= 4 ? 2 : 1));
    const char *e = getenv("ATLJ_COPY_THREADS");
    if (e && atoi(e) > 0) n = atoi(e);
    g_pipe.workers = new Workers(n);
  }
  return g_pipe.workers;
}

int ensure_ring(int which, size_t bytes) {
  if (bytes <= g_pipe.ring_bytes[which]) return ATLJ_OK;
  if (g_pipe.ring[which]) cudaFreeHost(g_pipe.ring[which]);
  g_pipe.ring[which] = nullptr, g_pipe.ring_bytes[which] = 0;
  const size_t want = bytes + bytes / 8;
  ATLJ_CUDA(cudaHostAlloc((void **)&g_pipe.ring[which], want, cudaHostAllocDefault));
  g_pipe.ring_bytes[which] = want;
  return ATLJ_OK;
}

int ensure_events(size_t n) {
  while (g_pipe.ev.size() < n) {
    cudaEvent_t e;
    ATLJ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_pipe.ev.push_back(e);
  }
  return ATLJ_OK;
}

}  // namespace

bool host_ptr_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int hostpipe_threads() { return workers()->size(); }

// host -> device, complete in stream order on `st` when the call returns (the source may be reused right away only if
// it was pageable: a page-locked source is still being read until the stream reaches this copy)
int hostpipe_h2d(cudaStream_t st, void *dst_dev, const void *src_host, size_t bytes) {
  if (bytes == 0) return ATLJ_OK;
  if (host_ptr_is_pinned(src_host) || bytes < (1u << 20)) {
    ATLJ_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
    return ATLJ_OK;
  }
  int rc = ensure_ring(0, bytes);
  if (rc) return rc;
  // the ring is reused by every call: the previous call's copies out of it must have been consumed
  ATLJ_CUDA(cudaStreamSynchronize(st));
  const size_t nchunk = (bytes + CHUNK - 1) / CHUNK;
  std::vector<std::atomic<int>> done(nchunk);
  for (auto &d : done) d.store(0, std::memory_order_relaxed);
  Workers *w = workers();
  unsigned char *ring = g_pipe.ring[0];
  const unsigned char *src = static_cast<const unsigned char *>(src_host);
  for (size_t k = 0; k < nchunk; k++) {
    const size_t off = k * CHUNK, len = bytes - off < CHUNK ? bytes - off : CHUNK;
    std::atomic<int> *flag = &done[k];
    w->submit([=] {
      memcpy(ring + off, src + off, len);
      flag->store(1, std::memory_order_release);
    });
  }
  for (size_t k = 0; k < nchunk; k++) {
    while (done[k].load(std::memory_order_acquire) == 0) std::this_thread::yield();
    const size_t off = k * CHUNK, len = bytes - off < CHUNK ? bytes - off : CHUNK;
    ATLJ_CUDA(cudaMemcpyAsync(static_cast<unsigned char *>(dst_dev) + off, ring + off, len, cudaMemcpyHostToDevice, st));
  }
  return ATLJ_OK;
}

// device -> host; returns when the data is in dst_host (the stream is synchronised)
int hostpipe_d2h(cudaStream_t st, void *dst_host, const void *src_dev, size_t bytes) {
  if (bytes == 0) return ATLJ_OK;
  if (host_ptr_is_pinned(dst_host) || bytes < (1u << 20)) {
    ATLJ_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
    ATLJ_CUDA(cudaStreamSynchronize(st));
    return ATLJ_OK;
  }
  int rc = ensure_ring(1, bytes);
  if (rc) return rc;
  const size_t nchunk = (bytes + CHUNK - 1) / CHUNK;
  if ((rc = ensure_events(nchunk))) return rc;
  unsigned char *ring = g_pipe.ring[1];
  unsigned char *dst = static_cast<unsigned char *>(dst_host);
  for (size_t k = 0; k < nchunk; k++) {
    const size_t off = k * CHUNK, len = bytes - off < CHUNK ? bytes - off : CHUNK;
    ATLJ_CUDA(cudaMemcpyAsync(ring + off, static_cast<const unsigned char *>(src_dev) + off, len, cudaMemcpyDeviceToHost, st));
    ATLJ_CUDA(cudaEventRecord(g_pipe.ev[k], st));
  }
  std::atomic<size_t> left(nchunk);
  Workers *w = workers();
  for (size_t k = 0; k < nchunk; k++) {
    ATLJ_CUDA(cudaEventSynchronize(g_pipe.ev[k]));
    const size_t off = k * CHUNK, len = bytes - off < CHUNK ? bytes - off : CHUNK;
    std::atomic<size_t> *l = &left;
    w->submit([=] {
      memcpy(dst + off, ring + off, len);
      l->fetch_sub(1, std::memory_order_release);
    });
  }
  while (left.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  return ATLJ_OK;
}

}  // namespace atlj
