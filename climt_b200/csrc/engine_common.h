// climt_b200 -- bits shared by the engine translation units (host only).
#pragma once
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace cb {
inline std::string g_error;
inline std::mutex g_mu;
inline void set_global_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_error = s;
}
constexpr int kBlock = 128;

// A few persistent host threads for the scans below (created on first use; a call wakes them through a condition variable and
// takes part in the work itself).
class WorkerPool {
 public:
  static WorkerPool& get() {
    static WorkerPool p;
    return p;
  }
  // fn(task) for task in [0, ntasks), distributed dynamically; returns when all are done.  Calls are serialised.
  template <class F>
  void parallel_for(int ntasks, F&& fn) {
    if (ntasks <= 0) return;
    std::lock_guard<std::mutex> call(call_mu_);
    std::function<void(int)> f = std::forward<F>(fn);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &f;
      ntasks_ = ntasks;
      next_.store(0);
      pending_ = (int)threads_.size();
      ++generation_;
    }
    cv_.notify_all();
    run_tasks(f, ntasks);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }
  int size() const { return (int)threads_.size() + 1; }

 private:
  WorkerPool() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)(hw ? hw : 4u) / 2;
    // one process per GPU under torchrun: share the host cores between the ranks of the box
    if (const char* lws = std::getenv("LOCAL_WORLD_SIZE")) { const int r = std::atoi(lws); if (r > 1) n /= r; }
    if (const char* e = std::getenv("CLIMT_B200_HOST_THREADS")) n = std::atoi(e);
    n = n < 1 ? 1 : (n > 8 ? 8 : n);
    for (int t = 1; t < n; ++t) threads_.emplace_back([this] { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void run_tasks(const std::function<void(int)>& f, int ntasks) {
    for (;;) {
      const int i = next_.fetch_add(1);
      if (i >= ntasks) return;
      f(i);
    }
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int)>* f;
      int ntasks;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        f = fn_;
        ntasks = ntasks_;
      }
      run_tasks(*f, ntasks);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --pending_;
      }
      done_cv_.notify_one();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  int ntasks_ = 0, pending_ = 0;
  unsigned long generation_ = 0;
  std::atomic<int> next_{0};
  bool stop_ = false;
};

// Where the reference-named init symbols find their table blob (see lw_engine.cu / sw_engine.cu: default_blob)
inline std::string find_table_blob(const char* env_name, const char* file, void* symbol_in_this_library) {
  if (const char* p = std::getenv(env_name)) return p;
  auto readable = [](const std::string& f) { FILE* fp = std::fopen(f.c_str(), "rb"); if (fp) std::fclose(fp); return fp != nullptr; };
  if (const char* c = std::getenv("CLIMT_B200_CACHE")) {
    const std::string f = std::string(c) + "/" + file;
    if (readable(f)) return f;
  }
  std::string dir = ".";
  Dl_info info;
  if (dladdr(symbol_in_this_library, &info) && info.dli_fname) {
    const std::string so = info.dli_fname;
    const size_t k = so.find_last_of('/');
    if (k != std::string::npos) dir = so.substr(0, k);
  }
  const std::string cached = dir + "/data/_cache/" + file, shipped = dir + "/data/" + file;
  return readable(cached) ? cached : shipped;
}

// One array of a host-pointer call as the scan sees it: columns [c0, c0 + n) of a (rows, ncol) array of doubles.
struct ZeroView {
  const double* base;
  size_t rows, ncol, c0, n;
};
// Is every byte of each view zero (+0.0 everywhere)?  Rows are cut into blocks of ~256 KiB that the pool's threads pull from a
// shared counter; a view is dropped from the scan as soon as one of its blocks holds a set bit.  Used by the host-pointer calls,
// chunk by chunk, to replace the PCIe transfer of an all-zero input (the reference ABI's always-present halocarbon, aerosol
// optical depth and cloud arrays: ~60 % of the longwave call's input bytes in the default configuration) by a device-side
// memset -- the same values in HBM, so results are unchanged.  The scan of chunk k + 1 overlaps the GPU work of chunk k.
// -0.0 counts as non-zero (the array is then simply transferred).
// Self-check of the scan: if it runs slower than 20 GB/s three times in a row (a host with few free cores: PCIe would have moved
// the bytes faster), the caller stops scanning and transfers everything.
struct ScanGuard {
  int slow = 0;
  bool note(size_t bytes, double seconds) {  // -> false: stop scanning
    if (bytes < (8u << 20)) return true;  // small scans are dominated by the wake-up of the pool, not by bandwidth
    slow = (double)bytes < 20e9 * seconds ? slow + 1 : 0;
    return slow < 3;
  }
};
inline void all_zero_parallel(const ZeroView* view, int nview, bool* zero) {
  struct Block { int v; size_t r0, r1; };
  std::vector<Block> blocks;
  std::vector<std::atomic<int>> dirty(nview);
  for (int a = 0; a < nview; ++a) {
    dirty[a].store(0);
    const size_t row_bytes = view[a].n * sizeof(double);
    size_t per = row_bytes ? (256u << 10) / row_bytes : 1;
    if (per < 1) per = 1;
    for (size_t r = 0; r < view[a].rows; r += per) blocks.push_back({a, r, r + per < view[a].rows ? r + per : view[a].rows});
  }
  WorkerPool::get().parallel_for((int)blocks.size(), [&](int i) {
    const Block& b = blocks[i];
    if (dirty[b.v].load(std::memory_order_relaxed)) return;
    const ZeroView& v = view[b.v];
    uint64_t acc = 0;
    for (size_t r = b.r0; r < b.r1 && !acc; ++r) {
      const uint64_t* p = reinterpret_cast<const uint64_t*>(v.base + r * v.ncol + v.c0);
      size_t k = 0;
      for (; k + 8 <= v.n; k += 8) acc |= p[k] | p[k + 1] | p[k + 2] | p[k + 3] | p[k + 4] | p[k + 5] | p[k + 6] | p[k + 7];
      for (; k < v.n; ++k) acc |= p[k];
    }
    if (acc) dirty[b.v].store(1, std::memory_order_relaxed);
  });
  for (int a = 0; a < nview; ++a) zero[a] = dirty[a].load() == 0;
}
}  // namespace cb

// ---------------------------------------------------------------------------------------------
// Host-buffer pipeline shared by the LW and SW engines (needs <cuda_runtime.h> before this header).
// A host call is cut into column chunks; chunk k's inputs are gathered from the caller's (nrow, ncol) arrays into a
// chunk-contiguous device slot with strided 2-D copies on the H2D stream while chunk k-1 computes and chunk k-2's
// fluxes drain on the D2H stream.  Two input slots and two output slots; events order slot reuse.
#ifdef __CUDACC__
namespace cb {
// marshal.cu: specific humidity -> volume mixing ratio (in place when h2ovmr == q) and ln-p interface temperatures of one chunk
cudaError_t marshal_launch(int ncol, int nlay, const double* q, const double* t, const double* tsfc, const double* p, const double* p_int,
                           double* h2ovmr, double* tlev, cudaStream_t st);

struct HostPipe {
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  cudaEvent_t in_done[2] = {nullptr, nullptr}, cmp_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
  double* d_in[2] = {nullptr, nullptr};
  double* d_out[2] = {nullptr, nullptr};
  size_t in_cap = 0, out_cap = 0;
  // Pinned staging for callers whose arrays are ordinary pageable memory (numpy arrays handed down by sympl): the driver would
  // otherwise bounce every strided 2-D copy through its own small staging buffer synchronously (r02, 8192 x 60 through
  // RRTMGLongwave / RRTMGShortwave.array_call: 35 ms per step against 5.6 ms from pinned buffers).  Rows are copied into / out of
  // these slots by the worker pool and cross PCIe as ONE contiguous asynchronous copy per array and chunk.
  double* h_in[2] = {nullptr, nullptr};
  double* h_out[2] = {nullptr, nullptr};
  size_t hin_cap = 0, hout_cap = 0;
  // development trace (CLIMT_B200_PIPE_TRACE=1): timing events at the stage boundaries of every chunk, printed by trace_dump()
  struct Mark { cudaEvent_t ev; int chunk, kind; double host_ms; };
  bool trace = false;
  std::vector<Mark> marks;
  static cudaEvent_t& epoch() { static cudaEvent_t e = nullptr; return e; }
  static double& epoch_host() { static double t = 0; return t; }
  static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  void mark(cudaStream_t st, int chunk_index, int kind) {
    if (!trace) return;
    if (!epoch()) {
      cudaEventCreate(&epoch());
      cudaEventRecord(epoch(), st);
      cudaEventSynchronize(epoch());
      epoch_host() = now_ms();
    }
    Mark m{nullptr, chunk_index, kind, now_ms() - epoch_host()};
    cudaEventCreate(&m.ev);
    cudaEventRecord(m.ev, st);
    marks.push_back(m);
  }
  void trace_dump(const char* tag) {
    if (!trace) return;
    static const char* names[5] = {"h2d_start", "h2d_done", "compute_start", "compute_done", "d2h_done"};
    for (auto& m : marks) {
      float ms = 0;
      cudaEventSynchronize(m.ev);
      cudaEventElapsedTime(&ms, epoch(), m.ev);
      std::fprintf(stderr, "PIPE %s chunk %d %-13s gpu %.3f ms  (enqueued at host %.3f ms)\n", tag, m.chunk, names[m.kind], ms, m.host_ms);
      cudaEventDestroy(m.ev);
    }
    std::fprintf(stderr, "PIPE %s returned to caller at host %.3f ms\n", tag, now_ms() - epoch_host());
    marks.clear();
  }
  int chunk = 4096;  // columns per full pipeline chunk (r01 B200, 8192 x 60, LW + SW overlapped: 2048 -> 6.33, 4096 -> 6.02, 8192 -> 5.99 ms with the ramp)

  // Columns of chunk k: the first two chunks are a quarter and a half of `chunk`, so that the kernels start after a short first
  // copy instead of idling through a full-sized one (r01 trace: 0.63 ms of 5.7 with uniform 4096-column chunks).
  bool ramp = true;
  int chunk_size(int k, int remaining) const {
    int n = chunk;
    if (ramp && k == 0) n = chunk / 4;
    else if (ramp && k == 1) n = chunk / 2;
    n = n / 128 * 128;
    if (n < 128) n = 128;
    return n < remaining ? n : remaining;
  }

  cudaError_t init() {
    if (s_in) return cudaSuccess;
    cudaError_t ce;
    if ((ce = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)) != cudaSuccess) return ce;
    if ((ce = cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking)) != cudaSuccess) return ce;
    if ((ce = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking)) != cudaSuccess) return ce;
    for (int i = 0; i < 2; ++i) {
      if ((ce = cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming)) != cudaSuccess) return ce;
      if ((ce = cudaEventCreateWithFlags(&cmp_done[i], cudaEventDisableTiming)) != cudaSuccess) return ce;
      if ((ce = cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming)) != cudaSuccess) return ce;
    }
    if (const char* hc = std::getenv("CLIMT_B200_HOST_CHUNK")) chunk = std::max(128, std::atoi(hc));
    if (const char* tr = std::getenv("CLIMT_B200_PIPE_TRACE")) trace = std::atoi(tr) != 0;
    if (const char* rp = std::getenv("CLIMT_B200_HOST_RAMP")) ramp = std::atoi(rp) != 0;
    return cudaSuccess;
  }
  cudaError_t ensure(size_t in_doubles, size_t out_doubles) {
    cudaError_t ce;
    if (in_doubles > in_cap) {
      for (int i = 0; i < 2; ++i) { cudaFree(d_in[i]); d_in[i] = nullptr; }
      in_cap = 0;
      for (int i = 0; i < 2; ++i)
        if ((ce = cudaMalloc(&d_in[i], in_doubles * sizeof(double))) != cudaSuccess) return ce;
      in_cap = in_doubles;
    }
    if (out_doubles > out_cap) {
      for (int i = 0; i < 2; ++i) { cudaFree(d_out[i]); d_out[i] = nullptr; }
      out_cap = 0;
      for (int i = 0; i < 2; ++i)
        if ((ce = cudaMalloc(&d_out[i], out_doubles * sizeof(double))) != cudaSuccess) return ce;
      out_cap = out_doubles;
    }
    return cudaSuccess;
  }
  cudaError_t ensure_staging(size_t in_doubles, size_t out_doubles) {
    cudaError_t ce;
    if (in_doubles > hin_cap) {
      for (int i = 0; i < 2; ++i) { if (h_in[i]) cudaFreeHost(h_in[i]); h_in[i] = nullptr; }
      hin_cap = 0;
      for (int i = 0; i < 2; ++i)
        if ((ce = cudaMallocHost(&h_in[i], in_doubles * sizeof(double))) != cudaSuccess) return ce;
      hin_cap = in_doubles;
    }
    if (out_doubles > hout_cap) {
      for (int i = 0; i < 2; ++i) { if (h_out[i]) cudaFreeHost(h_out[i]); h_out[i] = nullptr; }
      hout_cap = 0;
      for (int i = 0; i < 2; ++i)
        if ((ce = cudaMallocHost(&h_out[i], out_doubles * sizeof(double))) != cudaSuccess) return ce;
      hout_cap = out_doubles;
    }
    return cudaSuccess;
  }
  // Can the copy engine read / write this host pointer directly (page-locked: cudaMallocHost, cudaHostRegister, torch pin_memory)?
  static bool dma_able(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
  }
  // rows x n columns [c0, c0+n) of a pageable host (rows, ncol) array -> staging slot -> contiguous (rows, n) device block
  cudaError_t gather_staged(double* dst_dev, double* stage, const double* src, int rows, int ncol, int c0, int n, int inner = 1) const {
    const size_t w = (size_t)n * inner;
    const int per = (int)std::max<size_t>(1, (256u << 10) / (w * sizeof(double)));
    WorkerPool::get().parallel_for((rows + per - 1) / per, [&](int t) {
      for (int r = t * per; r < rows && r < (t + 1) * per; ++r)
        std::memcpy(stage + (size_t)r * w, src + ((size_t)r * ncol + c0) * inner, w * sizeof(double));
    });
    return cudaMemcpyAsync(dst_dev, stage, (size_t)rows * w * sizeof(double), cudaMemcpyHostToDevice, s_in);
  }
  cudaError_t scatter_staged_issue(double* stage, const double* src_dev, int rows, int n) const {
    return cudaMemcpyAsync(stage, src_dev, (size_t)rows * n * sizeof(double), cudaMemcpyDeviceToHost, s_out);
  }
  static void scatter_staged_finish(double* dst, const double* stage, int rows, int ncol, int c0, int n) {
    const int per = (int)std::max<size_t>(1, (256u << 10) / ((size_t)n * sizeof(double)));
    WorkerPool::get().parallel_for((rows + per - 1) / per, [&](int t) {
      for (int r = t * per; r < rows && r < (t + 1) * per; ++r)
        std::memcpy(dst + (size_t)r * ncol + c0, stage + (size_t)r * n, (size_t)n * sizeof(double));
    });
  }
  void destroy() {
    for (int i = 0; i < 2; ++i) {
      if (h_in[i]) cudaFreeHost(h_in[i]);
      if (h_out[i]) cudaFreeHost(h_out[i]);
      cudaFree(d_in[i]); cudaFree(d_out[i]);
      if (in_done[i]) cudaEventDestroy(in_done[i]);
      if (cmp_done[i]) cudaEventDestroy(cmp_done[i]);
      if (out_done[i]) cudaEventDestroy(out_done[i]);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_cmp) cudaStreamDestroy(s_cmp);
    if (s_out) cudaStreamDestroy(s_out);
  }
  // rows x n columns [c0, c0+n) of a host (rows, ncol) array -> contiguous (rows, n) device block
  // (`inner` > 1: arrays whose fastest axis is a per-column vector, e.g. taucld(nbnd, ncol, nlay) in Fortran order)
  cudaError_t gather(double* dst, const double* src, int rows, int ncol, int c0, int n, int inner = 1) const {
    return cudaMemcpy2DAsync(dst, (size_t)n * inner * sizeof(double), src + (size_t)c0 * inner,
                             (size_t)ncol * inner * sizeof(double), (size_t)n * inner * sizeof(double), (size_t)rows,
                             cudaMemcpyHostToDevice, s_in);
  }
  cudaError_t scatter(double* dst, const double* src, int rows, int ncol, int c0, int n) const {
    return cudaMemcpy2DAsync(dst + c0, (size_t)ncol * sizeof(double), src, (size_t)n * sizeof(double),
                             (size_t)n * sizeof(double), (size_t)rows, cudaMemcpyDeviceToHost, s_out);
  }
};
}  // namespace cb
#endif
