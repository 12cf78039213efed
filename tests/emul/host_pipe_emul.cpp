// Test-only host build of the host-side helpers of the engines' host-pointer path (climt_b200/csrc/engine_common.h: the worker pool,
// the all-zero scan and its guard).  No CUDA: the HostPipe struct itself needs the runtime and is exercised by the GPU tests.
#include "../../climt_b200/csrc/engine_common.h"

#include <thread>

extern "C" int emul_pool_size() { return cb::WorkerPool::get().size(); }

// columns [c0, c0 + n) of nview (rows[i], ncol) arrays -> zero[i]
extern "C" void emul_all_zero(const double* const* base, const long* rows, int nview, long ncol, long c0, long n, int* zero) {
  std::vector<cb::ZeroView> v(nview);
  for (int i = 0; i < nview; ++i) v[i] = cb::ZeroView{base[i], (size_t)rows[i], (size_t)ncol, (size_t)c0, (size_t)n};
  std::vector<char> z(nview);
  bool* zb = reinterpret_cast<bool*>(z.data());
  cb::all_zero_parallel(v.data(), nview, zb);
  for (int i = 0; i < nview; ++i) zero[i] = zb[i] ? 1 : 0;
}

// two caller threads hammer the pool at once (the LW and SW engines' enqueue threads do): every task must run exactly once
extern "C" long emul_pool_stress(int calls, int ntasks) {
  std::atomic<long> done{0};
  auto caller = [&]() {
    for (int k = 0; k < calls; ++k) cb::WorkerPool::get().parallel_for(ntasks, [&](int) { done.fetch_add(1); });
  };
  std::thread a(caller), b(caller);
  a.join();
  b.join();
  return done.load();
}

extern "C" int emul_scan_guard(const double* gbytes_per_s, int n) {  // feeds n scans of 64 MiB at the given rates; -> index at which it trips, or -1
  cb::ScanGuard g;
  for (int i = 0; i < n; ++i)
    if (!g.note((size_t)64 << 20, (double)((size_t)64 << 20) / (gbytes_per_s[i] * 1e9))) return i;
  return -1;
}
