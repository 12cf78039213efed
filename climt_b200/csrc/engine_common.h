// climt_b200 -- bits shared by the engine translation units (host only).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>

namespace cb {
inline std::string g_error;
inline std::mutex g_mu;
inline void set_global_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_error = s;
}
constexpr int kBlock = 128;
}  // namespace cb

// ---------------------------------------------------------------------------------------------
// Host-buffer pipeline shared by the LW and SW engines (needs <cuda_runtime.h> before this header).
// A host call is cut into column chunks; chunk k's inputs are gathered from the caller's (nrow, ncol) arrays into a
// chunk-contiguous device slot with strided 2-D copies on the H2D stream while chunk k-1 computes and chunk k-2's
// fluxes drain on the D2H stream.  Two input slots and two output slots; events order slot reuse.
#ifdef __CUDACC__
namespace cb {
struct HostPipe {
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  cudaEvent_t in_done[2] = {nullptr, nullptr}, cmp_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
  double* d_in[2] = {nullptr, nullptr};
  double* d_out[2] = {nullptr, nullptr};
  size_t in_cap = 0, out_cap = 0;
  int chunk = 2048;  // columns per pipeline chunk (r01 B200 sweep: 1024-2048 best for LW+SW overlapped, 4096 for a lone engine)

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
  void destroy() {
    for (int i = 0; i < 2; ++i) {
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
