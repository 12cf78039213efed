// climt_b200 -- grey-atmosphere longwave (config 1 plumbing path), sm_100a.
// One thread per column, columns fastest: every access of a warp is one contiguous 256-byte row, each input is read
// once per sweep and each output written once -- a pure HBM-streaming kernel (1 480 B per 30-level column).
// Reference: climt/_components/radiation.py:162-190 (`_gray_lw_kernel_np`) and :93-99 (tendency).
#include <cuda_runtime.h>

#include <string>

#include "../../include/climt_b200.h"
#include "engine_common.h"

namespace {
__global__ void __launch_bounds__(128) k_gray_lw(int ncol, int nlay, const double* __restrict__ T,
                                                 const double* __restrict__ p_int, const double* __restrict__ T_surface,
                                                 const double* __restrict__ tau, double sigma, double g_over_cpd,
                                                 double* __restrict__ down, double* __restrict__ up,
                                                 double* __restrict__ tend) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncol) return;
  const size_t n = (size_t)ncol;
  const double ts = T_surface[i];
  double f = sigma * ((ts * ts) * (ts * ts));
  up[i] = f;
  double tau_lo = tau[i];
  for (int k = 1; k <= nlay; ++k) {
    const double tau_hi = tau[(size_t)k * n + i];
    const double trans = exp(-(tau_hi - tau_lo));
    const double t = T[(size_t)(k - 1) * n + i];
    const double t4 = sigma * ((t * t) * (t * t));
    f = f * trans + t4 * (1.0 - trans);
    up[(size_t)k * n + i] = f;
    tau_lo = tau_hi;
  }
  // downward sweep fused with the flux divergence: net(k+1) and net(k) are both known when level k is reached
  double d = 0.0;
  down[(size_t)nlay * n + i] = 0.0;
  double tau_hi = tau[(size_t)nlay * n + i];
  double net_hi = f - 0.0;
  double p_hi = p_int[(size_t)nlay * n + i];
  for (int k = nlay - 1; k >= 0; --k) {
    const double tau_k = tau[(size_t)k * n + i];
    const double trans = exp(-(tau_hi - tau_k));
    const double t = T[(size_t)k * n + i];
    const double t4 = sigma * ((t * t) * (t * t));
    d = d * trans + t4 * (1.0 - trans);
    down[(size_t)k * n + i] = d;
    const double net_k = up[(size_t)k * n + i] - d;
    const double p_k = p_int[(size_t)k * n + i];
    tend[(size_t)k * n + i] = g_over_cpd * (net_hi - net_k) / (p_hi - p_k);
    tau_hi = tau_k;
    net_hi = net_k;
    p_hi = p_k;
  }
}
}  // namespace

extern "C" int cb200_gray_lw_run_device(int device, int ncol, int nlay, const double* t, const double* p_int,
                                        const double* t_surf, const double* tau, double sigma, double g, double cpd,
                                        double* lw_down, double* lw_up, double* tendency, void* stream) {
  if (ncol <= 0 || nlay <= 0) { cb::set_global_error("gray: bad ncol/nlay"); return -3; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { cb::set_global_error(cudaGetErrorString(e)); return -1; }
  k_gray_lw<<<(ncol + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ncol, nlay, t, p_int, t_surf, tau, sigma, g / cpd,
                                                                  lw_down, lw_up, tendency);
  e = cudaGetLastError();
  if (e != cudaSuccess) { cb::set_global_error(cudaGetErrorString(e)); return -1; }
  return 0;
}

extern "C" int cb200_gray_lw_run_host(int device, int ncol, int nlay, const double* t, const double* p_int,
                                      const double* t_surf, const double* tau, double sigma, double g, double cpd,
                                      double* lw_down, double* lw_up, double* tendency) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { cb::set_global_error(cudaGetErrorString(e)); return -1; }
  const size_t n = (size_t)ncol, L = (size_t)nlay;
  const size_t tot = L * n + 2 * (L + 1) * n + n + 2 * (L + 1) * n + L * n;
  double* d = nullptr;
  e = cudaMalloc(&d, tot * sizeof(double));
  if (e != cudaSuccess) { cb::set_global_error(cudaGetErrorString(e)); return -1; }
  double *dT = d, *dp = dT + L * n, *dtau = dp + (L + 1) * n, *dts = dtau + (L + 1) * n, *ddn = dts + n,
         *dup = ddn + (L + 1) * n, *dtd = dup + (L + 1) * n;
  cudaMemcpyAsync(dT, t, L * n * 8, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(dp, p_int, (L + 1) * n * 8, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(dtau, tau, (L + 1) * n * 8, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(dts, t_surf, n * 8, cudaMemcpyHostToDevice, 0);
  int rc = cb200_gray_lw_run_device(device, ncol, nlay, dT, dp, dts, dtau, sigma, g, cpd, ddn, dup, dtd, 0);
  if (rc == 0) {
    cudaMemcpyAsync(lw_down, ddn, (L + 1) * n * 8, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(lw_up, dup, (L + 1) * n * 8, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(tendency, dtd, L * n * 8, cudaMemcpyDeviceToHost, 0);
    e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) { cb::set_global_error(cudaGetErrorString(e)); rc = -1; }
  }
  cudaFree(d);
  return rc;
}
