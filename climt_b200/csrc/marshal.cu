// climt_b200 -- device-side host-marshal arithmetic of the drop-in components, for callers whose
// model state already lives in HBM: specific humidity -> volume mixing ratio (climt/_core/util.py:47-86), temperature on
// interface levels by ln-p weights (climt/_core/util.py:89-142), cos(zenith) (rrtmg/sw/component.py:591).
// One thread per (column, level); column-fastest rows.
#include <cuda_runtime.h>

#include "../../include/climt_b200.h"
#include "engine_common.h"

namespace {
__global__ void __launch_bounds__(128) k_marshal(int ncol, int nlay, const double* __restrict__ q, const double* __restrict__ t,
                                                 const double* __restrict__ tsfc, const double* __restrict__ p,
                                                 const double* __restrict__ p_int, const double* __restrict__ zenith,
                                                 double* __restrict__ h2ovmr, double* __restrict__ tlev, double* __restrict__ coszen) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;  // interface level 0..nlay
  if (c >= ncol) return;
  const size_t n = (size_t)ncol;
  if (h2ovmr && k < nlay) h2ovmr[k * n + c] = q[k * n + c] * 28.964 / 18.02;
  if (tlev) {
    double v;
    if (k == 0) v = tsfc[c];
    else if (k == nlay) v = t[(size_t)(nlay - 1) * n + c];
    else {
      const double lp1 = log(p[k * n + c]), lp0 = log(p[(size_t)(k - 1) * n + c]);
      const double w = (log(p_int[k * n + c]) - lp1) / (lp0 - lp1);
      v = t[k * n + c] - w * (t[k * n + c] - t[(size_t)(k - 1) * n + c]);
    }
    tlev[k * n + c] = v;
  }
  if (coszen && k == 0) coszen[c] = cos(zenith[c]);
}
}  // namespace

// used by the host-pointer calls of the engines when the caller hands over specific humidity / no interface temperatures
// (cb200_{lw,sw}_set_host_marshal): the same kernel on one chunk's device buffers
namespace cb {
cudaError_t marshal_launch(int ncol, int nlay, const double* q, const double* t, const double* tsfc, const double* p, const double* p_int,
                           double* h2ovmr, double* tlev, cudaStream_t st) {
  k_marshal<<<dim3((ncol + 127) / 128, nlay + 1), 128, 0, st>>>(ncol, nlay, q, t, tsfc, p, p_int, nullptr, h2ovmr, tlev, nullptr);
  return cudaGetLastError();
}
}  // namespace cb

extern "C" int cb200_marshal_device(int device, int ncol, int nlay, const double* q, const double* t, const double* tsfc,
                                    const double* p, const double* p_int, const double* zenith, double* h2ovmr, double* tlev,
                                    double* coszen, void* stream) {
  if (ncol <= 0 || nlay <= 0) { cb::set_global_error("marshal: bad ncol/nlay"); return -3; }
  if ((h2ovmr && !q) || (tlev && (!t || !tsfc || !p || !p_int)) || (coszen && !zenith)) {
    cb::set_global_error("marshal: missing input for a requested output");
    return -3;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) {
    k_marshal<<<dim3((ncol + 127) / 128, nlay + 1), 128, 0, (cudaStream_t)stream>>>(ncol, nlay, q, t, tsfc, p, p_int, zenith, h2ovmr,
                                                                                   tlev, coszen);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { cb::set_global_error(cudaGetErrorString(e)); return -1; }
  return 0;
}
