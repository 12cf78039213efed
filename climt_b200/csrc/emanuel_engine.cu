// climt_b200 -- Emanuel moist convection engine: CUDA kernel (sm_100a, one warp per column, persistent grid), host pipeline, C ABI.
// Warp-cooperative code: emanuel_core.cuh.  Reference: climt/_lib/emanuel/convect43c.f90, climt/_components/emanuel/
// _emanuel_convection.pyx, component.py, pure_python_v3.py (see include/climt_b200.h for the entry-point mapping).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/climt_b200.h"
#include "engine_common.h"
#include "emanuel_core.cuh"

using namespace cb::emanuel;

namespace {
#ifndef CB_EMANUEL_WARPS
#define CB_EMANUEL_WARPS 4  // columns (warps) per block; the block's shared memory is CB_EMANUEL_WARPS x V_COUNT x n1 doubles
#endif

// One warp per column, persistent grid: warp slot w handles columns w, w + nslots, ... and owns slot w of the matrix workspace.
__global__ void __launch_bounds__(32 * CB_EMANUEL_WARPS)
    k_emanuel(const Par par, const __grid_constant__ In in, const __grid_constant__ Work W, const __grid_constant__ Out out, int c0, int n,
              int NL, double dt) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * CB_EMANUEL_WARPS + warp, nslots = gridDim.x * CB_EMANUEL_WARPS;
  double* sv = smem + (size_t)warp * V_COUNT * W.n1;
  double* mw = W.m + (size_t)slot * M_COUNT * W.nm * W.nm;
  for (int c = slot; c < n; c += nslots) convect_warp(par, in, W, out, (size_t)c0 + c, lane, sv, mw, NL, dt);
}

#define CUDA_OK(call)                                                                         \
  do {                                                                                        \
    cudaError_t err__ = (call);                                                               \
    if (err__ != cudaSuccess) {                                                               \
      e->error = std::string(#call) + ": " + cudaGetErrorString(err__);                       \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

}  // namespace

struct cb200_emanuel_engine {
  int device = 0;
  Par par{};
  Work W{};
  int nslots = 0, cap_nl = 0, blocks_per_sm = 0, sm_count = 0;
  size_t smem_bytes = 0;
  int cap_nlev = 0;
  cb::HostPipe pipe;
  int32_t* d_iflag[2] = {nullptr, nullptr};
  int iflag_cap = 0;
  std::string error;
  int launches = 0;
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double kernel_ms = 0.0;

  void free_work() {
    cudaFree(W.m);
    W = Work{};
    nslots = cap_nl = cap_nlev = 0;
  }
  // shared memory per block and the persistent grid (blocks per SM from the occupancy calculator); one matrix slot per resident warp
  int ensure_work(int nlev, int NL) {
    cb200_emanuel_engine* e = this;
    if (W.m && nlev == cap_nlev && NL == cap_nl) return 0;
    free_work();
    W.n1 = nlev + 4;
    W.nm = NL + 2;
    smem_bytes = sizeof(double) * (size_t)CB_EMANUEL_WARPS * V_COUNT * W.n1;
    if (smem_bytes > 200 * 1024) { error = "emanuel: too many levels for the shared-memory column slices"; return -3; }
    CUDA_OK(cudaFuncSetAttribute(k_emanuel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_emanuel, 32 * CB_EMANUEL_WARPS, smem_bytes));
    if (blocks_per_sm < 1) { error = "emanuel: kernel does not fit on an SM"; return -1; }
    CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    nslots = sm_count * blocks_per_sm * CB_EMANUEL_WARPS;
    CUDA_OK(cudaMalloc(&W.m, sizeof(double) * (size_t)nslots * M_COUNT * W.nm * W.nm));
    cap_nlev = nlev; cap_nl = NL;
    return 0;
  }
};

extern "C" int cb200_emanuel_create(cb200_emanuel_engine** out, const cb200_emanuel_params* p, int device) {
  *out = nullptr;
  static_assert(sizeof(cb200_emanuel_params) == sizeof(Par), "cb200_emanuel_params layout");
  if (!p) { cb::set_global_error("emanuel: null parameters"); return -1; }
  // the argument checks of EmanuelConvection.__init__ (component.py:201-216)
  if (p->cu < 0 || p->cu > 1) { cb::set_global_error("Momentum transfer coefficient must be between 0 and 1."); return -3; }
  if (p->sigd < 0 || p->sigd > 1) { cb::set_global_error("Downdraft fraction must be between 0 and 1."); return -3; }
  if (p->sigs < 0 || p->sigs > 1) { cb::set_global_error("Outside cloud precipitation fraction must be between 0 and 1."); return -3; }
  if (p->minorig < 1) { cb::set_global_error("emanuel: minimum_convecting_layer must be >= 1"); return -3; }
  auto* e = new cb200_emanuel_engine();
  e->device = device;
  std::memcpy(&e->par, p, sizeof(Par));
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev1);
  if (ce != cudaSuccess) {
    cb::set_global_error(std::string("emanuel create: ") + cudaGetErrorString(ce));
    delete e;
    return -1;
  }
  *out = e;
  return 0;
}

extern "C" void cb200_emanuel_destroy(cb200_emanuel_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  e->free_work();
  cudaFree(e->d_iflag[0]); cudaFree(e->d_iflag[1]);
  e->pipe.destroy();
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  delete e;
}
extern "C" const char* cb200_emanuel_last_error(cb200_emanuel_engine* e) { return e ? e->error.c_str() : cb::g_error.c_str(); }
extern "C" int cb200_emanuel_last_launches(cb200_emanuel_engine* e) { return e->launches; }
extern "C" int cb200_emanuel_enable_timing(cb200_emanuel_engine* e, int on) { e->timing = on != 0; return 0; }
extern "C" double cb200_emanuel_last_kernel_ms(cb200_emanuel_engine* e) { return e->kernel_ms; }

namespace {

int validate(cb200_emanuel_engine* e, int ncol, int nlev, int NL, double dt, int qs_mode, const cb200_emanuel_inputs* in,
             const cb200_emanuel_outputs* out) {
  if (ncol <= 0 || nlev < 6) { e->error = "emanuel: need ncol > 0 and at least 6 levels"; return -3; }
  if (NL < 4 || NL > nlev - 2) { e->error = "emanuel: max_conv_lev must be in [4, nlev - 2] (the component passes nlev - 3)"; return -3; }
  if (!(dt > 0)) { e->error = "emanuel: the time step must be positive"; return -3; }
  if (qs_mode < 0 || qs_mode > 2) { e->error = "emanuel: qs_mode must be 0 (given), 1 (bolton_q_sat) or 2 (compute_qs)"; return -3; }
  if (!in->t || !in->q || !in->u || !in->v || !in->p || !in->ph || !in->cbmf || (qs_mode == QS_GIVEN && !in->qs)) {
    e->error = "emanuel: missing input array"; return -3;
  }
  if (!out->ft || !out->fq || !out->fu || !out->fv || !out->precip || !out->wd || !out->tprime || !out->qprime || !out->cbmf ||
      !out->cape || !out->iflag) {
    e->error = "emanuel: missing output array"; return -3;
  }
  return 0;
}

// one launch over n columns starting at column c0 of the arrays `in` / `out` describe
int launch(cb200_emanuel_engine* e, const In& in, const Out& out, int c0, int n, int NL, double dt, cudaStream_t st) {
  if (e->timing) cudaEventRecord(e->ev0, st);
  int blocks = (n + CB_EMANUEL_WARPS - 1) / CB_EMANUEL_WARPS;
  const int resident = e->sm_count * e->blocks_per_sm;
  if (blocks > resident) blocks = resident;
  k_emanuel<<<blocks, 32 * CB_EMANUEL_WARPS, e->smem_bytes, st>>>(e->par, in, e->W, out, c0, n, NL, dt);
  e->launches += 1;
  if (e->timing) {
    cudaEventRecord(e->ev1, st);
    CUDA_OK(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->kernel_ms += ms;
  }
  return 0;
}

// strides of the two layouts: 0 = (level, column) column-fastest over `ld` columns, 1 = (column, level) level-fastest
In make_in(int nlev, int layout, size_t ld, const cb200_emanuel_inputs& p, int qs_mode) {
  const size_t L = (size_t)nlev;
  In in{};
  in.nlev = nlev;
  if (layout == 0) { in.ls = ld; in.cs = 1; in.ls_i = ld; in.cs_i = 1; }
  else { in.ls = 1; in.cs = L; in.ls_i = 1; in.cs_i = L + 1; }
  in.t = p.t; in.q = p.q; in.u = p.u; in.v = p.v; in.p = p.p; in.ph = p.ph; in.qs = qs_mode == QS_GIVEN ? p.qs : nullptr; in.cbmf = p.cbmf;
  in.qs_mode = qs_mode;
  return in;
}
Out make_out(int nlev, int layout, size_t ld, const cb200_emanuel_outputs& p) {
  Out o{};
  if (layout == 0) { o.ls = ld; o.cs = 1; }
  else { o.ls = 1; o.cs = (size_t)nlev; }
  o.ft = p.ft; o.fq = p.fq; o.fu = p.fu; o.fv = p.fv; o.precip = p.precip; o.wd = p.wd; o.tprime = p.tprime; o.qprime = p.qprime;
  o.cbmf = p.cbmf; o.cape = p.cape; o.iflag = p.iflag;
  return o;
}

}  // namespace

extern "C" int cb200_emanuel_run_device(cb200_emanuel_engine* e, int ncol, int nlev, int max_conv_lev, double dt, int qs_mode, int layout,
                                        const cb200_emanuel_inputs* in, const cb200_emanuel_outputs* out, void* stream) {
  if (int rc = validate(e, ncol, nlev, max_conv_lev, dt, qs_mode, in, out)) return rc;
  if (layout != 0 && layout != 1) { e->error = "emanuel: layout must be 0 (level, column) or 1 (column, level)"; return -3; }
  CUDA_OK(cudaSetDevice(e->device));
  if (int rc = e->ensure_work(nlev, max_conv_lev)) return rc;
  e->launches = 0;
  e->kernel_ms = 0.0;
  if (launch(e, make_in(nlev, layout, (size_t)ncol, *in, qs_mode), make_out(nlev, layout, (size_t)ncol, *out), 0, ncol, max_conv_lev, dt,
             (cudaStream_t)stream))
    return -1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// Host-pointer call in the component's layout ((ncol, nlev) C order): chunked 3-stream pipeline.  A chunk of columns is one
// contiguous block of every array, so H2D / D2H are plain copies and the kernel reads the chunk in place.
extern "C" int cb200_emanuel_run_host(cb200_emanuel_engine* e, int ncol, int nlev, int max_conv_lev, double dt, int qs_mode,
                                      const cb200_emanuel_inputs* hin, const cb200_emanuel_outputs* hout) {
  if (int rc = validate(e, ncol, nlev, max_conv_lev, dt, qs_mode, hin, hout)) return rc;
  CUDA_OK(cudaSetDevice(e->device));
  cb::HostPipe& P = e->pipe;
  CUDA_OK(P.init());
  int chunk = P.chunk * 4;  // convection moves 15x fewer bytes per column than radiation: larger chunks keep the grid filled
  if (chunk > ncol) chunk = ncol;
  if (int rc = e->ensure_work(nlev, max_conv_lev)) return rc;
  const size_t L = (size_t)nlev;
  const size_t in_per_col = 6 * L + 1 + 1 + (qs_mode == QS_GIVEN ? L : 0), out_per_col = 4 * L + 6;
  CUDA_OK(P.ensure(in_per_col * chunk, out_per_col * chunk));
  if (chunk > e->iflag_cap) {
    for (int i = 0; i < 2; ++i) { cudaFree(e->d_iflag[i]); e->d_iflag[i] = nullptr; }
    for (int i = 0; i < 2; ++i) CUDA_OK(cudaMalloc(&e->d_iflag[i], sizeof(int32_t) * chunk));
    e->iflag_cap = chunk;
  }
  e->launches = 0;
  e->kernel_ms = 0.0;
  int k = 0;
  for (int c0 = 0; c0 < ncol; c0 += chunk, ++k) {
    const size_t n = (size_t)((ncol - c0) < chunk ? (ncol - c0) : chunk);
    const int s = k & 1;
    CUDA_OK(cudaStreamWaitEvent(P.s_in, P.cmp_done[s], 0));
    double* d = P.d_in[s];
    cb200_emanuel_inputs di{};
    const double* hsrc[8] = {hin->t, hin->q, hin->u, hin->v, hin->p, hin->ph, hin->qs, hin->cbmf};
    const size_t per[8] = {L, L, L, L, L, L + 1, qs_mode == QS_GIVEN ? L : 0, 1};
    const double** dptr = reinterpret_cast<const double**>(&di);
    for (int i = 0; i < 8; ++i) {
      if (!per[i]) { dptr[i] = nullptr; continue; }
      CUDA_OK(cudaMemcpyAsync(d, hsrc[i] + (size_t)c0 * per[i], per[i] * n * sizeof(double), cudaMemcpyHostToDevice, P.s_in));
      dptr[i] = d;
      d += per[i] * n;
    }
    CUDA_OK(cudaEventRecord(P.in_done[s], P.s_in));
    double* o = P.d_out[s];
    cb200_emanuel_outputs dout{o, o + L * n, o + 2 * L * n, o + 3 * L * n, o + 4 * L * n, o + 4 * L * n + n, o + 4 * L * n + 2 * n,
                               o + 4 * L * n + 3 * n, o + 4 * L * n + 4 * n, o + 4 * L * n + 5 * n, e->d_iflag[s]};
    CUDA_OK(cudaStreamWaitEvent(P.s_cmp, P.in_done[s], 0));
    CUDA_OK(cudaStreamWaitEvent(P.s_cmp, P.out_done[s], 0));
    if (launch(e, make_in(nlev, 1, n, di, qs_mode), make_out(nlev, 1, n, dout), 0, (int)n, max_conv_lev, dt, P.s_cmp)) return -1;
    CUDA_OK(cudaEventRecord(P.cmp_done[s], P.s_cmp));
    CUDA_OK(cudaStreamWaitEvent(P.s_out, P.cmp_done[s], 0));
    double* hdst[10] = {hout->ft, hout->fq, hout->fu, hout->fv, hout->precip, hout->wd, hout->tprime, hout->qprime, hout->cbmf, hout->cape};
    double* const* dsrc = reinterpret_cast<double* const*>(&dout);
    for (int i = 0; i < 10; ++i) {
      const size_t pc = i < 4 ? L : 1;
      CUDA_OK(cudaMemcpyAsync(hdst[i] + (size_t)c0 * pc, dsrc[i], pc * n * sizeof(double), cudaMemcpyDeviceToHost, P.s_out));
    }
    CUDA_OK(cudaMemcpyAsync(hout->iflag + c0, e->d_iflag[s], n * sizeof(int32_t), cudaMemcpyDeviceToHost, P.s_out));
    CUDA_OK(cudaEventRecord(P.out_done[s], P.s_out));
  }
  CUDA_OK(cudaStreamSynchronize(P.s_out));
  CUDA_OK(cudaGetLastError());
  return 0;
}

static_assert(sizeof(cb200_emanuel_inputs) == 8 * sizeof(double*), "cb200_emanuel_inputs layout");
static_assert(sizeof(cb200_emanuel_outputs) == 11 * sizeof(double*), "cb200_emanuel_outputs layout");

// ---- the reference's own C symbols (convect43c.f90:91-137 bind(c) 'init_emanuel_convection_fortran', :146-150 'emanuel_convection'):
// process-global parameters, one column per call, everything by pointer, void.
namespace {
cb200_emanuel_engine* g_engine = nullptr;
}

extern "C" void init_emanuel_convection_fortran(int* pbl, int* least_conv_level, double* thresh_water_level, double* crit_temp,
                                                double* entrain_coeff, double* downdraft_frac_area, double* precip_frac_outside_cloud,
                                                double* rain_speed, double* snow_speed, double* rain_evap_coeff, double* snow_evap_coeff,
                                                double* mom_tran_coeff, double* max_neg_temp_pert, double* beta, double* alpha,
                                                double* damp_amp, double* Cpd, double* Cpv, double* Cl, double* gas_const_vapour,
                                                double* gas_const_air, double* lat_heat, double* grav, double* density_water,
                                                double* reference_mass_flux_timescale) {
  if (*pbl != 0) std::fprintf(stderr, "climt_b200: the dry adiabatic adjustment (IPBL != 0) is not provided; climt always passes 0\n");
  cb200_emanuel_params p{(double)*least_conv_level, *thresh_water_level, *crit_temp, *entrain_coeff, *downdraft_frac_area,
                         *precip_frac_outside_cloud, *rain_speed, *snow_speed, *rain_evap_coeff, *snow_evap_coeff, *mom_tran_coeff, *beta,
                         *max_neg_temp_pert, *alpha, *damp_amp, *Cpd, *Cpv, *Cl, *gas_const_vapour, *gas_const_air, *lat_heat, *grav,
                         *density_water, *reference_mass_flux_timescale, 273.0};
  if (g_engine) { cb200_emanuel_destroy(g_engine); g_engine = nullptr; }
  int dev = 0;
  if (const char* d = std::getenv("CLIMT_B200_DEVICE")) dev = std::atoi(d);
  if (cb200_emanuel_create(&g_engine, &p, dev)) std::fprintf(stderr, "climt_b200: init_emanuel_convection_fortran failed: %s\n", cb200_global_error());
}

extern "C" void emanuel_convection(double* temp, double* q, double* qs, double* u, double* v, double* pmid, double* pint, int* nlevs,
                                   int* max_conv_lev, int* num_tracers, double* dt, int* conv_state, double* dtemp, double* dq, double* du,
                                   double* dv, double* precip, double* downdraft_vel_scale, double* downdraft_temp_scale,
                                   double* downdraft_q_scale, double* cloud_base_mass_flux, double* cape, double* tracers,
                                   double* dtracers) {
  (void)tracers; (void)dtracers;
  if (!g_engine) { std::fprintf(stderr, "climt_b200: init_emanuel_convection_fortran has not been called\n"); return; }
  if (*num_tracers != 0) { std::fprintf(stderr, "climt_b200: emanuel_convection: tracers are not provided (climt passes 0)\n"); return; }
  cb200_emanuel_inputs in{temp, q, u, v, pmid, pint, qs, cloud_base_mass_flux};
  int32_t flag = 0;
  cb200_emanuel_outputs out{dtemp, dq, du, dv, precip, downdraft_vel_scale, downdraft_temp_scale, downdraft_q_scale, cloud_base_mass_flux,
                            cape, &flag};
  if (cb200_emanuel_run_host(g_engine, 1, *nlevs, *max_conv_lev, *dt, 0, &in, &out))
    std::fprintf(stderr, "climt_b200: %s\n", cb200_emanuel_last_error(g_engine));
  *conv_state = flag;
}
