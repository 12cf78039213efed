// climt_b200 -- Reed-Jablonowski simple physics: large-scale condensation, bulk surface fluxes and the
// implicit boundary-layer diffusion of one column, sm_100a.
// Replaces
//   simple_physics_func / set_physical_constants_func   climt/_lib/simple_physics/simple_physics_custom.f90:59-565, 28-57
//   get_new_state / do_simple_physics (level flip, pdel) climt/_components/simple_physics/_simple_physics.pyx:84-180
// One thread per column, columns fastest: every access of a warp is one contiguous 256-byte row.  A column is three short
// serial passes over the levels (condensation; Thomas forward sweep surface -> top; back substitution top -> surface); the four
// right-hand sides of the tridiagonal systems are parked in the output arrays and the two elimination factors in a caller-provided
// workspace, so nothing is thread-local beyond scalars.  ~10 KB per 60-level column: HBM streaming.
//
// Kept from the Fortran on purpose: its un-suffixed literals are single precision (1.0007, 3.46e-8, 611.21, 17.966, 273., 247.15,
// 0.378, 1.0003, 4.18e-8, 611.15, 22.452, 272.5, q0 = 0.021, pi = 4.*atan(1.)); the `test` switch is fed from the component's
// `simulate_cyclone` flag (_simple_physics.pyx:168) and selects the baroclinic-wave SST when 1; latitude arrives in degrees.
// Not reproduced: boundary layer without surface fluxes -- the Fortran then reads Km / Ke uninitialised (:368-381 skipped, :441-452
// read them); the entry points reject that combination.
#include <cuda_runtime.h>
#include <math.h>

#include <cstdio>
#include <cstdlib>

#include <string>

#include "../../include/climt_b200.h"
#include "engine_common.h"
#include "simple_physics_core.cuh"

namespace {
using cb::sp::Geo;

__global__ void __launch_bounds__(128) k_simple_physics(const Geo G, const double dtime, const cb200_simple_physics_params P,
                                                         const cb200_simple_physics_inputs in,
                                                         const cb200_simple_physics_outputs out, double* __restrict__ work) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < G.ncol) cb::sp::simple_physics_column(G, dtime, P, in, out, work, c);
}

int fail(const std::string& s) {
  cb::set_global_error(s);
  return -1;
}

int check(int ncol, int nlev, double dtime, const cb200_simple_physics_params* p) {
  if (ncol <= 0 || nlev < 2 || !(dtime > 0.0)) { cb::set_global_error("simple physics: bad ncol / nlev / dtime"); return -3; }
  if (p->do_pbl == 1 && p->do_surf_flux != 1) {
    cb::set_global_error("simple physics: boundary_layer without surface_fluxes reads uninitialised diffusivities in the reference "
                         "(simple_physics_custom.f90:368-381 is skipped, :441-452 reads Km / Ke); not provided");
    return -3;
  }
  return 0;
}

cb200_simple_physics_params g_params = {};  // the reference-named entry points' process-global configuration
}  // namespace

extern "C" int cb200_simple_physics_run_device(int device, int ncol, int nlev, int order, double dtime,
                                               const cb200_simple_physics_params* p, const cb200_simple_physics_inputs* in,
                                               const cb200_simple_physics_outputs* out, double* workspace, void* stream) {
  if (int rc = check(ncol, nlev, dtime, p)) return rc;
  if (p->do_pbl == 1 && !workspace) { cb::set_global_error("simple physics: the boundary layer needs a 2 * nlev * ncol workspace"); return -3; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  const Geo G{ncol, nlev, order ? 1 : 0};
  k_simple_physics<<<(ncol + 127) / 128, 128, 0, (cudaStream_t)stream>>>(G, dtime, *p, *in, *out, workspace);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail(cudaGetErrorString(e));
}

extern "C" int cb200_simple_physics_run_host(int device, int ncol, int nlev, int order, double dtime,
                                             const cb200_simple_physics_params* p, const cb200_simple_physics_inputs* in,
                                             const cb200_simple_physics_outputs* out) {
  if (int rc = check(ncol, nlev, dtime, p)) return rc;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  const size_t n = (size_t)ncol, L = (size_t)nlev;
  // t q u v pmid (L) | pint (L+1) | ps ts qsurf lat (1) || t q u v (L) | precl sens lat (1) || workspace 2L
  const size_t tot = (5 * L + (L + 1) + 4 + 4 * L + 3 + 2 * L) * n;
  double* d = nullptr;
  if ((e = cudaMalloc(&d, tot * sizeof(double))) != cudaSuccess) return fail(cudaGetErrorString(e));
  cb200_simple_physics_inputs di;
  cb200_simple_physics_outputs dout;
  double* q = d;
  auto take = [&](size_t rows) { double* r = q; q += rows * n; return r; };
  const double* src[10] = {in->t, in->q, in->u, in->v, in->pmid, in->pint, in->ps, in->ts, in->qsurf, in->lat};
  const size_t rows[10] = {L, L, L, L, L, L + 1, 1, 1, 1, 1};
  const double** dst[10] = {&di.t, &di.q, &di.u, &di.v, &di.pmid, &di.pint, &di.ps, &di.ts, &di.qsurf, &di.lat};
  for (int i = 0; i < 10; ++i) {
    double* r = take(rows[i]);
    *dst[i] = r;
    if (src[i]) cudaMemcpyAsync(r, src[i], rows[i] * n * sizeof(double), cudaMemcpyHostToDevice, 0);
    else cudaMemsetAsync(r, 0, rows[i] * n * sizeof(double), 0);
  }
  dout.t = take(L); dout.q = take(L); dout.u = take(L); dout.v = take(L);
  dout.precl = take(1); dout.sens_ht_flux = take(1); dout.lat_ht_flux = take(1);
  double* work = take(2 * L);
  int rc = cb200_simple_physics_run_device(device, ncol, nlev, order, dtime, p, &di, &dout, work, nullptr);
  if (rc == 0) {
    double* hdst[7] = {out->t, out->q, out->u, out->v, out->precl, out->sens_ht_flux, out->lat_ht_flux};
    const double* hsrc[7] = {dout.t, dout.q, dout.u, dout.v, dout.precl, dout.sens_ht_flux, dout.lat_ht_flux};
    const size_t hrows[7] = {L, L, L, L, 1, 1, 1};
    for (int i = 0; i < 7; ++i) cudaMemcpyAsync(hdst[i], hsrc[i], hrows[i] * n * sizeof(double), cudaMemcpyDeviceToHost, 0);
    if ((e = cudaStreamSynchronize(0)) != cudaSuccess) rc = fail(cudaGetErrorString(e));
  }
  cudaFree(d);
  return rc;
}

// ---- the reference's own symbols (bind(c) names of simple_physics_custom.f90:30-32, 59-63; declared in _simple_physics.pyx:6-27):
// process-global constants, arrays (pver, pcols) C order = Fortran (pcols, pver), model top first, t / q / u / v updated in place.
extern "C" void set_fortran_constants(double* g, double* cpd, double* r_air, double* latent_heat, double* r_cond, double* radius,
                                      double* rotation, double* density_cond, double* top_pbl, double* pbl_decay,
                                      double* drag_coeff_sens_lat, double* Cd0_ext, double* Cd1_ext, double* Cm_ext) {
  g_params.gravit = *g; g_params.cpair = *cpd; g_params.rair = *r_air; g_params.latvap = *latent_heat; g_params.rh2o = *r_cond;
  g_params.radius = *radius; g_params.omega = *rotation; g_params.rhow = *density_cond; g_params.pbltop = *top_pbl;
  g_params.pblconst = *pbl_decay; g_params.C = *drag_coeff_sens_lat; g_params.Cd0 = *Cd0_ext; g_params.Cd1 = *Cd1_ext;
  g_params.Cm = *Cm_ext;
}

extern "C" void simple_physics(int* pcols, int* pver, double* dtime, double* lat, double* t, double* q, double* u, double* v,
                               double* pmid, double* pint, double* pdel, double* rpdel, double* ps, double* precl, int* test,
                               int* do_lsc, int* do_pbl, int* do_surf_flux, int* use_ts_ext, double* ts, int* use_qsurf_ext,
                               double* qsurf, double* sens_ht_flux, double* lat_ht_flux) {
  (void)pdel; (void)rpdel;  // the layer thickness is re-formed from pint exactly as _simple_physics.pyx:113-116 formed it
  cb200_simple_physics_params p = g_params;
  p.test = *test; p.do_lsc = *do_lsc; p.do_pbl = *do_pbl; p.do_surf_flux = *do_surf_flux; p.use_ts_ext = *use_ts_ext;
  p.use_qsurf_ext = *use_qsurf_ext; p.clamp_latent_heat_flux = 0;
  const cb200_simple_physics_inputs in{t, q, u, v, pmid, pint, ps, ts, qsurf, lat};
  const cb200_simple_physics_outputs out{t, q, u, v, precl, sens_ht_flux, lat_ht_flux};
  int dev = 0;
  if (const char* d = std::getenv("CLIMT_B200_DEVICE")) dev = std::atoi(d);
  if (cb200_simple_physics_run_host(dev, *pcols, *pver, 1, *dtime, &p, &in, &out))
    std::fprintf(stderr, "climt_b200: simple_physics failed: %s\n", cb200_global_error());
}
