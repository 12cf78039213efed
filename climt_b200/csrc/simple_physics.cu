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

namespace {
struct Geo {
  int ncol, nlev, order;  // order 0: level 0 is the surface (climt); 1: level 0 is the model top (the Fortran's own order)
  // storage index of the Fortran's full level k = 1 (top) .. pver (surface) and interface k = 1 (top) .. pver + 1 (surface)
  __device__ size_t lev(int k, int c) const { return (size_t)(order ? k - 1 : nlev - k) * ncol + c; }
  __device__ size_t ifc(int k, int c) const { return (size_t)(order ? k - 1 : nlev + 1 - k) * ncol + c; }
};

__global__ void __launch_bounds__(128) k_simple_physics(const Geo G, const double dtime, const cb200_simple_physics_params P,
                                                         const cb200_simple_physics_inputs in,
                                                         const cb200_simple_physics_outputs out, double* __restrict__ work) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= G.ncol) return;
  const int pver = G.nlev;
  const double gravit = P.gravit, rair = P.rair, cpair = P.cpair, latvap = P.latvap, rh2o = P.rh2o;
  const double epsilo = rair / rh2o, zvir = (rh2o / rair) - 1.0;
  const double T0 = 273.16, e0 = 610.78, v20 = 20.0, p0 = 100000.0;
  const double kappa = rair / cpair;
  const double ps = in.ps[c];
  // height of the lowest full level, from the state before any process (:291-294)
  const double t_low0 = in.t[G.lev(pver, c)], q_low0 = in.q[G.lev(pver, c)];
  const double za = rair / gravit * t_low0 * (1.0 + zvir * q_low0) * 0.5 * (log(ps) - log(in.pint[G.ifc(pver, c)]));
  double Tsurf;
  if (P.use_ts_ext == 1) {
    Tsurf = in.ts[c];
  } else if (P.test == 1) {  // SST of the moist baroclinic wave (:304-311), single-precision pi and q0
    const double pi = (double)(4.f * atanf(1.f)), T00 = 288.0, u0 = 35.0, eta0 = 0.252, q0 = (double)0.021f;
    const double latw = 2.0 * pi / 9.0, etav = (1.0 - eta0) * 0.5 * pi;
    const double lat = in.lat[c], sl = sin(lat), cl = cos(lat);
    const double sl2 = sl * sl, sl6 = sl2 * sl2 * sl2, r = lat / latw, r2 = r * r;
    Tsurf = (T00 + pi * u0 / rair * 1.5 * sin(etav) * pow(cos(etav), 0.5) *
                       ((-2.0 * sl6 * (cl * cl + 1.0 / 3.0) + 10.0 / 63.0) * u0 * pow(cos(etav), 1.5) +
                        (8.0 / 5.0 * (cl * cl * cl) * (sl2 + 2.0 / 3.0) - pi / 4.0) * P.radius * P.omega * 0.5)) /
            (1.0 + zvir * q0 * exp(-(r2 * r2)));
  } else {
    Tsurf = 302.15;
  }
  // ---- large-scale condensation (:330-352); the updated T and q go to the output arrays
  double precl = 0.0;
  for (int k = 1; k <= pver; ++k) {
    const size_t o = G.lev(k, c);
    double t = in.t[o], q = in.q[o];
    if (P.do_lsc == 1) {
      const double qsat = epsilo * e0 / in.pmid[o] * exp(-latvap / rh2o * ((1.0 / t) - 1.0 / T0));
      double dtdt = 0.0, dqdt = 0.0;
      if (q > qsat) {
        const double tmp = 1.0 / dtime * (q - qsat) / (1.0 + (latvap / cpair) * (epsilo * latvap * qsat / (rair * (t * t))));
        dtdt = latvap / cpair * tmp;
        dqdt = -tmp;
        precl = precl + tmp * (in.pint[G.ifc(k + 1, c)] - in.pint[G.ifc(k, c)]) / (gravit * P.rhow);
      }
      t = t + dtdt * dtime;
      q = q + dqdt * dtime;
    }
    out.t[o] = t;
    out.q[o] = q;
    out.u[o] = in.u[o];
    out.v[o] = in.v[o];
  }
  out.precl[c] = precl;
  // ---- surface fluxes (:358-430)
  double Km_s = 0.0, Ke_s = 0.0, sens = 0.0, lath = 0.0;
  if (P.do_surf_flux == 1) {
    const size_t o = G.lev(pver, c);
    const double u = out.u[o], v = out.v[o];
    const double wind = sqrt(u * u + v * v);
    Ke_s = P.C * wind * za;
    double Cd;
    if (wind < v20) {
      Cd = P.Cd0 + P.Cd1 * wind;
      Km_s = Cd * wind * za;
    } else {
      Cd = P.Cm;
      Km_s = P.Cm * wind * za;
    }
    const double damp = 1.0 + Cd * wind * dtime / za;
    out.u[o] = u / damp;
    out.v[o] = v / damp;
    const double pm = in.pmid[o], dps = in.pint[G.ifc(pver + 1, c)] - in.pint[G.ifc(pver, c)];
    double t = out.t[o], q = out.q[o];
    double rho = pm / (rair * t);
    double flux = P.C * wind * (Tsurf - t);
    sens = rho * cpair * flux;
    t = t + (flux * (rho * gravit) / dps) * dtime;
    double qsats;
    if (P.use_qsurf_ext == 1) {
      qsats = in.qsurf[c];
    } else {
      double esats;
      if (Tsurf > 271) {
        esats = ((double)1.0007f + (double)3.46e-8f * ps) * (double)611.21f *
                exp((double)17.966f * (Tsurf - 273.) / ((double)247.15f + (Tsurf - 273.)));
      } else {
        esats = ((double)1.0003f + (double)4.18e-8f * ps) * (double)611.15f *
                exp((double)22.452f * (Tsurf - 273.) / ((double)272.5f + (Tsurf - 273.)));
      }
      qsats = epsilo * esats / (ps - (double)0.378f * esats);
    }
    rho = pm / (rair * t);
    flux = P.C * wind * (qsats - q);
    lath = latvap * rho * flux;
    q = q + (flux * (rho * gravit) / dps) * dtime;
    out.t[o] = t;
    out.q[o] = q;
  }
  out.sens_ht_flux[c] = sens;
  out.lat_ht_flux[c] = (P.clamp_latent_heat_flux && lath < 0.0) ? 0.0 : lath;  // component.py:257
  if (P.do_pbl != 1) return;
  // ---- boundary layer: implicit diffusion of u, v, theta, q (:436-520).  Forward sweep k = pver .. 1 (surface -> top): the
  // coefficients of level k need the interface densities above (k, k-1) and below (k+1, k); CE / CEm go to the workspace, the
  // four CF right-hand sides overwrite the output arrays (each level's state is consumed before it is overwritten).
  const size_t wrow = (size_t)pver * G.ncol;
  double* __restrict__ wCE = work;
  double* __restrict__ wCEm = work + wrow;
  const double pc2 = P.pblconst * P.pblconst;
  auto taper = [&](int k) {  // Km(k) / Km(pver + 1) at interface k (:368-380)
    const double pk = in.pint[G.ifc(k, c)];
    return pk >= P.pbltop ? 1.0 : exp(-((P.pbltop - pk) * (P.pbltop - pk)) / pc2);
  };
  double CE_b = 0.0, CEm_b = 0.0, CFu_b = 0.0, CFv_b = 0.0, CFt_b = 0.0, CFq_b = 0.0;  // values at k + 1
  double t_k = out.t[G.lev(pver, c)], pm_k = in.pmid[G.lev(pver, c)];
  double CA = 0.0, CAm = 0.0;  // CA(pver) = CAm(pver) = 0
  for (int k = pver; k >= 1; --k) {
    const size_t o = G.lev(k, c);
    const double rpdel = 1.0 / (in.pint[G.ifc(k + 1, c)] - in.pint[G.ifc(k, c)]);
    double CC = 0.0, CCm = 0.0, CA_up = 0.0, CAm_up = 0.0, t_up = 0.0, pm_up = 0.0;
    if (k > 1) {  // interface k, between levels k-1 and k: CC(k), CCm(k) and the CA(k-1), CAm(k-1) of the level above
      const size_t ou = G.lev(k - 1, c);
      t_up = out.t[ou];
      pm_up = in.pmid[ou];
      const double rho = in.pint[G.ifc(k, c)] / (rair * (t_k + t_up) / 2.0);
      const double tp = taper(k), Km = Km_s * tp, Ke = Ke_s * tp;
      const double dpm = pm_k - pm_up;
      CCm = rpdel * dtime * gravit * gravit * Km * rho * rho / dpm;
      CC = rpdel * dtime * gravit * gravit * Ke * rho * rho / dpm;
      const double rpdel_up = 1.0 / (in.pint[G.ifc(k, c)] - in.pint[G.ifc(k - 1, c)]);
      CAm_up = rpdel_up * dtime * gravit * gravit * Km * rho * rho / dpm;
      CA_up = rpdel_up * dtime * gravit * gravit * Ke * rho * rho / dpm;
    }
    const double den = 1.0 + CA + CC - CA * CE_b, denm = 1.0 + CAm + CCm - CAm * CEm_b;
    const double CE = CC / den, CEm = CCm / denm;
    const double CFu = (out.u[o] + CAm * CFu_b) / denm, CFv = (out.v[o] + CAm * CFv_b) / denm;
    const double CFt = (pow(p0 / pm_k, kappa) * t_k + CA * CFt_b) / den, CFq = (out.q[o] + CA * CFq_b) / den;
    wCE[o] = CE;
    wCEm[o] = CEm;
    out.u[o] = CFu; out.v[o] = CFv; out.t[o] = CFt; out.q[o] = CFq;
    CE_b = CE; CEm_b = CEm; CFu_b = CFu; CFv_b = CFv; CFt_b = CFt; CFq_b = CFq;
    CA = CA_up; CAm = CAm_up; t_k = t_up; pm_k = pm_up;
  }
  // back substitution k = 1 .. pver (top -> surface) (:498-520)
  double u_a, v_a, t_a, q_a, pw_a;  // new values at k - 1 and (p0 / pmid(k-1))^kappa
  {
    const size_t o = G.lev(1, c);
    const double pm = in.pmid[o];
    u_a = out.u[o]; v_a = out.v[o]; q_a = out.q[o];
    t_a = out.t[o] * pow(pm / p0, kappa);
    out.t[o] = t_a;
    pw_a = pow(p0 / pm, kappa);
  }
  for (int k = 2; k <= pver; ++k) {
    const size_t o = G.lev(k, c);
    const double CE = wCE[o], CEm = wCEm[o], pm = in.pmid[o];
    const double u_n = CEm * u_a + out.u[o], v_n = CEm * v_a + out.v[o];
    const double t_n = (CE * t_a * pw_a + out.t[o]) * pow(pm / p0, kappa);
    const double q_n = CE * q_a + out.q[o];
    out.u[o] = u_n; out.v[o] = v_n; out.t[o] = t_n; out.q[o] = q_n;
    u_a = u_n; v_a = v_n; q_a = q_n; t_a = t_n;
    pw_a = pow(p0 / pm, kappa);
  }
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
