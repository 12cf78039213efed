// climt_b200 -- Reed-Jablonowski simple physics: the per-column code of k_simple_physics (also compiled for the host by
// tests/emul/simple_physics_emul.cpp, which steps it column by column against the oracle).  See simple_physics.cu for what it
// replaces and which quirks of the Fortran are kept.
#pragma once
#include <math.h>

#include "../../include/climt_b200.h"
#include "cb_common.h"

namespace cb {
namespace sp {
struct Geo {
  int ncol, nlev, order;  // order 0: level 0 is the surface (climt); 1: level 0 is the model top (the Fortran's own order)
  // storage index of the Fortran's full level k = 1 (top) .. pver (surface) and interface k = 1 (top) .. pver + 1 (surface)
  CB_HD size_t lev(int k, int c) const { return (size_t)(order ? k - 1 : nlev - k) * ncol + c; }
  CB_HD size_t ifc(int k, int c) const { return (size_t)(order ? k - 1 : nlev + 1 - k) * ncol + c; }
};

CB_HD void simple_physics_column(const Geo& G, const double dtime, const cb200_simple_physics_params& P,
                                 const cb200_simple_physics_inputs& in, const cb200_simple_physics_outputs& out,
                                 double* __restrict__ work, const int c) {
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

}  // namespace sp
}  // namespace cb
