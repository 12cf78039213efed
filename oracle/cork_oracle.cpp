// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's CORK correlated-k radiation kernels.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// Restates (file:line of the reference each function follows):
//   ck_bracket            cork/optics/correlated_k.py:30-43   (np.searchsorted(side="left") - 1, clamped; fraction in [0,1])
//   ck_txx / ck_txx_cont  cork/optics/correlated_k.py:46-78   (trilinear in T, log p, log X_H2O; term order kept)
//   orc_cork_tau          cork/optics/correlated_k.py:81-117  (7-D table: CO2 axis, geometric interpolation in k)
//                         cork/optics/correlated_k.py:222-330, 420-470 (6-D / 5-D tables via interpolate_k + continuum)
//   orc_cork_planck       cork/lw/kernels.py:9-68
//   orc_cork_lw_transport cork/lw/kernels.py:71-121
//   orc_cork_sw_two_stream cork/sw/kernels.py:18-263 (_delta_scale, _sw_dif_and_source, _adding, _sw_two_stream_core)
//   orc_cork_column_amount / orc_cork_heating  cork/common.py:38-80
// Pinned against tests/golden/cork_reference.npz (outputs of the reference's numba kernels run in the build container,
// tests/golden/make_cork_golden.py).  The reference evaluates exp/log with libm, as this file does.
#include <cmath>
#include <cstddef>
#include <vector>

namespace {

struct Br { int i; double f; };

inline Br bracket(const double* grid, int n, double v) {
  // np.searchsorted(grid, v) (side='left'): first index with grid[idx] >= v
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    if (grid[mid] < v) lo = mid + 1; else hi = mid;
  }
  int i = lo - 1;
  if (i < 0) i = 0;
  else if (i > n - 2) i = n - 2;
  double f = (v - grid[i]) / (grid[i + 1] - grid[i]);
  if (f < 0.0) f = 0.0;
  else if (f > 1.0) f = 1.0;
  return {i, f};
}

}  // namespace

extern "C" {

// tau[b][g][lev][col].  nX == 0: 5-D table (T, P); nC == 0: no CO2 axis.  k in the table's own dtype (float32 entries are
// promoted on use, exactly as numba/numpy promote them); planck_fraction is handed over promoted to double.
void orc_cork_tau(const void* kptr, int k_is_f64, int ngas, int nband, int ngpt, int nT, int nP, int nX, int nC, const double* T_grid,
                  const double* p_grid_log, const double* log_x_grid, const double* log_c_grid, const double* T,
                  const double* log_p, const double* log_x, const double* log_c, const double* gas_amounts, int has_cont,
                  const double* log_cont, int co2_logk, int nlev, int ncol, double* tau) {
  const double FLOOR = 1e-40;
  const int nXe = nX > 0 ? nX : 1, nCe = nC > 0 ? nC : 1;
  auto K = [&](int ig, int ib, int igp, int iT, int iP, int iX, int iC) -> double {
    const size_t o = ((((((size_t)ig * nband + ib) * ngpt + igp) * nT + iT) * nP + iP) * nXe + iX) * nCe + iC;
    return k_is_f64 ? static_cast<const double*>(kptr)[o] : (double)static_cast<const float*>(kptr)[o];
  };
  auto LC = [&](int ib, int iT, int iP, int iX) -> double { return log_cont[(((size_t)ib * nT + iT) * nP + iP) * nXe + iX]; };
  for (int i = 0; i < ncol; ++i)
    for (int kk = 0; kk < nlev; ++kk) {
      const size_t o = (size_t)kk * ncol + i;
      const Br bT = bracket(T_grid, nT, T[o]);
      const Br bP = bracket(p_grid_log, nP, log_p[o]);
      Br bX{0, 0.0}, bC{0, 0.0};
      if (nX > 0) bX = bracket(log_x_grid, nX, log_x[o]);
      if (nC > 0) bC = bracket(log_c_grid, nC, log_c[o]);
      const int iT = bT.i, iP = bP.i, iX = bX.i, iC = bC.i;
      const double fT = bT.f, fP = bP.f, fX = bX.f, fC = bC.f;
      auto txx = [&](int ig, int ib, int igp, int ic) -> double {
        if (nX == 0) {  // bilinear (_tp)
          return K(ig, ib, igp, iT, iP, 0, ic) * (1.0 - fT) * (1.0 - fP) + K(ig, ib, igp, iT + 1, iP, 0, ic) * fT * (1.0 - fP) +
                 K(ig, ib, igp, iT, iP + 1, 0, ic) * (1.0 - fT) * fP + K(ig, ib, igp, iT + 1, iP + 1, 0, ic) * fT * fP;
        }
        const double x0 = K(ig, ib, igp, iT, iP, iX, ic) * (1.0 - fT) * (1.0 - fP) + K(ig, ib, igp, iT + 1, iP, iX, ic) * fT * (1.0 - fP) +
                          K(ig, ib, igp, iT, iP + 1, iX, ic) * (1.0 - fT) * fP + K(ig, ib, igp, iT + 1, iP + 1, iX, ic) * fT * fP;
        const double x1 = K(ig, ib, igp, iT, iP, iX + 1, ic) * (1.0 - fT) * (1.0 - fP) +
                          K(ig, ib, igp, iT + 1, iP, iX + 1, ic) * fT * (1.0 - fP) +
                          K(ig, ib, igp, iT, iP + 1, iX + 1, ic) * (1.0 - fT) * fP + K(ig, ib, igp, iT + 1, iP + 1, iX + 1, ic) * fT * fP;
        return x0 * (1.0 - fX) + x1 * fX;
      };
      for (int ib = 0; ib < nband; ++ib) {
        double cont_val = 0.0;
        if (has_cont) {
          const double x0 = LC(ib, iT, iP, iX) * (1.0 - fT) * (1.0 - fP) + LC(ib, iT + 1, iP, iX) * fT * (1.0 - fP) +
                            LC(ib, iT, iP + 1, iX) * (1.0 - fT) * fP + LC(ib, iT + 1, iP + 1, iX) * fT * fP;
          const double x1 = LC(ib, iT, iP, iX + 1) * (1.0 - fT) * (1.0 - fP) + LC(ib, iT + 1, iP, iX + 1) * fT * (1.0 - fP) +
                            LC(ib, iT, iP + 1, iX + 1) * (1.0 - fT) * fP + LC(ib, iT + 1, iP + 1, iX + 1) * fT * fP;
          cont_val = std::exp(x0 * (1.0 - fX) + x1 * fX);
        }
        for (int igp = 0; igp < ngpt; ++igp) {
          double acc = 0.0;
          for (int ig = 0; ig < ngas; ++ig) {
            double kv;
            if (nC > 0) {
              const double c0 = txx(ig, ib, igp, iC), c1 = txx(ig, ib, igp, iC + 1);
              if (co2_logk) {
                const double l0 = std::log(c0 > FLOOR ? c0 : FLOOR), l1 = std::log(c1 > FLOOR ? c1 : FLOOR);
                kv = std::exp(l0 * (1.0 - fC) + l1 * fC);
              } else {
                kv = c0 * (1.0 - fC) + c1 * fC;
              }
            } else {
              kv = txx(ig, ib, igp, 0);
            }
            acc += kv * gas_amounts[((size_t)ig * nlev + kk) * ncol + i];
          }
          if (has_cont) acc += cont_val * gas_amounts[(size_t)kk * ncol + i];
          tau[(((size_t)ib * ngpt + igp) * nlev + kk) * ncol + i] = acc;
        }
      }
    }
}

void orc_cork_planck(const double* planck_frac, int nband_orig, int ngpt_orig, int nT, const double* T_grid, const double* T,
                     const double* T_surf, double sigma, int nband, int ngpt, int is_esft, int nlev, int ncol,
                     double* planck_src, double* surf_src) {
  auto PF = [&](int b, int g, int t) -> double { return (double)planck_frac[((size_t)b * ngpt_orig + g) * nT + t]; };
  for (int icol = 0; icol < ncol; ++icol) {
    const double T_s = T_surf[icol];
    const Br bs = bracket(T_grid, nT, T_s);
    const double surf_planck = sigma * ((T_s * T_s) * (T_s * T_s));  // numba lowers `** 4` to powi: (x^2)^2
    for (int ib = 0; ib < nband; ++ib) {
      const int ibo = ib < nband_orig ? ib : nband_orig - 1;
      for (int igp = 0; igp < ngpt; ++igp) {
        const int go = is_esft ? igp % ngpt_orig : igp;
        const double frac = PF(ibo, go, bs.i) * (1.0 - bs.f) + PF(ibo, go, bs.i + 1) * bs.f;
        surf_src[((size_t)ib * ngpt + igp) * ncol + icol] = frac * surf_planck;
      }
    }
    for (int kk = 0; kk < nlev; ++kk) {
      const double T_l = T[(size_t)kk * ncol + icol];
      const Br bl = bracket(T_grid, nT, T_l);
      const double layer_planck = sigma * ((T_l * T_l) * (T_l * T_l));
      for (int ib = 0; ib < nband; ++ib) {
        const int ibo = ib < nband_orig ? ib : nband_orig - 1;
        for (int igp = 0; igp < ngpt; ++igp) {
          const int go = is_esft ? igp % ngpt_orig : igp;
          const double frac = PF(ibo, go, bl.i) * (1.0 - bl.f) + PF(ibo, go, bl.i + 1) * bl.f;
          planck_src[(((size_t)ib * ngpt + igp) * nlev + kk) * ncol + icol] = frac * layer_planck;
        }
      }
    }
  }
}

void orc_cork_lw_transport(const double* tau, const double* planck_source, const double* surface_source, const double* emissivity,
                           const double* weights, int nband, int ngpt, int nlev, int ncol, double D, double* up_band,
                           double* down_band, double* up_broad, double* down_broad) {
  auto I4 = [&](int b, int g, int k, int i) { return (((size_t)b * ngpt + g) * nlev + k) * ncol + i; };
  auto IB = [&](int b, int k, int i) { return ((size_t)b * (nlev + 1) + k) * ncol + i; };
  for (int i = 0; i < ncol; ++i) {
    for (int k = 0; k <= nlev; ++k) { up_broad[(size_t)k * ncol + i] = 0.0; down_broad[(size_t)k * ncol + i] = 0.0; }
    for (int b = 0; b < nband; ++b) {
      for (int k = 0; k <= nlev; ++k) { up_band[IB(b, k, i)] = 0.0; down_band[IB(b, k, i)] = 0.0; }
      for (int g = 0; g < ngpt; ++g) {
        const double w = weights[(size_t)b * ngpt + g];
        double up_prev = emissivity[(size_t)b * ncol + i] * surface_source[((size_t)b * ngpt + g) * ncol + i];
        up_band[IB(b, 0, i)] += w * up_prev;
        for (int k = 0; k < nlev; ++k) {
          const double trans = std::exp(-D * tau[I4(b, g, k, i)]);
          const double up_cur = up_prev * trans + planck_source[I4(b, g, k, i)] * (1.0 - trans);
          up_band[IB(b, k + 1, i)] += w * up_cur;
          up_prev = up_cur;
        }
        double dn_prev = 0.0;
        for (int k = nlev - 1; k >= 0; --k) {
          const double trans = std::exp(-D * tau[I4(b, g, k, i)]);
          const double dn_cur = dn_prev * trans + planck_source[I4(b, g, k, i)] * (1.0 - trans);
          down_band[IB(b, k, i)] += w * dn_cur;
          dn_prev = dn_cur;
        }
      }
      for (int k = 0; k <= nlev; ++k) {
        up_broad[(size_t)k * ncol + i] += up_band[IB(b, k, i)];
        down_broad[(size_t)k * ncol + i] += down_band[IB(b, k, i)];
      }
    }
  }
}

void orc_cork_sw_two_stream(const double* tau, const double* ssa, const double* asym, const double* zenith, const double* albedo,
                            const double* solar_flux, const double* weights, int nband, int ngpt, int nlev, int ncol,
                            double* up_band, double* down_band, double* up_broad, double* down_broad) {
  const double MIN_K = 1.0e-12, MIN_MU0 = 1.0e-8;
  auto I4 = [&](int b, int g, int k, int i) { return (((size_t)b * ngpt + g) * nlev + k) * ncol + i; };
  auto IB = [&](int b, int k, int i) { return ((size_t)b * (nlev + 1) + k) * ncol + i; };
  for (size_t j = 0; j < (size_t)nband * (nlev + 1) * ncol; ++j) { up_band[j] = 0.0; down_band[j] = 0.0; }
  std::vector<double> Rdif(nlev), Tdif(nlev), src_up(nlev), src_dn(nlev), flux_dn_dir(nlev + 1), alb(nlev + 1), src(nlev + 1),
      denom(nlev), flux_up(nlev + 1), flux_dn(nlev + 1);
  for (int b = 0; b < nband; ++b)
    for (int g = 0; g < ngpt; ++g) {
      const double w = weights[(size_t)b * ngpt + g];
      for (int i = 0; i < ncol; ++i) {
        const double mu0 = std::cos(zenith[i]);
        if (mu0 <= 1e-4) continue;  // night
        flux_dn_dir[nlev] = 1.0;
        for (int k = nlev - 1; k >= 0; --k) {
          // _delta_scale
          const double t0 = tau[I4(b, g, k, i)], s0 = ssa[I4(b, g, k, i)], g0 = asym[I4(b, g, k, i)];
          const double f = g0 * g0;
          const double tau_s = t0 * (1.0 - s0 * f);
          const double w0 = (1.0 - s0 * f) > 1e-30 ? s0 * (1.0 - f) / (1.0 - s0 * f) : 0.0;
          const double gs = (1.0 - f) > 1e-30 ? (g0 - f) / (1.0 - f) : 0.0;
          // _sw_dif_and_source
          const double gamma1 = (8.0 - w0 * (5.0 + 3.0 * gs)) * 0.25;
          const double gamma2 = 3.0 * (w0 * (1.0 - gs)) * 0.25;
          const double kk = std::sqrt(std::fmax((gamma1 - gamma2) * (gamma1 + gamma2), MIN_K));
          const double e1 = std::exp(-tau_s * kk);
          const double e2 = e1 * e1;
          const double RT = 1.0 / (kk * (1.0 + e2) + gamma1 * (1.0 - e2));
          const double rdif = RT * gamma2 * (1.0 - e2);
          const double tdif = RT * 2.0 * kk * e1;
          const double mu0_s = std::fmax(mu0, MIN_MU0);
          const double Tnoscat = std::exp(-tau_s / mu0_s);
          const double k_mu = kk * mu0_s;
          double denom_dir = 1.0 - k_mu * k_mu;
          if (std::fabs(denom_dir) < 1e-30) denom_dir = 1e-30;
          const double RTd = w0 * RT / denom_dir;
          const double gamma3 = (2.0 - 3.0 * mu0_s * gs) * 0.25;
          const double gamma4 = 1.0 - gamma3;
          const double alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
          const double alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
          const double k_g3 = kk * gamma3, k_g4 = kk * gamma4;
          double Rdir = RTd * ((1.0 - k_mu) * (alpha2 + k_g3) - (1.0 + k_mu) * (alpha2 - k_g3) * e2 -
                               2.0 * (k_g3 - alpha2 * k_mu) * e1 * Tnoscat);
          double Tdir = -RTd * ((1.0 + k_mu) * (alpha1 + k_g4) * Tnoscat - (1.0 - k_mu) * (alpha1 - k_g4) * e2 * Tnoscat -
                                2.0 * (k_g4 + alpha1 * k_mu) * e1);
          Rdir = std::fmax(0.0, std::fmin(Rdir, 1.0 - Tnoscat));
          Tdir = std::fmax(0.0, std::fmin(Tdir, 1.0 - Tnoscat - Rdir));
          Rdif[k] = rdif; Tdif[k] = tdif;
          flux_dn_dir[k] = Tnoscat * flux_dn_dir[k + 1];
          src_up[k] = Rdir * flux_dn_dir[k + 1];
          src_dn[k] = Tdir * flux_dn_dir[k + 1];
        }
        const double src_sfc = flux_dn_dir[0] * albedo[i];
        // _adding
        alb[0] = albedo[i];
        src[0] = src_sfc;
        for (int k = 0; k < nlev; ++k) {
          denom[k] = 1.0 / (1.0 - Rdif[k] * alb[k]);
          alb[k + 1] = Rdif[k] + Tdif[k] * Tdif[k] * alb[k] * denom[k];
          src[k + 1] = src_up[k] + Tdif[k] * denom[k] * (src[k] + alb[k] * src_dn[k]);
        }
        flux_dn[nlev] = 0.0;
        flux_up[nlev] = flux_dn[nlev] * alb[nlev] + src[nlev];
        for (int k = nlev - 1; k >= 0; --k) {
          flux_dn[k] = (Tdif[k] * flux_dn[k + 1] + Rdif[k] * src[k] + src_dn[k]) * denom[k];
          flux_up[k] = flux_dn[k] * alb[k] + src[k];
        }
        const double scale = solar_flux[(size_t)b * ngpt + g] * mu0 * w;
        for (int k = 0; k <= nlev; ++k) {
          up_band[IB(b, k, i)] += flux_up[k] * scale;
          down_band[IB(b, k, i)] += (flux_dn_dir[k] + flux_dn[k]) * scale;
        }
      }
    }
  for (size_t j = 0; j < (size_t)(nlev + 1) * ncol; ++j) { up_broad[j] = 0.0; down_broad[j] = 0.0; }
  for (int b = 0; b < nband; ++b)
    for (int k = 0; k <= nlev; ++k)
      for (int i = 0; i < ncol; ++i) {
        up_broad[(size_t)k * ncol + i] += up_band[IB(b, k, i)];
        down_broad[(size_t)k * ncol + i] += down_band[IB(b, k, i)];
      }
}

void orc_cork_heating(const double* net_flux, const double* p_interface, double g, double cpd, int nlev, int ncol, double* hr) {
  for (int i = 0; i < ncol; ++i)
    for (int k = 0; k < nlev; ++k) {
      const double dp = p_interface[(size_t)(k + 1) * ncol + i] - p_interface[(size_t)k * ncol + i];
      const double dflux = net_flux[(size_t)(k + 1) * ncol + i] - net_flux[(size_t)k * ncol + i];
      hr[(size_t)k * ncol + i] = g / cpd * dflux / dp;
    }
}

void orc_cork_column_amount(const double* q, const double* p_interface, double g, int nlev, int ncol, double* amount) {
  for (int i = 0; i < ncol; ++i)
    for (int k = 0; k < nlev; ++k) {
      const double dp = std::fabs(p_interface[(size_t)(k + 1) * ncol + i] - p_interface[(size_t)k * ncol + i]);
      amount[(size_t)k * ncol + i] = q[(size_t)k * ncol + i] * dp / g;
    }
}

}  // extern "C"
