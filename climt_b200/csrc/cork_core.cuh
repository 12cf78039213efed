// climt_b200 -- CORK correlated-k radiation: per-thread device code (also compiled for the host by tests/emul).
//
// Replaces, fused, the reference's numba kernels
//   _ck_tau_additive_co2_kernel   cork/optics/correlated_k.py:81-117   (and interpolate_k/_continuum :222-330, 420-470)
//   planck_sources_kernel         cork/lw/kernels.py:9-68
//   _lw_transport_kernel          cork/lw/kernels.py:71-121
//   _sw_two_stream_core           cork/sw/kernels.py:18-263  (_delta_scale, _sw_dif_and_source, _adding)
//   compute_column_amount / compute_heating_rate   cork/common.py:38-80
// and the array glue between them in cork/lw/component.py:243-357, cork/sw/component.py:305-447.
//
// Work decomposition: one thread = (column, unit), a unit = U consecutive g-points of one band (U = 1, 2, 4 or 8; all
// units of a table share U).  Lanes of a warp are adjacent columns: every state/workspace/scratch access is a
// contiguous row; the k-table is re-laid-out so that the U g-points of a unit and the two CO2 neighbours of a corner
// are one aligned vector load.  The (tau, B) pairs of pass 1 are parked in per-thread scratch rows for pass 2, the
// per-unit weighted flux sums go to `part` and are reduced over units in a fixed order by cork_reduce_level.
#pragma once
#include <math.h>

#include "cb_common.h"

namespace cb {
namespace cork {

constexpr double kMolarMassDryAir = 28.970;  // cork/common.py:9
constexpr double kMolarMassH2O = 18.015;     // cork/common.py:11

// Picket-fence ("parmentier") optics: the coefficient files of cork/optics/parmentier.py, by value
// (freedman2014.npz, solar_composition.npz; same members as cb200_picket_coeffs).
constexpr int kPicketMaxRegions = 8;
struct Picket {
  int nregion;
  double bounds[kPicketMaxRegions + 1];
  double gv1[kPicketMaxRegions][2], gv2[kPicketMaxRegions][2], gv3[kPicketMaxRegions][2], beta[kPicketMaxRegions][2];
  double quad[3];
  double T_boundary, a_hi, b_hi, c_hi, a_lo, b_lo, c_lo;
};

struct Table {
  int optics;                             // 0 = correlated-k table, 1 = picket fence (no table: `pk` below)
  int ngas, nband, ngpt, nT, nP, nX, nC;  // nX / nC are 1 when the axis is absent
  int hasX, hasC, has_cont, co2_logk;
  int U, nchunk;                          // g-points per unit; units per band
  int nband_pf, ngpt_pf;
  const void* k;                          // [gas][iT][iP][iX][band][chunk][iC][U], float or double as shipped (promoted on use)
  const double* T_grid;
  const double* p_grid_log;
  const double* log_x_grid;
  const double* log_c_grid;
  double x_lo, x_hi, c_lo, c_hi;          // clip range of the VMR axes (correlated_k.py:539-542)
  const double* weights;                  // [band][g]
  const double* planck;                   // [iT][band_pf][g_pf]
  const double* log_cont;                 // [iT][iP][iX][band]  = log(max(continuum_kappa, 1e-40))
  const double* solar;                    // [band][g]
  const double* rayleigh;                 // [band] or null
  Picket pk;                              // optics == 1
};

struct Consts {
  double g, cpd, sigma, D;
};

// inputs shared by LW and SW (device pointers, (nlev, ncol) column-fastest unless noted)
struct In {
  int ncol, nlev;
  const double *T, *p, *p_int, *T_surf;
  const double* q_h2o;    // specific humidity, or null (no H2O axis)
  const double* co2_vmr;  // or null
  const double* gas_q;    // (ngas, nlev, ncol) mass mixing ratios of a non-premixed table, or null -> column mass of air
  // LW
  const double* emissivity;  // (nband, ncol)
  const double* tau_cloud;   // (nlev, ncol, nband) band-fastest, or null
  // SW
  const double *zenith, *albedo;
  const double *ssa_cloud, *g_cloud;  // (nlev, ncol, nband), with tau_cloud
  const double* solar_flux;           // (nband, ngpt): solar_source * earth_sun_factor, prepared by the caller
  // picket fence
  const double *T_irr, *T_int;        // (ncol) irradiation / internal temperature
  const double* bond_albedo;          // (ncol) or null (= 0): SW Bond-albedo feedback, second pass
};

struct Out {
  double *up_broad, *down_broad;  // (nlev+1, ncol)
  double* heating;                // (nlev, ncol) K s-1
  double *up_band, *down_band;    // (nband, nlev+1, ncol) or null
  double *tau_band, *trans_band, *hr_band;  // (nband, nlev, ncol) or null (trans_band: LW only)
  // diagnostics_level >= 1 (cork/lw/component.py:341-358, cork/sw/component.py:455-492): per-band g-point averages / sums of the
  // kernels' per-g-point quantities, band-major like up_band; null = not wanted.  Indexed by DiagLw / DiagSw.
  double* diag[10];
  int diag_ncol, diag_c0;  // their column stride and the first column of the running chunk in them
};
// LW: weighted averages  sum_g w x / sum_g w  of the layer transmittance and of the WEIGHTED per-g-point fluxes (lw/kernels.py:103-116)
enum DiagLw { DL_TRANS = 0, DL_UP, DL_DOWN, DL_N };
// SW: weighted averages of the layer quantities, plain sums over g of the interface quantities (sw/kernels.py:299-394)
enum DiagSw { DS_RDIF = 0, DS_TDIF, DS_TNOSCAT, DS_DIRECT, DS_RDIR, DS_TDIR, DS_TAU_D, DS_SSA_D, DS_G_D, DS_COMB_ALB, DS_N };
CB_HD bool diag_sw_is_interface(int j) { return j == DS_DIRECT || j == DS_COMB_ALB; }

// workspace of one column chunk (ncc columns)
enum WsField { F_FT = 0, F_FP, F_FX, F_FC, F_AMT0 };  // + ngas amounts
struct Work {
  int ncc;
  double* ws;     // [field][lev][ncc]
  int* idx;       // [lev][ncc]  iT | iP<<8 | iX<<16 | iC<<24
  double* scr;    // [unit][row][lev][ncc]
  double* part;   // [unit][3][lev+1][ncc]   0 = up, 1 = down, 2 = tau (lev rows used)
  int nscr;       // scratch rows per unit
  double* dpart;  // [unit][ndiag][lev+1][ncc] per-unit sums of the diagnostics (diagnostics_level >= 1 only)
  int ndiag;      // DL_N | DS_N
  int diag_level; // 0, 1, 2
  const double* wsum;  // (nband) sum of each band's g-point weights as the caller evaluated it (the reference sums the table's own
                       // dtype: a float32 table gives a float32 sum, cork/lw/component.py:342), or null -> summed here in float64
};

CB_HD int pack_idx(int iT, int iP, int iX, int iC) { return iT | (iP << 8) | (iX << 16) | (iC << 24); }

struct Br {
  int i;
  double f;
};
// _ck_bracket (correlated_k.py:30-43): np.searchsorted(grid, v) - 1 clamped to [0, n-2]; fraction clamped to [0, 1]
CB_HD Br bracket(const double* __restrict__ grid, int n, double v) {
  int lo = 0;
  while (lo < n && CB_LDG(grid + lo) < v) ++lo;
  int i = lo - 1;
  if (i < 0) i = 0;
  else if (i > n - 2) i = n - 2;
  const double g0 = CB_LDG(grid + i), g1 = CB_LDG(grid + i + 1);
  double f = (v - g0) / (g1 - g0);
  if (f < 0.0) f = 0.0;
  else if (f > 1.0) f = 1.0;
  return {i, f};
}

// ---- picket-fence optics ----------------------------------------------------------------------------------------
// Freedman et al. (2014) Rosseland mean opacity [m2 kg-1] (compute_rosseland_mean_opacity, cork/optics/parmentier.py:55-74)
CB_HD double picket_kappa_R(const Picket& P, double T, double p) {
  const double log_T = log10(fmax(T, 10.0));
  const double log_P = log10(fmax(p * 10.0, 1.0));
  const double log_k = T < P.T_boundary ? P.a_lo * log_T + P.b_lo * log_P + P.c_lo : P.a_hi * log_T + P.b_hi * log_P + P.c_hi;
  return pow(10.0, log_k) * 0.1;
}

struct PicketCol {
  double gv[3], beta, R;
};
// Per-column ratio coefficients: T_eff of Lee et al. (2021) Eq. 20 with mu* = 1/4 (cork/lw/component.py:386-396,
// cork/sw/component.py:511-514) and lookup_ratio_coefficients (cork/optics/parmentier.py:100-153).  The region search
// keeps the reference's fall-back to region 0 when no interval contains T_eff.
CB_HD PicketCol picket_column(const Picket& P, double T_irr, double T_int, double A_B) {
  double T_eff = pow(pow(T_int, 4.0) + (1.0 - A_B) * 0.25 * pow(T_irr, 4.0), 0.25);
  T_eff = fmax(T_eff, 100.0);
  const double X = log10(fmax(T_eff, 10.0));
  int r = 0;
  for (int i = 0; i < P.nregion; ++i)
    if (T_eff >= P.bounds[i] && T_eff < P.bounds[i + 1]) { r = i; break; }
  PicketCol o;
  o.gv[0] = pow(10.0, P.gv1[r][0] + P.gv1[r][1] * X);
  o.gv[1] = pow(10.0, P.gv2[r][0] + P.gv2[r][1] * X);
  o.gv[2] = pow(10.0, P.gv3[r][0] + P.gv3[r][1] * X);
  o.beta = fmin(fmax(P.beta[r][0] + P.beta[r][1] * X, 0.01), 0.99);
  const double gamma_P = fmax(pow(10.0, P.quad[0] + P.quad[1] * X + P.quad[2] * (X * X)), 1.0);
  const double gm1 = gamma_P - 1.0;
  const double disc = gm1 * gm1 + 4.0 * o.beta * (1.0 - o.beta) * gm1;
  if (disc < 0) {
    o.R = 1.0;
  } else {
    const double den = 2.0 * o.beta * (1.0 - o.beta);
    o.R = fmax(1.0 + gm1 / den + sqrt(disc) / den, 1.0);
  }
  return o;
}

// ---- prep: interpolation coordinates and layer amounts of one (column, level) ------------------------------------
// (_additive_co2_fast, correlated_k.py:526-548; component glue cork/lw/component.py:259-287)
CB_HD void prep_cell(const Table& Tb, const Consts& K, const In& in, const Work& W, int c0, int c, int l) {
  const int ncol = in.ncol, nlev = in.nlev, ncc = W.ncc;
  const size_t o = (size_t)l * ncol + c0 + c;
  const size_t fs = (size_t)nlev * ncc;
  double* ws = W.ws + (size_t)l * ncc + c;
  if (Tb.optics == 1) {  // picket fence: Rosseland mean opacity and layer mass (cork/lw/component.py:402-410)
    ws[F_FT * fs] = picket_kappa_R(Tb.pk, in.T[o], in.p[o]);
    ws[F_AMT0 * fs] = fabs(in.p_int[(size_t)(l + 1) * ncol + c0 + c] - in.p_int[o]) / K.g;
    W.idx[(size_t)l * ncc + c] = 0;
    return;
  }
  const Br bT = bracket(Tb.T_grid, Tb.nT, in.T[o]);
  const Br bP = bracket(Tb.p_grid_log, Tb.nP, log(fmax(in.p[o], 1.0)));
  Br bX{0, 0.0}, bC{0, 0.0};
  if (Tb.hasX) {
    const double q = in.q_h2o[o];
    const double M = kMolarMassH2O / kMolarMassDryAir;
    double x = q / fmax(q + (1.0 - q) * M, 1e-30);
    x = fmin(fmax(x, Tb.x_lo), Tb.x_hi);
    bX = bracket(Tb.log_x_grid, Tb.nX, log(fmax(x, 1e-30)));
  }
  if (Tb.hasC) {
    double x = in.co2_vmr[o];
    x = fmin(fmax(x, Tb.c_lo), Tb.c_hi);
    bC = bracket(Tb.log_c_grid, Tb.nC, log(fmax(x, 1e-30)));
  }
  ws[F_FT * fs] = bT.f; ws[F_FP * fs] = bP.f; ws[F_FX * fs] = bX.f; ws[F_FC * fs] = bC.f;
  W.idx[(size_t)l * ncc + c] = pack_idx(bT.i, bP.i, bX.i, bC.i);
  const double dp = fabs(in.p_int[(size_t)(l + 1) * ncol + c0 + c] - in.p_int[o]);
  for (int ig = 0; ig < Tb.ngas; ++ig) {
    const double q = in.gas_q ? in.gas_q[((size_t)ig * nlev + l) * ncol + c0 + c] : 1.0;
    ws[(F_AMT0 + ig) * fs] = q * dp / K.g;  // compute_column_amount, cork/common.py:62-80
  }
}

// U consecutive table entries -> double (128-bit loads; float32 tables are promoted here)
template <int U>
CB_HD void ldk(const float* __restrict__ p, double* v) {
#if defined(__CUDA_ARCH__)
  if (U == 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else if (U == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else if (U == 2) {
    const float2 a = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = a.x; v[1] = a.y;
  } else {
    v[0] = __ldg(p);
  }
#else
  for (int u = 0; u < U; ++u) v[u] = p[u];
#endif
}
template <int U>
CB_HD void ldk(const double* __restrict__ p, double* v) {
  if (U == 8) {
    const Row<4> a = ldrow<4>(p), b = ldrow<4>(p + 4);
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = a[u]; v[4 + u] = b[u]; }
  } else {
    const Row<(U < 8 ? U : 4)> a = ldrow<(U < 8 ? U : 4)>(p);
#pragma unroll
    for (int u = 0; u < (U < 8 ? U : 4); ++u) v[u] = a[u];
  }
}

// Gas optical depth of U g-points of one band at one (level, column): sum over gases of k * amount, k interpolated
// trilinearly in (T, log p, log X_H2O) at the two bracketing CO2 nodes and geometrically between them
// (_ck_txx7 / _ck_tau_additive_co2_kernel, correlated_k.py:46-60, 97-117), plus the band-grey continuum (:62-78).
template <int U, typename KT>
CB_HD void gas_tau(const Table& Tb, const double* __restrict__ ws, size_t fs, int idx, int band, int chunk, double* tau) {
  const int iT = idx & 255, iP = (idx >> 8) & 255, iX = (idx >> 16) & 255, iC = (idx >> 24) & 255;
  const double fT = ws[F_FT * fs], fP = ws[F_FP * fs], fX = ws[F_FX * fs], fC = ws[F_FC * fs];
  const double aT = 1.0 - fT, aP = 1.0 - fP;
  const size_t sC = (size_t)U, sCh = (size_t)Tb.nC * U, sB = sCh * Tb.nchunk, sX = sB * Tb.nband, sP = sX * Tb.nX,
               sT = sP * Tb.nP, sG = sT * Tb.nT;
#pragma unroll
  for (int u = 0; u < U; ++u) tau[u] = 0.0;
  for (int ig = 0; ig < Tb.ngas; ++ig) {
    const KT* __restrict__ b0 = static_cast<const KT*>(Tb.k) + ig * sG + iT * sT + iP * sP + iX * sX + band * sB + chunk * sCh + iC * sC;
    double kv[U];
    const int ncn = Tb.hasC ? 2 : 1;
    double cc[2][U];
    for (int ic = 0; ic < ncn; ++ic) {
      const KT* __restrict__ b = b0 + ic * sC;
      double v00[U], v10[U], v01[U], v11[U];
      ldk<U>(b, v00); ldk<U>(b + sT, v10); ldk<U>(b + sP, v01); ldk<U>(b + sT + sP, v11);
      double x0[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x0[u] = v00[u] * aT * aP + v10[u] * fT * aP + v01[u] * aT * fP + v11[u] * fT * fP;
      if (Tb.hasX) {
        ldk<U>(b + sX, v00); ldk<U>(b + sX + sT, v10); ldk<U>(b + sX + sP, v01); ldk<U>(b + sX + sT + sP, v11);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double x1 = v00[u] * aT * aP + v10[u] * fT * aP + v01[u] * aT * fP + v11[u] * fT * fP;
          cc[ic][u] = x0[u] * (1.0 - fX) + x1 * fX;
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) cc[ic][u] = x0[u];
      }
    }
    if (Tb.hasC) {
      const double FLOOR = 1e-40;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (Tb.co2_logk) {
          const double l0 = log(cc[0][u] > FLOOR ? cc[0][u] : FLOOR), l1 = log(cc[1][u] > FLOOR ? cc[1][u] : FLOOR);
          kv[u] = exp(l0 * (1.0 - fC) + l1 * fC);
        } else {
          kv[u] = cc[0][u] * (1.0 - fC) + cc[1][u] * fC;
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) kv[u] = cc[0][u];
    }
    const double amt = ws[(F_AMT0 + ig) * fs];
#pragma unroll
    for (int u = 0; u < U; ++u) tau[u] += kv[u] * amt;
  }
  if (Tb.has_cont) {
    const size_t cX = (size_t)Tb.nband, cP = cX * Tb.nX, cT = cP * Tb.nP;
    const double* __restrict__ b = Tb.log_cont + iT * cT + iP * cP + iX * cX + band;
    const double x0 = CB_LDG(b) * aT * aP + CB_LDG(b + cT) * fT * aP + CB_LDG(b + cP) * aT * fP + CB_LDG(b + cT + cP) * fT * fP;
    const double x1 = CB_LDG(b + cX) * aT * aP + CB_LDG(b + cX + cT) * fT * aP + CB_LDG(b + cX + cP) * aT * fP +
                      CB_LDG(b + cX + cT + cP) * fT * fP;
    const double cont = exp(x0 * (1.0 - fX) + x1 * fX);
    const double amt0 = ws[F_AMT0 * fs];
#pragma unroll
    for (int u = 0; u < U; ++u) tau[u] += cont * amt0;
  }
}

// ---- longwave unit -------------------------------------------------------------------------------------------------
// scratch rows per unit: 2U  (trans, planck source per g-point)
// OPT = 1: picket-fence optics (U = 1; band 0 = kappa_1 with Planck share beta, band 1 = kappa_2 with 1 - beta;
// CorkLongwaveRadiation._parmentier_optics, cork/lw/component.py:375-422) in place of the k-table and planck_fraction.
template <int U, typename KT, int OPT = 0, bool DIAG = false>
CB_HD void lw_unit(const Table& Tb, const Consts& K, const In& in, const Work& W, int c0, int c, int band, int chunk, int unit) {
  const int ncol = in.ncol, nlev = in.nlev, ncc = W.ncc;
  const size_t gc = (size_t)c0 + c;
  const size_t fs = (size_t)nlev * ncc;
  const size_t ps = (size_t)(nlev + 1) * ncc;
  const int g0 = chunk * U;
  double w[U];
#pragma unroll
  for (int u = 0; u < U; ++u) w[u] = CB_LDG(Tb.weights + band * Tb.ngpt + g0 + u);
  const int bpf = band < Tb.nband_pf ? band : Tb.nband_pf - 1;
  const double* __restrict__ pf = OPT ? nullptr : Tb.planck + (size_t)bpf * Tb.ngpt_pf + g0;
  const size_t pfT = (size_t)Tb.nband_pf * Tb.ngpt_pf;
  double* __restrict__ part = W.part + (size_t)unit * 3 * ps + c;
  double* __restrict__ scr = W.scr + (size_t)unit * W.nscr * fs + c;
  double* __restrict__ dpart = DIAG ? W.dpart + (size_t)unit * DL_N * ps + c : nullptr;
  double pk_frac = 0.0, pk_c2 = 0.0, pk_R = 1.0;
  if (OPT == 1) {
    const PicketCol pc = picket_column(Tb.pk, in.T_irr[gc], in.T_int[gc], 0.0);
    pk_frac = band == 0 ? pc.beta : 1.0 - pc.beta;
    pk_c2 = pc.beta / pc.R + 1.0 - pc.beta;  // compute_thermal_opacities, cork/optics/parmentier.py:31-33
    pk_R = pc.R;
  }
  // surface source and upward boundary (lw/kernels.py:29-47, 92-95)
  double up[U];
  if (OPT == 1) {
    const double Ts = in.T_surf[gc];
    const double em = in.emissivity[(size_t)band * ncol + gc];
    up[0] = em * (pk_frac * (K.sigma * pow(Ts, 4.0)));  // cork/lw/component.py:418-420
    part[0] = w[0] * up[0];
    if (DIAG) dpart[DL_UP * ps] = w[0] * (w[0] * up[0]);
  } else {
    const double Ts = in.T_surf[gc];
    const Br bs = bracket(Tb.T_grid, Tb.nT, Ts);
    const double planck = K.sigma * ((Ts * Ts) * (Ts * Ts));
    double f0[U], f1[U];
    ldk<U>(pf + bs.i * pfT, f0); ldk<U>(pf + (bs.i + 1) * pfT, f1);
    const double em = in.emissivity[(size_t)band * ncol + gc];
    double s = 0.0, sdu = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double frac = f0[u] * (1.0 - bs.f) + f1[u] * bs.f;
      up[u] = em * (frac * planck);
      s += w[u] * up[u];
      if (DIAG) sdu += w[u] * (w[u] * up[u]);
    }
    part[0] = s;
    if (DIAG) dpart[DL_UP * ps] = sdu;
  }
  // pass 1: surface -> top.  optical depth, Planck source, upward sweep
  for (int l = 0; l < nlev; ++l) {
    const double* __restrict__ ws = W.ws + (size_t)l * ncc + c;
    const int idx = W.idx[(size_t)l * ncc + c];
    double tau[U];
    const double tc = in.tau_cloud ? in.tau_cloud[((size_t)l * ncol + gc) * Tb.nband + band] : 0.0;
    const int iT = idx & 255;
    const double fT = ws[F_FT * fs];
    const double Tl = in.T[(size_t)l * ncol + gc];
    double planck;
    double f0[U], f1[U];
    if (OPT == 1) {
      const double kappa_2 = ws[F_FT * fs] * pk_c2;  // ws[F_FT] holds kappa_R here (prep_cell)
      tau[0] = (band == 0 ? pk_R * kappa_2 : kappa_2) * ws[F_AMT0 * fs];
      planck = K.sigma * pow(Tl, 4.0);
      f0[0] = pk_frac; f1[0] = 0.0;
    } else {
      gas_tau<U, KT>(Tb, ws, fs, idx, band, chunk, tau);
      planck = K.sigma * ((Tl * Tl) * (Tl * Tl));
      ldk<U>(pf + iT * pfT, f0); ldk<U>(pf + (iT + 1) * pfT, f1);
    }
    double su = 0.0, st = 0.0, sdt = 0.0, sdu = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double t = tau[u] + tc;
      const double src = OPT == 1 ? f0[u] * planck : (f0[u] * (1.0 - fT) + f1[u] * fT) * planck;
      const double trans = exp(-K.D * t);
      up[u] = up[u] * trans + src * (1.0 - trans);
      su += w[u] * up[u];
      st += w[u] * t;
      if (DIAG) { sdt += w[u] * trans; sdu += w[u] * (w[u] * up[u]); }
      scr[(size_t)(2 * u) * fs + (size_t)l * ncc] = trans;
      scr[(size_t)(2 * u + 1) * fs + (size_t)l * ncc] = src;
    }
    part[(size_t)(l + 1) * ncc] = su;
    part[2 * ps + (size_t)l * ncc] = st;
    if (DIAG) { dpart[DL_TRANS * ps + (size_t)l * ncc] = sdt; dpart[DL_UP * ps + (size_t)(l + 1) * ncc] = sdu; }
  }
  // pass 2: top -> surface (lw/kernels.py:108-117)
  double dn[U];
#pragma unroll
  for (int u = 0; u < U; ++u) dn[u] = 0.0;
  part[ps + (size_t)nlev * ncc] = 0.0;
  if (DIAG) dpart[DL_DOWN * ps + (size_t)nlev * ncc] = 0.0;
  for (int l = nlev - 1; l >= 0; --l) {
    double sd = 0.0, sdd = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double trans = scr[(size_t)(2 * u) * fs + (size_t)l * ncc];
      const double src = scr[(size_t)(2 * u + 1) * fs + (size_t)l * ncc];
      dn[u] = dn[u] * trans + src * (1.0 - trans);
      sd += w[u] * dn[u];
      if (DIAG) sdd += w[u] * (w[u] * dn[u]);
    }
    part[ps + (size_t)l * ncc] = sd;
    if (DIAG) dpart[DL_DOWN * ps + (size_t)l * ncc] = sdd;
  }
}

// ---- shortwave unit ------------------------------------------------------------------------------------------------
// scratch rows per unit: 7U  (Rdif, Tdif, src_up -> denom, src_dn, direct beam at layer base, albedo, src)
// OPT = 1: picket-fence optics (U = 1; three purely absorbing visible bands, kappa_v = gamma_v * kappa_R;
// CorkShortwaveRadiation._parmentier_sw_optics, cork/sw/component.py:498-532).
template <int U, typename KT, int OPT = 0, bool DIAG = false>
CB_HD void sw_unit(const Table& Tb, const Consts& K, const In& in, const Work& W, int c0, int c, int band, int chunk, int unit) {
  const int ncol = in.ncol, nlev = in.nlev, ncc = W.ncc;
  const size_t gc = (size_t)c0 + c;
  const size_t fs = (size_t)nlev * ncc;
  const size_t ps = (size_t)(nlev + 1) * ncc;
  const int g0 = chunk * U;
  double* __restrict__ part = W.part + (size_t)unit * 3 * ps + c;
  double* __restrict__ scr = W.scr + (size_t)unit * W.nscr * fs + c;
  double* __restrict__ dpart = DIAG ? W.dpart + (size_t)unit * DS_N * ps + c : nullptr;
  const bool diag2 = DIAG && W.diag_level >= 2;
  const double mu0 = cos(in.zenith[gc]);
  const bool night = mu0 <= 1e-4;  // sw/kernels.py:218-219: no contribution
  const double alb_s = in.albedo[gc];
  double w[U], scale[U], sol[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    w[u] = CB_LDG(Tb.weights + band * Tb.ngpt + g0 + u);
    sol[u] = CB_LDG(in.solar_flux + band * Tb.ngpt + g0 + u);
    // scale = solar_flux * mu0 * w (sw/kernels.py:257)
    scale[u] = sol[u] * mu0 * w[u];
  }
  (void)sol;
  const double ray = Tb.rayleigh ? CB_LDG(Tb.rayleigh + band) : 0.0;
  double pk_gv = 0.0;
  if (OPT == 1) pk_gv = picket_column(Tb.pk, in.T_irr[gc], in.T_int[gc], in.bond_albedo ? in.bond_albedo[gc] : 0.0).gv[band];
  const double MIN_K = 1.0e-12, MIN_MU0 = 1.0e-8;
  const double mu0_s = fmax(mu0, MIN_MU0);
  // pass 1: top -> surface.  layer optics, two-stream coefficients, direct beam (sw/kernels.py:233-249)
  double dir[U];
#pragma unroll
  for (int u = 0; u < U; ++u) dir[u] = 1.0;
  if (DIAG) {  // direct beam at the top interface: unit flux x solar_flux x mu0 (sw/kernels.py:355); zero at night (:318-319)
    double sdir = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) sdir += night ? 0.0 : (1.0 * sol[u]) * mu0;
    dpart[DS_DIRECT * ps + (size_t)nlev * ncc] = sdir;
  }
  for (int l = nlev - 1; l >= 0; --l) {
    const double* __restrict__ ws = W.ws + (size_t)l * ncc + c;
    const int idx = W.idx[(size_t)l * ncc + c];
    double tau_abs[U];
    if (OPT == 1) tau_abs[0] = (pk_gv * ws[F_FT * fs]) * ws[F_AMT0 * fs];  // kappa_v * mass, cork/sw/component.py:525-530
    else gas_tau<U, KT>(Tb, ws, fs, idx, band, chunk, tau_abs);
    double tau_ray = 0.0;
    if (Tb.rayleigh) {
      const double dp = fabs(in.p_int[(size_t)(l + 1) * ncol + gc] - in.p_int[(size_t)l * ncol + gc]);
      tau_ray = ray * dp / K.g;  // sw/component.py:359-360
    }
    double tau_c = 0.0, ssa_c = 0.0, g_c = 0.0;
    if (in.tau_cloud) {
      const size_t oc = ((size_t)l * ncol + gc) * Tb.nband + band;
      tau_c = in.tau_cloud[oc]; ssa_c = in.ssa_cloud[oc]; g_c = in.g_cloud[oc];
    }
    double st = 0.0;
    double dsum[DS_N];
    if (DIAG) for (int j = 0; j < DS_N; ++j) dsum[j] = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // gas + Rayleigh (sw/component.py:356-368), then gas + cloud mixing (:388-412)
      double tau = tau_abs[u], ssa = 0.0;
      if (Tb.rayleigh) {
        const double tot = tau_abs[u] + tau_ray;
        ssa = tot > 0 ? fdiv(tau_ray, tot) : 0.0;
        tau = tot;
      }
      const double tau_total = tau + tau_c;
      const double scat_gas = tau * ssa, scat_cloud = tau_c * ssa_c;
      const double scat_total = scat_gas + scat_cloud;
      const double ssa_t = tau_total > 0 ? fdiv(scat_total, tau_total) : 0.0;
      const double g_t = scat_total > 0 ? fdiv(scat_gas * 0.0 + scat_cloud * g_c, scat_total) : 0.0;
      st += w[u] * tau_total;
      if (night) continue;
      // _delta_scale (sw/kernels.py:18-35)
      const double f = g_t * g_t;
      const double tau_s = tau_total * (1.0 - ssa_t * f);
      const double w0 = (1.0 - ssa_t * f) > 1e-30 ? fdiv(ssa_t * (1.0 - f), 1.0 - ssa_t * f) : 0.0;
      const double gs = (1.0 - f) > 1e-30 ? fdiv(g_t - f, 1.0 - f) : 0.0;
      // _sw_dif_and_source (sw/kernels.py:38-118)
      const double gamma1 = (8.0 - w0 * (5.0 + 3.0 * gs)) * 0.25;
      const double gamma2 = 3.0 * (w0 * (1.0 - gs)) * 0.25;
      const double kk = sqrt(fmax((gamma1 - gamma2) * (gamma1 + gamma2), MIN_K));
      const double e1 = exp(-tau_s * kk);
      const double e2 = e1 * e1;
      const double RT = frcp(kk * (1.0 + e2) + gamma1 * (1.0 - e2));
      const double Rdif = RT * gamma2 * (1.0 - e2);
      const double Tdif = RT * 2.0 * kk * e1;
      const double Tnoscat = exp(-fdiv(tau_s, mu0_s));
      const double k_mu = kk * mu0_s;
      double denom_dir = 1.0 - k_mu * k_mu;
      if (fabs(denom_dir) < 1e-30) denom_dir = 1e-30;
      const double RTd = fdiv(w0 * RT, denom_dir);
      const double gamma3 = (2.0 - 3.0 * mu0_s * gs) * 0.25;
      const double gamma4 = 1.0 - gamma3;
      const double alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
      const double alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
      const double k_g3 = kk * gamma3, k_g4 = kk * gamma4;
      double Rdir = RTd * ((1.0 - k_mu) * (alpha2 + k_g3) - (1.0 + k_mu) * (alpha2 - k_g3) * e2 -
                           2.0 * (k_g3 - alpha2 * k_mu) * e1 * Tnoscat);
      double Tdir = -RTd * ((1.0 + k_mu) * (alpha1 + k_g4) * Tnoscat - (1.0 - k_mu) * (alpha1 - k_g4) * e2 * Tnoscat -
                            2.0 * (k_g4 + alpha1 * k_mu) * e1);
      Rdir = fmax(0.0, fmin(Rdir, 1.0 - Tnoscat));
      Tdir = fmax(0.0, fmin(Tdir, 1.0 - Tnoscat - Rdir));
      const double above = dir[u];
      dir[u] = Tnoscat * above;
      double* __restrict__ s = scr + (size_t)(7 * u) * fs + (size_t)l * ncc;
      s[0 * fs] = Rdif; s[1 * fs] = Tdif; s[2 * fs] = Rdir * above; s[3 * fs] = Tdir * above; s[4 * fs] = dir[u];
      if (DIAG) {
        dsum[DS_RDIF] += w[u] * Rdif; dsum[DS_TDIF] += w[u] * Tdif; dsum[DS_TNOSCAT] += w[u] * Tnoscat;
        dsum[DS_DIRECT] += (dir[u] * sol[u]) * mu0;
        if (diag2) {
          dsum[DS_RDIR] += w[u] * Rdir; dsum[DS_TDIR] += w[u] * Tdir;
          dsum[DS_TAU_D] += w[u] * tau_s; dsum[DS_SSA_D] += w[u] * w0; dsum[DS_G_D] += w[u] * gs;
        }
      }
    }
    part[2 * ps + (size_t)l * ncc] = st;
    if (DIAG) {  // (night: every sum is still zero, as in the reference, whose column loop `continue`s)
      for (int j = 0; j < DS_COMB_ALB; ++j) dpart[(size_t)j * ps + (size_t)l * ncc] = dsum[j];
    }
  }
  if (night) {
    for (int l = 0; l <= nlev; ++l) {
      part[(size_t)l * ncc] = 0.0; part[ps + (size_t)l * ncc] = 0.0;
      if (DIAG) dpart[DS_COMB_ALB * ps + (size_t)l * ncc] = 0.0;
    }
    return;
  }
  // pass 2: surface -> top.  combined albedo and source below each interface (_adding, sw/kernels.py:150-164)
  double alb[U], src[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { alb[u] = alb_s; src[u] = dir[u] * alb_s; }
  for (int l = 0; l < nlev; ++l) {
    double salb = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double* __restrict__ s = scr + (size_t)(7 * u) * fs + (size_t)l * ncc;
      const double Rdif = s[0 * fs], Tdif = s[1 * fs], src_up = s[2 * fs], src_dn = s[3 * fs];
      const double denom = frcp(1.0 - Rdif * alb[u]);
      s[2 * fs] = denom; s[5 * fs] = alb[u]; s[6 * fs] = src[u];
      if (DIAG) salb += alb[u];  // combined albedo below interface l (sw/kernels.py:360-367: the same recurrence)
      const double a1 = Rdif + Tdif * Tdif * alb[u] * denom;
      const double s1 = src_up + Tdif * denom * (src[u] + alb[u] * src_dn);
      alb[u] = a1; src[u] = s1;
    }
    if (DIAG) dpart[DS_COMB_ALB * ps + (size_t)l * ncc] = diag2 ? salb : 0.0;
  }
  if (DIAG) {
    double salb = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) salb += alb[u];
    dpart[DS_COMB_ALB * ps + (size_t)nlev * ncc] = diag2 ? salb : 0.0;
  }
  // pass 3: top -> surface.  diffuse fluxes (sw/kernels.py:170-186) and the weighted band sums (:257-260)
  double fdn[U];
  {
    double su = 0.0, sd = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      fdn[u] = 0.0;
      const double fup = fdn[u] * alb[u] + src[u];
      su += fup * scale[u];
      sd += (1.0 + fdn[u]) * scale[u];
    }
    part[(size_t)nlev * ncc] = su;
    part[ps + (size_t)nlev * ncc] = sd;
  }
  for (int l = nlev - 1; l >= 0; --l) {
    double su = 0.0, sd = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* __restrict__ s = scr + (size_t)(7 * u) * fs + (size_t)l * ncc;
      const double Rdif = s[0 * fs], Tdif = s[1 * fs], denom = s[2 * fs], src_dn = s[3 * fs], dirk = s[4 * fs], albk = s[5 * fs],
                   srck = s[6 * fs];
      fdn[u] = (Tdif * fdn[u] + Rdif * srck + src_dn) * denom;
      const double fup = fdn[u] * albk + srck;
      su += fup * scale[u];
      sd += (dirk + fdn[u]) * scale[u];
    }
    part[(size_t)l * ncc] = su;
    part[ps + (size_t)l * ncc] = sd;
  }
}

// ---- reduction over units: per-band and broadband fluxes at one interface (fixed order: g ascending, then band) ---
CB_HD void reduce_level(const Table& Tb, const Work& W, int nlev, int ncol, int c0, int c, int lev, const Out& out) {
  const int ncc = W.ncc;
  const size_t ps = (size_t)(nlev + 1) * ncc;
  const size_t gc = (size_t)c0 + c;
  double ub = 0.0, db = 0.0;
  for (int b = 0; b < Tb.nband; ++b) {
    double u = 0.0, d = 0.0, t = 0.0;
    for (int ch = 0; ch < Tb.nchunk; ++ch) {
      const double* p = W.part + (size_t)(b * Tb.nchunk + ch) * 3 * ps + (size_t)lev * ncc + c;
      u += p[0];
      d += p[ps];
      if (lev < nlev) t += p[2 * ps];
    }
    if (out.up_band) out.up_band[((size_t)b * (nlev + 1) + lev) * ncol + gc] = u;
    if (out.down_band) out.down_band[((size_t)b * (nlev + 1) + lev) * ncol + gc] = d;
    if (out.tau_band && lev < nlev) out.tau_band[((size_t)b * nlev + lev) * ncol + gc] = t;
    ub += u;
    db += d;
  }
  out.up_broad[(size_t)lev * ncol + gc] = ub;
  out.down_broad[(size_t)lev * ncol + gc] = db;
}

// diagnostics_level >= 1: the per-band g-point averages (weighted, / sum of the band's weights) or sums of the per-unit sums
// (cork/lw/component.py:341-358; cork/sw/component.py:463-492)
CB_HD void reduce_diag_level(const Table& Tb, const Work& W, int nlev, int c, int lev, bool lw, const Out& out) {
  const int ncc = W.ncc;
  const size_t ps = (size_t)(nlev + 1) * ncc;
  const size_t gc = (size_t)out.diag_c0 + c;
  for (int b = 0; b < Tb.nband; ++b) {
    double wsum = 0.0;
    if (W.wsum) wsum = W.wsum[b];
    else for (int g = 0; g < Tb.ngpt; ++g) wsum += Tb.weights[b * Tb.ngpt + g];
    for (int j = 0; j < W.ndiag; ++j) {
      if (!out.diag[j]) continue;
      const bool iface = lw ? (j != DL_TRANS) : diag_sw_is_interface(j);
      if (!iface && lev == nlev) continue;
      double v = 0.0;
      for (int ch = 0; ch < Tb.nchunk; ++ch)
        v += W.dpart[((size_t)(b * Tb.nchunk + ch) * W.ndiag + j) * ps + (size_t)lev * ncc + c];
      if (lw || !iface) v = v / wsum;
      out.diag[j][((size_t)b * (iface ? nlev + 1 : nlev) + lev) * out.diag_ncol + gc] = v;
    }
  }
}

// heating rates of one layer: broadband (K s-1) and per band (K day-1), transmittance per band
// (compute_heating_rate cork/common.py:38-59; cork/lw/component.py:320-333)
CB_HD void heat_layer(const Table& Tb, const Consts& K, const In& in, const Out& out, int c0, int c, int l, bool lw) {
  const int ncol = in.ncol, nlev = in.nlev;
  const size_t gc = (size_t)c0 + c;
  const size_t o0 = (size_t)l * ncol + gc, o1 = (size_t)(l + 1) * ncol + gc;
  const double dp = in.p_int[o1] - in.p_int[o0];
  const double net1 = out.up_broad[o1] - out.down_broad[o1], net0 = out.up_broad[o0] - out.down_broad[o0];
  out.heating[o0] = K.g / K.cpd * (net1 - net0) / dp;
  for (int b = 0; b < Tb.nband; ++b) {
    if (out.hr_band && out.up_band && out.down_band) {
      const size_t b0 = ((size_t)b * (nlev + 1) + l) * ncol + gc, b1 = ((size_t)b * (nlev + 1) + l + 1) * ncol + gc;
      const double n1 = out.up_band[b1] - out.down_band[b1], n0 = out.up_band[b0] - out.down_band[b0];
      out.hr_band[((size_t)b * nlev + l) * ncol + gc] = (K.g / K.cpd * (n1 - n0) / dp) * 86400.0;
    }
    if (lw && out.trans_band && out.tau_band) {
      const size_t ob = ((size_t)b * nlev + l) * ncol + gc;
      out.trans_band[ob] = exp(-K.D * out.tau_band[ob]);
    }
  }
}

}  // namespace cork
}  // namespace cb
