// climt_b200 -- CORK k-table re-layout (host only; shared by the CUDA engine and the test-only host emulation).
// Reference order k(gas, band, g, T, P[, X[, C]]) (correlated_k.py:9-21) -> [gas][T][P][X][band][chunk][C][U]: the U g-points
// of a unit and both CO2 neighbours of an interpolation corner are contiguous; planck_fraction(band, g, T) -> [T][band][g];
// continuum_kappa(band, T, P, X) -> log() as [T][P][X][band].
#pragma once
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/climt_b200.h"
#include "cork_core.cuh"

namespace cb {
namespace cork {

struct TableImage {
  std::vector<float> k32;   // exactly one of k32 / k64 is filled (the table's own dtype)
  std::vector<double> k64, planck;
  std::vector<double> d;
  bool k_f64 = false;
  size_t oT = 0, oP = 0, oX = 0, oC = 0, oW = 0, oCont = 0, oSol = 0, oRay = 0;
  bool is_lw = false, is_sw = false, has_ray = false;
};

inline std::string check_table(const cb200_cork_table* t) {
  if (!t || ((t->k_coefficients_f32 != nullptr) == (t->k_coefficients_f64 != nullptr)) || !t->temperature_grid || !t->pressure_grid_log || !t->gpoint_weights) return "cork: incomplete table";
  if (t->ngas < 1 || t->nband < 1 || t->ngpt < 1 || t->nT < 2 || t->nP < 2 || t->nT > 255 || t->nP > 255 || t->nX > 255 || t->nC > 255)
    return "cork: unsupported table dimensions";
  if (t->nC > 0 && t->nX == 0) return "cork: a CO2 axis requires an H2O axis (correlated_k.py:222-330)";
  if ((t->nX > 0 && !t->h2o_vmr_grid) || (t->nC > 0 && !t->co2_vmr_grid)) return "cork: missing VMR grid";
  if (t->planck_fraction && t->ngpt_pf != t->ngpt)
    return "cork: planck_fraction g-point count differs from the table's (ESFT tables are not supported)";
  return "";
}

// pointers of T are left null; bind() sets them from wherever the images end up (device or host)
inline void build_images(const cb200_cork_table* t, int umax, Table& T, TableImage& im) {
  im.is_lw = t->planck_fraction != nullptr;
  im.is_sw = t->solar_source_per_gpoint != nullptr;
  im.has_ray = t->rayleigh_coefficient != nullptr;
  T = Table{};
  T.ngas = t->ngas; T.nband = t->nband; T.ngpt = t->ngpt; T.nT = t->nT; T.nP = t->nP;
  T.hasX = t->nX > 0; T.hasC = t->nC > 0;
  T.nX = T.hasX ? t->nX : 1;
  T.nC = T.hasC ? t->nC : 1;
  T.has_cont = (t->continuum_kappa != nullptr && T.hasX) ? 1 : 0;
  T.co2_logk = t->co2_logk;
  int U = 1;
  for (int cand : {8, 4, 2})
    if (cand <= umax && t->ngpt % cand == 0) { U = cand; break; }
  T.U = U;
  T.nchunk = t->ngpt / U;
  T.nband_pf = t->nband_pf; T.ngpt_pf = t->ngpt_pf;
  const size_t nk = (size_t)T.ngas * T.nband * T.ngpt * T.nT * T.nP * T.nX * T.nC;
  im.k_f64 = t->k_coefficients_f64 != nullptr;
  if (im.k_f64) im.k64.resize(nk); else im.k32.resize(nk);
  size_t src = 0;
  for (int ig = 0; ig < T.ngas; ++ig)
    for (int b = 0; b < T.nband; ++b)
      for (int gp = 0; gp < T.ngpt; ++gp)
        for (int iT = 0; iT < T.nT; ++iT)
          for (int iP = 0; iP < T.nP; ++iP)
            for (int iX = 0; iX < T.nX; ++iX)
              for (int iC = 0; iC < T.nC; ++iC, ++src) {
                const int ch = gp / U, u = gp % U;
                const size_t dst = ((((((((size_t)ig * T.nT + iT) * T.nP + iP) * T.nX + iX) * T.nband + b) * T.nchunk + ch) * T.nC + iC) * U) + u;
                if (im.k_f64) im.k64[dst] = t->k_coefficients_f64[src]; else im.k32[dst] = t->k_coefficients_f32[src];
              }
  if (im.is_lw) {
    im.planck.resize((size_t)T.nT * T.nband_pf * T.ngpt_pf);
    for (int b = 0; b < T.nband_pf; ++b)
      for (int gp = 0; gp < T.ngpt_pf; ++gp)
        for (int iT = 0; iT < T.nT; ++iT)
          im.planck[((size_t)iT * T.nband_pf + b) * T.ngpt_pf + gp] = t->planck_fraction[((size_t)b * T.ngpt_pf + gp) * T.nT + iT];
  }
  auto put = [&](const double* p, size_t n) { size_t off = im.d.size(); im.d.insert(im.d.end(), p, p + n); return off; };
  im.oT = put(t->temperature_grid, T.nT);
  im.oP = put(t->pressure_grid_log, T.nP);
  if (T.hasX) {
    std::vector<double> lx(T.nX);
    for (int i = 0; i < T.nX; ++i) lx[i] = std::log(std::fmax(t->h2o_vmr_grid[i], 1e-30));  // correlated_k.py:534
    im.oX = put(lx.data(), lx.size());
    T.x_lo = t->h2o_vmr_grid[0]; T.x_hi = t->h2o_vmr_grid[T.nX - 1];
  }
  if (T.hasC) {
    std::vector<double> lc(T.nC);
    for (int i = 0; i < T.nC; ++i) lc[i] = std::log(std::fmax(t->co2_vmr_grid[i], 1e-30));  // :535
    im.oC = put(lc.data(), lc.size());
    T.c_lo = t->co2_vmr_grid[0]; T.c_hi = t->co2_vmr_grid[T.nC - 1];
  }
  im.oW = put(t->gpoint_weights, (size_t)T.nband * T.ngpt);
  if (T.has_cont) {
    std::vector<double> lc((size_t)T.nband * T.nT * T.nP * T.nX);
    for (int b = 0; b < T.nband; ++b)
      for (int iT = 0; iT < T.nT; ++iT)
        for (int iP = 0; iP < T.nP; ++iP)
          for (int iX = 0; iX < T.nX; ++iX)
            lc[(((size_t)iT * T.nP + iP) * T.nX + iX) * T.nband + b] =
                std::log(std::fmax(t->continuum_kappa[(((size_t)b * T.nT + iT) * T.nP + iP) * T.nX + iX], 1e-40));  // :550-552
    im.oCont = put(lc.data(), lc.size());
  }
  if (im.is_sw) im.oSol = put(t->solar_source_per_gpoint, (size_t)T.nband * T.ngpt);
  if (im.has_ray) im.oRay = put(t->rayleigh_coefficient, T.nband);
}

inline void bind(Table& T, const TableImage& im, const void* k, const double* planck, const double* dd) {
  T.k = k;
  T.planck = im.is_lw ? planck : nullptr;
  T.T_grid = dd + im.oT; T.p_grid_log = dd + im.oP;
  T.log_x_grid = T.hasX ? dd + im.oX : nullptr;
  T.log_c_grid = T.hasC ? dd + im.oC : nullptr;
  T.weights = dd + im.oW;
  T.log_cont = T.has_cont ? dd + im.oCont : nullptr;
  T.solar = im.is_sw ? dd + im.oSol : nullptr;
  T.rayleigh = im.has_ray ? dd + im.oRay : nullptr;
}

}  // namespace cork
}  // namespace cb
