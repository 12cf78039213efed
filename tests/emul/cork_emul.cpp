// TEST-ONLY host emulation of the CUDA CORK engine (same per-thread code as the kernels, stepped serially).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../climt_b200/csrc/cork_tables.h"

using namespace cb::cork;

template <int U, typename KT, int OPT = 0>
static void run_units(bool lw, const Table& T, const Consts& K, const In& in, const Work& W, int n) {
  for (int unit = 0; unit < T.nband * T.nchunk; ++unit) {
    const int band = unit / T.nchunk, chunk = unit - band * T.nchunk;
    for (int c = 0; c < n; ++c) {
      if (W.diag_level > 0) {
        if (lw) lw_unit<U, KT, OPT, true>(T, K, in, W, 0, c, band, chunk, unit);
        else sw_unit<U, KT, OPT, true>(T, K, in, W, 0, c, band, chunk, unit);
      } else {
        if (lw) lw_unit<U, KT, OPT>(T, K, in, W, 0, c, band, chunk, unit);
        else sw_unit<U, KT, OPT>(T, K, in, W, 0, c, band, chunk, unit);
      }
    }
  }
}

static int run(const Table& T, bool k_f64, bool premixed, int lw, const double* scal, const double* solar_flux, int ncol, int nlev,
               const double* const* inp, double* const* outp, int diag_level = 0, double* const* diagp = nullptr,
               const double* wsum = nullptr);

// the same with diagnostics_level >= 1: diagp = 10 pointers in cb200_cork_diagnostics.field order (null = not wanted)
extern "C" int emul_cork_run_diag(const cb200_cork_table* t, int lw, int umax, const double* scal, const double* solar_flux, int ncol,
                                  int nlev, const double* const* inp, double* const* outp, int diag_level, double* const* diagp,
                                  const double* wsum) {
  if (!check_table(t).empty()) return -1;
  Table T;
  TableImage im;
  build_images(t, umax, T, im);
  bind(T, im, im.k_f64 ? static_cast<const void*>(im.k64.data()) : static_cast<const void*>(im.k32.data()), im.planck.data(), im.d.data());
  return run(T, im.k_f64, t->premixed != 0, lw, scal, solar_flux, ncol, nlev, inp, outp, diag_level, diagp, wsum);
}

// scal = {g, cpd, sigma, D}; solar_flux (nband, ngpt) for sw; inp: 16 pointers in cb200_cork_inputs order; outp: 8 pointers
extern "C" int emul_cork_run(const cb200_cork_table* t, int lw, int umax, const double* scal, const double* solar_flux, int ncol, int nlev,
                             const double* const* inp, double* const* outp) {
  if (!check_table(t).empty()) return -1;
  Table T;
  TableImage im;
  build_images(t, umax, T, im);
  bind(T, im, im.k_f64 ? static_cast<const void*>(im.k64.data()) : static_cast<const void*>(im.k32.data()), im.planck.data(), im.d.data());
  return run(T, im.k_f64, t->premixed != 0, lw, scal, solar_flux, ncol, nlev, inp, outp);
}

// picket-fence engine (cb200_cork_create_picket): same driver, Table set up as the engine does
extern "C" int emul_picket_run(const cb200_picket_coeffs* c, int lw, const double* scal, const double* solar_flux, int ncol, int nlev,
                               const double* const* inp, double* const* outp) {
  static const double ones[3] = {1.0, 1.0, 1.0};
  Table T{};
  T.optics = 1;
  T.ngas = 1; T.nband = lw ? 2 : 3; T.ngpt = 1; T.U = 1; T.nchunk = 1;
  T.nT = T.nP = T.nX = T.nC = 1;
  std::memcpy(&T.pk, c, sizeof(Picket));
  T.weights = ones;
  return run(T, false, true, lw, scal, solar_flux, ncol, nlev, inp, outp);
}

static int run(const Table& T, bool k_f64, bool premixed, int lw, const double* scal, const double* solar_flux, int ncol, int nlev,
               const double* const* inp, double* const* outp, int diag_level, double* const* diagp, const double* wsum) {
  Consts K{scal[0], scal[1], scal[2], scal[3]};
  In in{};
  in.ncol = ncol; in.nlev = nlev;
  in.T = inp[0]; in.p = inp[1]; in.p_int = inp[2]; in.T_surf = inp[3]; in.q_h2o = inp[4]; in.co2_vmr = inp[5];
  in.gas_q = premixed ? nullptr : inp[6];
  in.emissivity = inp[7]; in.tau_cloud = inp[8]; in.zenith = inp[9]; in.albedo = inp[10]; in.ssa_cloud = inp[11]; in.g_cloud = inp[12];
  in.solar_flux = solar_flux;
  in.T_irr = inp[13]; in.T_int = inp[14]; in.bond_albedo = inp[15];
  Out out{};
  out.up_broad = outp[0]; out.down_broad = outp[1]; out.heating = outp[2]; out.up_band = outp[3]; out.down_band = outp[4];
  out.tau_band = outp[5]; out.trans_band = outp[6]; out.hr_band = outp[7];
  const int nunits = T.nband * T.nchunk;
  Work W{};
  W.ncc = ncol;
  W.nscr = lw ? 2 * T.U : 7 * T.U;
  std::vector<double> ws((size_t)(F_AMT0 + T.ngas) * nlev * ncol), scr((size_t)nunits * W.nscr * nlev * ncol),
      part((size_t)nunits * 3 * (nlev + 1) * ncol);
  std::vector<int> idx((size_t)nlev * ncol);
  W.ws = ws.data(); W.idx = idx.data(); W.scr = scr.data(); W.part = part.data();
  std::vector<double> dpart;
  if (diag_level > 0) {
    W.ndiag = lw ? (int)DL_N : (int)DS_N;
    W.diag_level = diag_level;
    W.wsum = wsum;
    dpart.assign((size_t)nunits * W.ndiag * (nlev + 1) * ncol, std::nan(""));
    W.dpart = dpart.data();
    for (int j = 0; j < W.ndiag; ++j) out.diag[j] = diagp[j];
    out.diag_ncol = ncol; out.diag_c0 = 0;
  }
  for (int l = 0; l < nlev; ++l)
    for (int c = 0; c < ncol; ++c) prep_cell(T, K, in, W, 0, c, l);
#define CB_RUN(U)                                                                 \
  case U:                                                                         \
    if (k_f64) run_units<U, double>(lw != 0, T, K, in, W, ncol);                  \
    else run_units<U, float>(lw != 0, T, K, in, W, ncol);                         \
    break;
  if (T.optics == 1) run_units<1, float, 1>(lw != 0, T, K, in, W, ncol);
  else switch (T.U) { CB_RUN(1) CB_RUN(2) CB_RUN(4) CB_RUN(8) }
#undef CB_RUN
  for (int lev = 0; lev <= nlev; ++lev)
    for (int c = 0; c < ncol; ++c) {
      reduce_level(T, W, nlev, ncol, 0, c, lev, out);
      if (diag_level > 0) reduce_diag_level(T, W, nlev, c, lev, lw != 0, out);
    }
  for (int l = 0; l < nlev; ++l)
    for (int c = 0; c < ncol; ++c) heat_layer(T, K, in, out, 0, c, l, lw != 0);
  return 0;
}
