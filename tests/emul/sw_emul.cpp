// TEST-ONLY host emulation of the CUDA SW engine (same per-thread code as the kernels, stepped serially).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../climt_b200/csrc/sw_tables.h"
#include "../../climt_b200/csrc/mcica_host.h"

using namespace cb::sw;

template <int B, int U>
static void run_taumol(const Tables& T, const Solar& sol, const In& in, const Work& W, int n, int g0) {
  // two layer chunks, like two blockIdx.z slices of the kernel
  const int half = in.nlay / 2;
  for (int c = 0; c < n; ++c) {
    sw_taumol_unit<B, U>(T, sol, in, W, 0, c, g0, half, in.nlay);
    sw_taumol_unit<B, U>(T, sol, in, W, 0, c, g0, 0, half);
  }
}

template <int U>
static void run_transfer(const Tables& T, const Solar& sol, const In& in, const Flags& fl, const Work& W, int n, int ib, int g0, int unit, bool mc) {
  for (int c = 0; c < n; ++c) {
    // the units of a group accumulate into the group's (zeroed) rows in unit order, like the warps of a block on the device
    const size_t pstride = (size_t)(in.nlay + 1) * W.ncc;
    SwPartDirect sink{W.part + (size_t)(unit / CB_SW_GROUP) * 4 * pstride + c, pstride, W.ncc};
    if (mc) sw_transfer_unit<U, true>(T, sol, in, fl, W, 0, c, ib, g0, sink);
    else sw_transfer_unit<U, false>(T, sol, in, fl, W, 0, c, ib, g0, sink);
  }
}

// The column-tile form of the transfer (sw_core.cuh: sw_tile_cell / sw_tile_sweeps; CUDA kernel k_sw_tile): the cells of one g-point
// are evaluated into NaN-poisoned row buffers, then the column's sweeps run over them -- same code, serial schedule.
static int g_tile = 0;
extern "C" void emul_sw_set_tile(int on) { g_tile = on; }

// The scan form of the sweeps (sw_core.cuh: scan_up_local / scan_dn_local / scan_walk; CUDA kernel k_sw_scan) for one stream of
// one column: the kScanLanes lanes of the column's group stepped one after the other, the shuffle scans as array passes in the
// same order (Hillis-Steele), per-lane accumulators scattered to the per-interface sums.
static void scan_column(const double* Ps, int nlay, double albdir, double albdif, double zinc, double* acc_up, double* acc_dn) {
  constexpr int W8 = kScanLanes, KLMAX = 32;
  const int KL = (nlay + 1 + W8 - 1) / W8;
  UpMap up[W8], upx[W8];
  DnMap dn[W8], dnx[W8];
  for (int i = 0; i < W8; ++i) {
    auto row = [&](int r, int k) { return Ps[(size_t)r * nlay + i * KL + k]; };
    scan_local<KLMAX>(row, nlay, i, KL, up[i], dn[i]);
  }
  for (int d = 1; d < W8; d <<= 1) {
    UpMap u2[W8];
    DnMap d2[W8];
    for (int i = 0; i < W8; ++i) {
      u2[i] = i >= d ? up_compose(up[i], up[i - d]) : up[i];
      d2[i] = i + d < W8 ? dn_compose(dn[i], dn[i + d]) : dn[i];
    }
    for (int i = 0; i < W8; ++i) { up[i] = u2[i]; dn[i] = d2[i]; }
  }
  for (int i = 0; i < W8; ++i) { upx[i] = i ? up[i - 1] : up_identity(); dnx[i] = i + 1 < W8 ? dn[i + 1] : dn_identity(); }
  for (int i = 0; i < W8; ++i) {
    double au[KLMAX] = {0.}, ad[KLMAX] = {0.};
    auto row = [&](int r, int k) { return Ps[(size_t)r * nlay + i * KL + k]; };
    scan_walk<KLMAX>(row, nlay, i, KL, upx[i], dnx[i], albdir, albdif, zinc, au, ad);
    for (int k = 0; k < KL; ++k)
      if (i * KL + k <= nlay) { acc_up[i * KL + k] += au[k]; acc_dn[i * KL + k] += ad[k]; }
  }
}

template <bool MC, bool CLOUDY>
static void run_tile(const Tables& T, const Solar& sol, const In& in, const Flags& fl, const Work& W, int n) {
  const int nlay = in.nlay;
  constexpr int NR = CLOUDY ? kSwTileRowsCloudy : kSwTileRowsClear;
  const size_t pstride = (size_t)(nlay + 1) * W.ncc, as = (size_t)(nlay + 1);
  std::vector<double> P((size_t)NR * nlay), R((size_t)2 * nlay), acc((size_t)4 * (nlay + 1));
  for (int group = 0; group < kTileGroups; ++group) {
    int ib0, ib1;
    sw_tile_group_bands(group, ib0, ib1);
    for (int c = 0; c < n; ++c) {
      std::fill(acc.begin(), acc.end(), 0.0);
      const bool cloudy_col = W.anycld[c] != 0;
      double prmu0 = in.coszen[c];
      if (prmu0 < 1.e-10) prmu0 = 1.e-10;
      for (int ib = ib0; ib < ib1; ++ib) {
        const bool nir = (ib <= 8) || ib == 13;
        const double albdir = nir ? in.aldir[c] : in.asdir[c], albdif = nir ? in.aldif[c] : in.asdif[c];
        for (int g = 0; g < band_ngpt(ib); ++g) {
          const int gabs = band_gstart(ib) + g;
          std::fill(P.begin(), P.end(), std::nan(""));
          std::fill(R.begin(), R.end(), std::nan(""));
          for (int l = 0; l < nlay; ++l) {
            double taug, taur;
            sw_tile_cell_load(in, W, c, l, gabs, taug, taur);
            sw_tile_cell<MC, CLOUDY>(T, in, fl, W, 0, c, l, ib, gabs, prmu0, cloudy_col, taug, taur, P.data() + l, (size_t)nlay);
          }
          const double zinc = sol.adjflux[ib] * W.src[(size_t)gabs * W.ncc + c] * prmu0;
          // clear-sky stream -> rows 2, 3; total-sky stream (cloudy form) -> rows 0, 1
          for (int stream = 0; stream < (CLOUDY ? 2 : 1); ++stream) {
            const double* Ps = P.data() + (size_t)stream * 5 * nlay;
            double* up = acc.data() + (size_t)(stream ? 0 : 2) * as;
            double* dn = up + as;
            if (g_tile == 2) scan_column(Ps, nlay, albdir, albdif, zinc, up, dn);
            else sw_tile_sweeps(Ps, (size_t)nlay, 1, nlay, albdir, albdif, zinc, R.data(), (size_t)nlay, 1, up, dn, 1);
          }
        }
      }
      for (int q = CLOUDY ? 0 : 2; q < 4; ++q)
        for (int lev = 0; lev <= nlay; ++lev)
          W.part[((size_t)group * 4 + q) * pstride + (size_t)lev * W.ncc + c] = acc[(size_t)q * as + lev];
    }
  }
}

// development hook (tools/adding_scan_study.py): when set, the per-g-point scratch rows [112][NSCR][nlay][ncol] of the last run are
// copied here after the transfer pass
static double* g_scr_export = nullptr;
extern "C" void emul_sw_export_scratch(double* dst) { g_scr_export = dst; }
extern "C" int emul_sw_nscr() { return NSCR; }

// scal = {adjes, scon, solcycfrac, indsolvar0, indsolvar1, bndsolvar[14]}; iopt = {icld, iaer, inflag, iceflag, liqflag, isolvar, dyofyr, mcica, irng, permuteseed}
extern "C" int emul_sw_run(const char* blob, const double* consts11, const int* iopt, const double* scal, int ncol, int nlay,
                           const double* const* inp /*29 pointers in struct In order*/, double* const* outp /*6*/) {
  try {
    Constants k;
    std::memcpy(&k, consts11, sizeof(k));
    std::vector<double> img;
    Tables T;
    build_tables(blob, k, img, T);
    T.base = img.data();
    In in;
    in.ncol = ncol; in.nlay = nlay;
    const double** ip = &in.play;
    for (int i = 0; i < 29; ++i) ip[i] = inp[i];
    Out out;
    double** op = &out.uflx;
    for (int i = 0; i < 6; ++i) op[i] = outp[i];
    Flags fl{iopt[0], iopt[1], iopt[2], iopt[3], iopt[4], iopt[7]};
    const int irng = iopt[8], seed = iopt[9];
    const bool mc = fl.mcica && fl.icld >= 1;
    SolarOptions so;
    so.isolvar = iopt[5]; so.scon = scal[1]; so.indsolvar[0] = scal[3]; so.indsolvar[1] = scal[4];
    for (int i = 0; i < 14; ++i) so.bndsolvar[i] = scal[5 + i];
    Solar sol = compute_solar(so, scal[0], iopt[6], scal[2]);
    Unit units[kMaxUnits];
    const int nunits = build_units(units, CB_SW_UMAX);
    Work W;
    W.ncc = ncol;
    std::vector<double> ws((size_t)NF * nlay * ncol), cld((size_t)42 * nlay * ncol), aer((size_t)42 * nlay * ncol),
        scr((size_t)112 * NSCR * nlay * ncol), srcv((size_t)112 * ncol), part((size_t)nunits * 4 * (nlay + 1) * ncol);
    std::vector<int> idx((size_t)nlay * ncol), lt(ncol), ls((size_t)14 * ncol), ac(ncol);
    std::vector<unsigned> mask((size_t)4 * nlay * ncol, 0u);
    int err = 0;
    W.mask = mask.data(); W.mstride = ncol; W.moff = 0;
    W.ws = ws.data(); W.idx = idx.data(); W.laytrop = lt.data(); W.laysolfr = ls.data(); W.anycld = ac.data();
    W.cld = cld.data(); W.aer = aer.data(); W.scr = scr.data(); W.src = srcv.data(); W.part = part.data(); W.err = &err;
    if (mc && irng == 1) { cb::mcica::mask_mt_host(in.cldfr, ncol, nlay, 112, 4, fl.icld, seed, mask); W.mask = mask.data(); }
    if (mc && irng == 0)
      for (int c = 0; c < ncol; ++c)
        if (cb::mcica::mask_column_kiss(in.play, in.cldfr, ncol, nlay, 112, 4, fl.icld, seed, W.mask, ncol, 0, c)) err = 9;
    for (int c = 0; c < ncol; ++c)
      for (int l = nlay - 1; l >= 0; --l) sw_prep_column<true, false>(T, in, fl, W, 0, c, l, l + 1);
    for (int c = 0; c < ncol; ++c) sw_prep_column<false, true>(T, in, fl, W, 0, c, 0, nlay);
    Unit tunits[kMaxUnits];
    const int ntau = build_units(tunits, CB_SW_TAU_UMAX);
    for (int k2 = 0; k2 < ntau; ++k2) {
      const Unit un = tunits[k2];
#define CASE(B) case B: if (un.u == 4) run_taumol<B, 4>(T, sol, in, W, ncol, un.g0); else run_taumol<B, 2>(T, sol, in, W, ncol, un.g0); break;
      switch (un.band) {
        CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22) CASE(23) CASE(24) CASE(25) CASE(26) CASE(27)
        CASE(28) CASE(29)
      }
#undef CASE
    }
    const bool tile = g_tile != 0;
    if (tile) {
      if (mc) run_tile<true, true>(T, sol, in, fl, W, ncol);
      else if (fl.icld >= 1) run_tile<false, true>(T, sol, in, fl, W, ncol);
      else run_tile<false, false>(T, sol, in, fl, W, ncol);
    }
    for (int k2 = 0; k2 < nunits && !tile; ++k2) {
      const Unit un = units[k2];
      if (un.u == 4) run_transfer<4>(T, sol, in, fl, W, ncol, un.band - 16, un.g0, k2, mc);
      else if (un.u == 1) run_transfer<1>(T, sol, in, fl, W, ncol, un.band - 16, un.g0, k2, mc);
      else run_transfer<2>(T, sol, in, fl, W, ncol, un.band - 16, un.g0, k2, mc);
    }
    if (g_scr_export) std::memcpy(g_scr_export, scr.data(), scr.size() * sizeof(double));
    for (int c = 0; c < ncol; ++c)
      for (int lev = 0; lev <= nlay; ++lev)
        sw_reduce_level(W, tile ? kTileGroups : (nunits + CB_SW_GROUP - 1) / CB_SW_GROUP, nlay, 0, c, lev, ncol, out);
    for (int c = 0; c < ncol; ++c)
      for (int l = 0; l < nlay; ++l) sw_heating(T, in, out, c, l);
    return err;
  } catch (std::exception& e) {
    std::fprintf(stderr, "emul_sw_run: %s\n", e.what());
    return -1;
  }
}

#ifdef CB_COVERAGE
// branch-coverage counters of the taumol evaluator (tools/taumol_coverage.py): [band 0..31][lower][id]
static long g_cov[32][2][cb::COV_N];
extern "C" void cb_cov_hit(int engine, int band, int lower, int id) { if (engine == 1 && band >= 0 && band < 32 && id < cb::COV_N) ++g_cov[band][lower][id]; }
extern "C" void emul_cov_reset() { std::memset(g_cov, 0, sizeof g_cov); }
extern "C" void emul_cov_get(long* out /*[32][2][COV_N]*/) { std::memcpy(out, g_cov, sizeof g_cov); }
extern "C" int emul_cov_n() { return cb::COV_N; }
#endif
