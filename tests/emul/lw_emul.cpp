// TEST-ONLY host emulation of the CUDA LW engine: runs the very same per-thread code
// (climt_b200/csrc/lw_core.cuh) serially on the CPU so the kernel logic can be checked against the
// oracle in a container without a GPU.  Not part of the product; the product never falls back to this.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../climt_b200/csrc/lw_tables.h"
#include "../../climt_b200/csrc/mcica_host.h"

using namespace cb::lw;

template <int B, int U>
static void run_taumol(const Tables& T, const In& in, const Work& W, int n, int g0) {
  const int cut = in.nlay / 3;  // two layer chunks, like two blockIdx.z slices of the kernel
  for (int c = 0; c < n; ++c) {
    lw_taumol_unit<B, U>(T, in, W, 0, c, g0, cut, in.nlay);
    lw_taumol_unit<B, U>(T, in, W, 0, c, g0, 0, cut);
  }
}

template <int U, bool DRV>
static void run_transfer(const Tables& T, const In& in, const Work& W, int n, int ib, int g0, int unit, bool mc, bool mr) {
  for (int c = 0; c < n; ++c) {
    // the units of a group accumulate into the group's (zeroed) rows in unit order, like the warps of a block on the device
    const size_t pstride = (size_t)(in.nlay + 1) * W.ncc;
    LwPartDirect sink{W.part + (size_t)(unit / CB_LW_GROUP) * W.npart * pstride + c, pstride, W.ncc};
    if (mc) lw_transfer_unit<U, true, false, DRV>(T, in, W, 0, c, ib, g0, sink);
    else if (mr) lw_transfer_unit<U, false, true, DRV>(T, in, W, 0, c, ib, g0, sink);
    else lw_transfer_unit<U, false, false, DRV>(T, in, W, 0, c, ib, g0, sink);
  }
}

// The column-tile form of the transfer (lw_core.cuh: lw_tile_cell / lw_tile_sweeps; CUDA kernel k_lw_tile): the cells of one
// g-point are evaluated into NaN-poisoned row buffers, then the column's sweeps run over them -- same code, serial schedule.
static int g_tile = 0;
extern "C" void emul_lw_set_tile(int on) { g_tile = on; }

template <bool MC, bool CLOUDY>
static void run_tile(const Tables& T, const In& in, const Work& W, int n) {
  const int nlay = in.nlay;
  constexpr int NR = CLOUDY ? kTileRowsCloudy : kTileRowsClear;
  const size_t pstride = (size_t)(nlay + 1) * W.ncc;
  std::vector<double> P((size_t)NR * nlay), acc((size_t)4 * (nlay + 1));
  for (int group = 0; group < kTileGroups; ++group) {
    int ib0, ib1;
    lw_tile_group_bands(group, ib0, ib1);
    for (int c = 0; c < n; ++c) {
      std::fill(acc.begin(), acc.end(), 0.0);
      const bool cloudy_col = CLOUDY && W.ncbands[c] > 0;
      for (int ib = ib0; ib < ib1; ++ib) {
        const TileBandCol bc = lw_tile_band_col(W, c, ib);
        const double* tp = T.base + T.totplnk + (size_t)ib * 181;
        const double semiss = in.emis[(size_t)ib * in.ncol + c];
        const double plankbnd = semiss * planck_band(tp, in.tsfc[c]);
        const double wband = 0.5 * T.base[T.delwave + ib];
        for (int g = 0; g < band_ngpt(ib); ++g) {
          std::fill(P.begin(), P.end(), std::nan(""));
          double frac1 = 0.;
          for (int l = nlay - 1; l >= 0; --l) {
            const TileCellCol cc = lw_tile_cell_col<MC>(in, W, 0, c, l);
            const TileCellBand cb_ = lw_tile_cell_band<CLOUDY>(T, in, W, 0, c, l, ib, bc, cc.cloudy);
            double tau, plfrac;
            lw_tile_cell_load(in, W, c, l, band_gstart(ib) + g, tau, plfrac);
            lw_tile_cell<MC, CLOUDY>(T, band_gstart(ib) + g, bc, cc, cb_, tau, plfrac, P.data() + l, (size_t)nlay);
            if (l == 0) frac1 = plfrac;
          }
          const size_t as = (size_t)(nlay + 1);
          // total-sky stream (the cloudy form's rows 3-5; cloud-free form: rows 0-2), then the clear-sky stream of a cloudy column
          lw_tile_sweeps(P.data() + (CLOUDY ? (size_t)TR_TT * nlay : 0), (size_t)nlay, 1, nlay, frac1 * plankbnd, 1. - semiss, wband,
                         acc.data() + as, acc.data(), 1);
          if (cloudy_col)
            lw_tile_sweeps(P.data(), (size_t)nlay, 1, nlay, frac1 * plankbnd, 1. - semiss, wband, acc.data() + 3 * as,
                           acc.data() + 2 * as, 1);
        }
      }
      for (int q = 0; q < 4; ++q) {
        if (q >= 2 && !cloudy_col) continue;
        for (int lev = 0; lev <= nlay; ++lev)
          W.part[((size_t)group * W.npart + q) * pstride + (size_t)lev * W.ncc + c] = acc[(size_t)q * (nlay + 1) + lev];
      }
    }
  }
}

// flags8 = {icld, idrv, inflag, iceflag, liqflag, mcica, irng, permuteseed}
extern "C" int emul_lw_run(const char* blob, const double* consts11, const int* flags5, int ncol, int nlay,
                           const double* const* inp /*23 pointers in struct In order*/,
                           double* const* outp /*8: 6 fluxes / heating rates + duflx_dt, duflxc_dt (used when idrv = 1)*/) {
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
    for (int i = 0; i < 23; ++i) ip[i] = inp[i];
    Out out;
    double** op = &out.uflx;
    for (int i = 0; i < 6; ++i) op[i] = outp[i];
    const bool drv = flags5[1] == 1;
    if (drv) { out.duflx_dt = outp[6]; out.duflxc_dt = outp[7]; }
    Flags fl{flags5[0], flags5[1], flags5[2], flags5[3], flags5[4], flags5[5]};
    const int irng = flags5[6], seed = flags5[7];
    const bool mc = fl.mcica && fl.icld >= 1;
    Unit units[kMaxUnits];
    const int nunits = build_units(units, CB_LW_UMAX);
    Work W;
    W.ncc = ncol;
    std::vector<double> ws((size_t)NF * nlay * ncol), pw(ncol), cld((size_t)32 * nlay * ncol),
        scr((size_t)140 * NSCR * nlay * ncol), ovl((size_t)OV_NROWS * (nlay + 2) * ncol), part((size_t)nunits * 6 * (nlay + 1) * ncol);
    W.npart = drv ? 6 : 4;
    std::vector<int> idx((size_t)nlay * ncol), lt(ncol), ncb(ncol);
    std::vector<unsigned> mask((size_t)5 * nlay * ncol, 0u);
    int err = 0;
    W.mask = mask.data(); W.mstride = ncol; W.moff = 0;
    W.ws = ws.data(); W.idx = idx.data(); W.laytrop = lt.data(); W.ncbands = ncb.data(); W.pwvcm = pw.data();
    W.cld = cld.data(); W.scr = scr.data(); W.ovl = ovl.data(); W.part = part.data(); W.err = &err;
    if (mc && irng == 1) { cb::mcica::mask_mt_host(in.cldfr, ncol, nlay, 140, 5, fl.icld, seed, mask); W.mask = mask.data(); }
    if (mc && irng == 0)
      for (int c = 0; c < ncol; ++c)
        if (cb::mcica::mask_column_kiss(in.play, in.cldfr, ncol, nlay, 140, 5, fl.icld, seed, W.mask, ncol, 0, c)) err = 9;
    for (int c = 0; c < ncol; ++c)
      for (int l = nlay - 1; l >= 0; --l) prep_column<true, false>(T, in, fl, W, 0, c, l, l + 1);
    for (int c = 0; c < ncol; ++c) prep_column<false, true>(T, in, fl, W, 0, c, 0, nlay);
    if (fl.icld >= 1)
      for (int c = 0; c < ncol; ++c)
        for (int l = 0; l < nlay; ++l) prep_cloud_scale(in, fl, W, 0, c, l);
    Unit tunits[kMaxUnits];
    const int ntau = build_units(tunits, CB_LW_TAU_UMAX);
    for (int k2 = 0; k2 < ntau; ++k2) {
      const Unit un = tunits[k2];
#define CASE(B) case B: if (un.u == 4) run_taumol<B, 4>(T, in, W, ncol, un.g0); else run_taumol<B, 2>(T, in, W, ncol, un.g0); break;
      switch (un.band) {
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
        CASE(14) CASE(15) CASE(16)
      }
#undef CASE
    }
    const bool tile = g_tile && !drv && (mc || fl.icld < 2);
    if (tile) {
      if (mc) run_tile<true, true>(T, in, W, ncol);
      else if (fl.icld >= 1) run_tile<false, true>(T, in, W, ncol);
      else run_tile<false, false>(T, in, W, ncol);
    }
    for (int k2 = 0; k2 < nunits && !tile; ++k2) {
      const Unit un = units[k2];
      const bool mr = !mc && fl.icld >= 2;
      if (drv) {
        if (un.u == 4) run_transfer<4, true>(T, in, W, ncol, un.band - 1, un.g0, k2, mc, mr);
        else run_transfer<2, true>(T, in, W, ncol, un.band - 1, un.g0, k2, mc, mr);
      } else {
        if (un.u == 4) run_transfer<4, false>(T, in, W, ncol, un.band - 1, un.g0, k2, mc, mr);
        else run_transfer<2, false>(T, in, W, ncol, un.band - 1, un.g0, k2, mc, mr);
      }
    }
    for (int c = 0; c < ncol; ++c)
      for (int lev = 0; lev <= nlay; ++lev)
        lw_reduce_level(T, W, tile ? kTileGroups : (nunits + CB_LW_GROUP - 1) / CB_LW_GROUP, nlay, 0, c, lev, ncol, out);
    for (int c = 0; c < ncol; ++c)
      for (int l = 0; l < nlay; ++l) lw_heating(T, in, out, c, l);
    return err;
  } catch (std::exception& e) {
    std::fprintf(stderr, "emul_lw_run: %s\n", e.what());
    return -1;
  }
}

#ifdef CB_COVERAGE
// branch-coverage counters of the taumol evaluator (tools/taumol_coverage.py): [band 0..31][lower][id]
static long g_cov[32][2][cb::COV_N];
extern "C" void cb_cov_hit(int engine, int band, int lower, int id) { if (engine == 0 && band >= 0 && band < 32 && id < cb::COV_N) ++g_cov[band][lower][id]; }
extern "C" void emul_cov_reset() { std::memset(g_cov, 0, sizeof g_cov); }
extern "C" void emul_cov_get(long* out /*[32][2][COV_N]*/) { std::memcpy(out, g_cov, sizeof g_cov); }
extern "C" int emul_cov_n() { return cb::COV_N; }
#endif
