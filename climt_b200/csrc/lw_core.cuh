// climt_b200 -- RRTMG longwave engine, per-thread core (sm_100a device code; also host-compilable
// so tests can single-step the very same code on the CPU without a GPU).
//
// Work decomposition (B200-first, not the reference's column-serial loop nest).  Lanes of a warp are always
// ADJACENT COLUMNS, so every access to the caller's state and to the workspace is a contiguous 256-byte row.
//   prep_column<LAYER,COLUMN>  inatm + setcoef per (column, layer); what couples the layers (pwvcm, laytrop, cldprop) per column
//   prep_overlap / prep_cloud_scale   maximum-random overlap factors (rtrnmr) / rtrn cloud prologue
//   lw_taumol_unit<B,U>        one thread per (column, unit of <=4 g-points of band B, chunk of layers): taug + Planck fractions;
//                              all band-specific code is block-uniform, table gathers of neighbouring columns share L1 lines
//   lw_transfer_unit<U,MC,MR>  one thread per (column, unit of <=2 g-points), band-generic: rtrn / rtrnmc / rtrnmr sweeps with
//                              the recurrences in registers
//   lw_reduce_level            one thread per (column, level): fixed-order sum of the per-unit partial fluxes (deterministic,
//                              no atomics), band weights, W m-2
//   lw_heating                 one thread per (column, layer): heating rates
//
// Reference being replaced (cited per function): climt/_lib/rrtmg_lw/rrtmg_lw_rad.nomcica.f90 (inatm, driver),
// rrtmg_lw_setcoef.f90, rrtmg_lw_cldprop.f90, rrtmg_lw_cldprmc.f90, rrtmg_lw_taumol.f90, rrtmg_lw_rtrn.f90, rrtmg_lw_rtrnmc.f90,
// rrtmg_lw_rtrnmr.f90.
// The 16 hand-specialised taugbNN routines are expressed here as ONE generic evaluator driven by a
// compile-time band description (struct Region) -- see region<B,LOWER>().
#pragma once
#include <cmath>

#include "cb_common.h"

namespace cb {
namespace lw {

constexpr int NBND = 16, NGPT = 140, NTBL = 10000;
constexpr int MAXU = 4;
// workspace fields per (layer, column)
enum WsField {
  F_FAC00 = 0, F_FAC01, F_FAC10, F_FAC11,
  F_COLH2O, F_COLCO2, F_COLO3, F_COLN2O, F_COLCO, F_COLCH4, F_COLO2,
  F_COLBRD, F_COLDRY, F_WX1, F_WX2, F_WX3, F_WX4,
  F_SELFFAC, F_SELFFRAC, F_FORFAC, F_FORFRAC, F_MINORFRAC, F_SCALEMINOR, F_SCALEMINORN2,
  NF
};

// Layout of a band's g-point tables in HBM: for every group of TGW = 4 consecutive g-points ONE contiguous block that holds every
// table of the band, row after row, TGW doubles per row (bands whose g-point count is not a multiple of 4 pad the last group):
//   element (table X, row r, g-point g)  ->  base + (g / 4) * rows * 4 + (X + r) * 4 + (g % 4)
// A taumol thread evaluates <= 4 g-points of one group: every table row it touches is one aligned 32-byte vector, neighbouring
// rows are neighbouring sectors of the same 128-byte line, and the whole block of a single-key-species band (<= 14 KB) is one
// cp.async.bulk copy into shared memory (lw_engine.cu: k_lw_taumol).
constexpr int TGW = 4;
struct BandOff {
  int base, rows;  // first block (offset into Tables::base, in doubles) and rows per block
  int absa, absb, selfref, forref, fracrefa, fracrefb;  // ROW offsets inside a block (-1: the band has no such table)
  int m[5];  // minor-gas tables (slot meaning per band: see kMinorNames in lw_tables.h)
  int x[2];  // cross-section vectors
  double refrat_planck_a, refrat_planck_b, refrat_m_a, refrat_m_b, refrat_m_a3;
};

struct Tables {
  const double* base;  // one HBM buffer holding every table, offsets below are in doubles
  BandOff b[NBND];
  int chi_mls, preflog, tref, rat, totplnk, totplnkderiv, delwave;
  int tau_tbl, exp_tbl, tfn_tbl;
  int et_tbl;  // interleaved {exp_tbl[i], tfn_tbl[i]} pairs: one 128-bit gather instead of two 64-bit ones
  int absice0, absice1, absice2, absice3, absliq1;
  double abscld1, absliq0;
  double bpade, heatfac, fluxfac, oneminus, avogad, grav;
};

struct In {  // reference ABI layout: (nlay[+1], ncol) column-fastest; emis (16,ncol); taucld (nlay,ncol,16); tauaer (16,nlay,ncol)
  int ncol, nlay;
  const double *play, *plev, *tlay, *tlev, *tsfc, *h2o, *o3, *co2, *ch4, *n2o, *o2, *cfc11, *cfc12, *cfc22, *ccl4,
      *emis, *cldfr, *taucld, *cicewp, *cliqwp, *reice, *reliq, *tauaer;
};
struct Out {
  double *uflx, *dflx, *hr, *uflxc, *dflxc, *hrc;
  double *duflx_dt = nullptr, *duflxc_dt = nullptr;  // idrv = 1: d(upward flux)/d(surface temperature), (nlay+1, ncol)
};
struct Flags {
  int icld, idrv, inflag, iceflag, liqflag;
  int mcica;  // 0: rtrn (band cloud optics, fractional cloud); 1: McICA (rtrnmc: per-g-point 0/1 cloud mask)
};
struct Work {  // all sized for a chunk of ncc columns
  int ncc;
  double* ws;    // [NF][nlay][ncc]
  int* idx;      // [nlay][ncc]  packed jp|jt|jt1|indself|indfor|indminor
  int* laytrop;  // [ncc]
  int* ncbands;  // [ncc]   0 = column has no cloudy layer
  double* pwvcm; // [ncc]
  double* cld;   // [2][16][nlay][ncc]  odcld, efclfrac
  double* scr;   // [140][NSCR][nlay][ncc] atrans, bbugas, atot, bbutot, taug, fracs
  double* part;  // [ngroups][npart][nlay+1][ncc] up, dn, upclr, dnclr [, d_up, d_upclr]  (band-weighted sums over a group's units)
  int npart;     // 4, or 6 with the surface-temperature derivative of the upward flux (idrv = 1)
  double* ovl;   // [14][nlay+2][ncc] maximum-random overlap factors of rtrnmr (OV_* rows), non-McICA icld = 2, 3 only
  unsigned* mask; // [nlay][5][mstride] McICA cloud mask (+ moff), bit (g & 31) of word (g >> 5) set = sub-column g cloudy
  int mstride, moff;
  int* err;      // [1]
};

CB_HD int pack_idx(int jp, int jt, int jt1, int inds, int indf, int indm) {
  return jp | (jt << 6) | (jt1 << 9) | (inds << 12) | (indf << 16) | (indm << 18);
}

// ---------------------------------------------------------------------------------------------
// prep_column: inatm (rrtmg_lw_rad.nomcica.f90:572-900) + setcoef (rrtmg_lw_setcoef.f90:257-412) +
// cldprop (rrtmg_lw_cldprop.f90:31-276) + the cloud prologue of rtrn (rrtmg_lw_rtrn.f90:260-318).
CB_HD double fmax2(double a, double b) { return a > b ? a : b; }
CB_HD double fmin2(double a, double b) { return a < b ? a : b; }
CB_HD double secdiff_band(double pwvcm, int ib /*0-based*/) {
  // a0 + a1*exp(a2*pwvcm) clamped to [1.5, 1.8] for bands 2-3 and 5-9, 1.66 otherwise (rrtmg_lw_rtrn.f90:246-270).
  // A switch, not indexed arrays: ib is a run-time value in the band-generic transfer kernel and indexed local arrays
  // would live in (and be initialised into) local memory by every thread.
  double a0, a1, a2;
  switch (ib) {
    case 1: a0 = 1.55; a1 = 0.25; a2 = -12.0; break;
    case 2: a0 = 1.58; a1 = 0.22; a2 = -11.7; break;
    case 4: a0 = 1.54; a1 = 0.13; a2 = -0.72; break;
    case 5: a0 = 1.454; a1 = 0.446; a2 = -0.243; break;
    case 6: a0 = 1.89; a1 = -0.10; a2 = 0.19; break;
    case 7: a0 = 1.33; a1 = 0.40; a2 = -0.062; break;
    case 8: a0 = 1.668; a1 = -0.006; a2 = 0.414; break;
    default: return 1.66;
  }
  double s = a0 + a1 * exp(a2 * pwvcm);
  if (s > 1.80) s = 1.80;
  if (s < 1.50) s = 1.50;
  return s;
}

CB_HD void prep_overlap(const In& in, const Work& W, int c0, int c);  // below

// Two launch modes (r01 ncu: the one-thread-per-column version cost as much as a transfer kernel -- 64 blocks on 148 SMs,
// 60 serial layers): LAYER_PART = inatm + setcoef of layers [l0, l1), independent per layer (one thread per (column, layer));
// COLUMN_PART = what couples the layers of a column: pwvcm, laytrop, cldprop / cldprmc (whose routine-locals persist from
// layer to layer in the Fortran) and the rtrn cloud prologue.  <true, true> over [0, nlay) is the original single pass.
template <bool LAYER_PART, bool COLUMN_PART>
CB_HD void prep_column(const Tables& T, const In& in, const Flags& fl, const Work& W, int c0, int c, int l0, int l1) {
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const size_t gc = (size_t)(c0 + c);
  const double* tb = T.base;
  const double amd = 28.9660, amw = 18.0160;
  const double stpfac = 296. / 1013.;
  bool clouds = fl.icld >= 1;
  double amttl = 0.0, wvttl = 0.0;
  int laytrop = 0;
  int ncbands = 1;
  bool anycld = false;
  // cldprop state that persists across layers exactly as in the Fortran (abscoice/abscoliq/iceind/liqind are
  // routine-local and keep their last value)
  double abscoice[16], abscoliq[16];
  for (int i = 0; i < 16; ++i) { abscoice[i] = 0.; abscoliq[i] = 0.; }
  int iceind = 0, liqind = 0;
  if (COLUMN_PART && clouds) {
    // no layer reaches the cloud threshold -> nothing downstream reads the cloud optics: skip cldprop altogether
    int any = 0;  // no short circuit: the loads of different layers stay independent of each other
    const double thr = fl.mcica ? 1.e-20 : 1.e-6;
#pragma unroll 8
    for (int l = 0; l < nlay; ++l) any |= (int)(CB_LDG(in.cldfr + (size_t)l * ncol + gc) >= thr);
    clouds = any != 0;
  }
  if (COLUMN_PART && !LAYER_PART && !clouds) {
    // Cloud-free column: what is left of the column pass are three sums over the layers.  No store in the loop, so the loads of
    // several layers are in flight together; the additions keep the layer order (bit-identical to the general loop below).
    // (the loads of 8 layers are issued together, then consumed: left to itself the compiler keeps one layer's loads per trip)
    constexpr int NB = 8;
    for (int l = 0; l < nlay; l += NB) {
      double h[NB], pa[NB], pb[NB], pm[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const size_t o = (size_t)(l + j < nlay ? l + j : nlay - 1) * ncol + gc;
        h[j] = CB_LDG(in.h2o + o); pa[j] = CB_LDG(in.plev + o); pb[j] = CB_LDG(in.plev + o + ncol); pm[j] = CB_LDG(in.play + o);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (l + j >= nlay) break;
        const double h2o = h[j];
        const double amm = (1. - h2o) * amd + h2o * amw;
        const double coldry = (pa[j] - pb[j]) * 1.e3 * T.avogad / (1.e2 * T.grav * amm * (1. + h2o));
        const double w0 = coldry * h2o;
        amttl = amttl + coldry + w0;
        wvttl = wvttl + w0;
        if (!(log(pm[j]) <= 4.56)) laytrop = laytrop + 1;
      }
    }
    const double wvsh = (amw * wvttl) / (amd * amttl);
    W.pwvcm[c] = wvsh * (1.e3 * in.plev[gc]) / (1.e2 * T.grav);
    W.laytrop[c] = laytrop;
    W.ncbands[c] = 0;
    return;
  }
#define WS(f, l) W.ws[((size_t)(f) * nlay + (l)) * ncc + c]
  for (int l = l0; l < l1; ++l) {
    const size_t o = (size_t)l * ncol + gc;
    const double pavel = in.play[o], tavel = in.tlay[o];
    const double pz_below = in.plev[o];  // pz(l-1)
    const double pz = in.plev[o + ncol];
    double wkl[7];
    wkl[0] = in.h2o[o]; wkl[1] = in.co2[o]; wkl[2] = in.o3[o]; wkl[3] = in.n2o[o];
    wkl[4] = 0.0; wkl[5] = in.ch4[o]; wkl[6] = in.o2[o];
    const double amm = (1. - wkl[0]) * amd + wkl[0] * amw;
    const double coldry = (pz_below - pz) * 1.e3 * T.avogad / (1.e2 * T.grav * amm * (1. + wkl[0]));
    double summol = 0.0;
    for (int i = 1; i < 7; ++i) summol = summol + wkl[i];
    const double wbrodl = coldry * (1. - summol);
    for (int i = 0; i < 7; ++i) wkl[i] = coldry * wkl[i];
    if (COLUMN_PART) {
      amttl = amttl + coldry + wkl[0];
      wvttl = wvttl + wkl[0];
    }
    const double plog = log(pavel);
    if (COLUMN_PART && !(plog <= 4.56)) laytrop = laytrop + 1;  // rrtmg_lw_setcoef.f90:293
    double factor = 0.;
    if (LAYER_PART) {
      WS(F_WX1, l) = coldry * in.ccl4[o] * 1.e-20;
      WS(F_WX2, l) = coldry * in.cfc11[o] * 1.e-20;
      WS(F_WX3, l) = coldry * in.cfc12[o] * 1.e-20;
      WS(F_WX4, l) = coldry * in.cfc22[o] * 1.e-20;
      // ---- setcoef, rrtmg_lw_setcoef.f90:257-412
      int jp = (int)(CB_MULADD_2R(-5., (plog + 0.04), 36.));
      if (jp < 1) jp = 1; else if (jp > 58) jp = 58;
      const double fp = 5. * (tb[T.preflog + jp - 1] - plog);
      const double tr0 = tb[T.tref + jp - 1], tr1 = tb[T.tref + jp];
      int jt = (int)(3. + (tavel - tr0) / 15.);
      if (jt < 1) jt = 1; else if (jt > 4) jt = 4;
      const double ft = ((tavel - tr0) / 15.) - (double)(jt - 3);
      int jt1 = (int)(3. + (tavel - tr1) / 15.);
      if (jt1 < 1) jt1 = 1; else if (jt1 > 4) jt1 = 4;
      const double ft1 = ((tavel - tr1) / 15.) - (double)(jt1 - 3);
      const double water = wkl[0] / coldry;
      const double scalefac = pavel * stpfac / tavel;
      double forfac = scalefac / (1. + water), forfrac, selffac = water * forfac, selffrac = 0.;
      int indfor, indself = 1;
      if (!(plog <= 4.56)) {
        factor = (332.0 - tavel) / 36.0;
        indfor = imin(2, imax(1, (int)factor));
        forfrac = factor - (double)indfor;
        factor = (tavel - 188.0) / 7.2;
        indself = imin(9, imax(1, (int)factor - 7));
        selffrac = factor - (double)(indself + 7);
      } else {
        factor = (tavel - 188.0) / 36.0;
        indfor = 3;
        forfrac = factor - 1.0;
      }
      const double scaleminor = pavel / tavel;
      const double scaleminorn2 = (pavel / tavel) * (wbrodl / (coldry + wkl[0]));
      factor = (tavel - 180.8) / 7.2;
      const int indminor = imin(18, imax(1, (int)factor));
      const double minorfrac = factor - (double)indminor;
      double colg[7];
      for (int i = 0; i < 7; ++i) colg[i] = 1.e-20 * wkl[i];
      if (colg[1] == 0.) colg[1] = 1.e-32 * coldry;
      if (colg[2] == 0.) colg[2] = 1.e-32 * coldry;
      if (colg[3] == 0.) colg[3] = 1.e-32 * coldry;
      if (colg[4] == 0.) colg[4] = 1.e-32 * coldry;
      if (colg[5] == 0.) colg[5] = 1.e-32 * coldry;
      const double compfp = 1. - fp;
      WS(F_FAC10, l) = compfp * ft;
      WS(F_FAC00, l) = compfp * (1. - ft);
      WS(F_FAC11, l) = fp * ft1;
      WS(F_FAC01, l) = fp * (1. - ft1);
      for (int i = 0; i < 7; ++i) WS(F_COLH2O + i, l) = colg[i];
      WS(F_COLBRD, l) = 1.e-20 * wbrodl;
      WS(F_COLDRY, l) = coldry;
      WS(F_SELFFAC, l) = colg[0] * selffac;
      WS(F_SELFFRAC, l) = selffrac;
      WS(F_FORFAC, l) = colg[0] * forfac;
      WS(F_FORFRAC, l) = forfrac;
      WS(F_MINORFRAC, l) = minorfrac;
      WS(F_SCALEMINOR, l) = scaleminor;
      WS(F_SCALEMINORN2, l) = scaleminorn2;
      W.idx[(size_t)l * ncc + c] = pack_idx(jp, jt, jt1, indself, indfor, indminor);
    }
    // ---- McICA: cldprmc (rrtmg_lw_cldprmc.f90:160-247).  Every cloudy sub-column of a layer carries the layer's
    // water paths, so the per-g-point optical depth only depends on the g-point's band: one value per (layer, band).
    if (LAYER_PART && fl.icld >= 1 && fl.mcica) {  // layer-independent: no routine-local survives from layer to layer here
      const double cldmin = 1.e-20;
      const double ciwp = in.cicewp[o], clwp = in.cliqwp[o];
      const double* tc = in.taucld ? in.taucld + 16 * ((size_t)l * ncol + gc) : nullptr;  // null = all zero
      const double cwp = ciwp + clwp;
      const int pat5[16] = {0, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4};
      for (int ib = 0; ib < 16; ++ib) {
        double tau = tc ? tc[ib] : 0.0;
        if (cwp >= cldmin || tau >= cldmin) {
          if (fl.inflag == 1) *W.err = 8;  // 'INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA'
          if (fl.inflag == 2) {
            double aice = 0.0, aliq = 0.0;
            const double radice = in.reice[o];
            if (ciwp == 0.0) aice = 0.0;
            else if (fl.iceflag == 0) {
              if (radice < 10.0) *W.err = 1;
              aice = tb[T.absice0] + tb[T.absice0 + 1] / radice;
            } else if (fl.iceflag == 1) {
              if (radice < 13.0 || radice > 130.) *W.err = 2;
              aice = tb[T.absice1 + 2 * pat5[ib]] + tb[T.absice1 + 2 * pat5[ib] + 1] / radice;
            } else if (fl.iceflag == 2 || fl.iceflag == 3) {
              const double rmax = fl.iceflag == 2 ? 131.0 : 140.0;
              if (radice < 5.0 || radice > rmax) { *W.err = fl.iceflag == 2 ? 2 : 3; }
              else {
                factor = (radice - 2.) / 3.;
                int index = (int)factor;
                const int top = fl.iceflag == 2 ? 43 : 46;
                if (index == top) index = top - 1;
                const double fint = factor - (double)index;
                const int t0 = fl.iceflag == 2 ? T.absice2 : T.absice3;
                const double k0 = tb[t0 + (index - 1) * 16 + ib], k1 = tb[t0 + index * 16 + ib];
                aice = k0 + fint * (k1 - (k0));
              }
            }
            if (clwp == 0.0) aliq = 0.0;
            else if (fl.liqflag == 0) aliq = T.absliq0;
            else if (fl.liqflag == 1) {
              const double radliq = in.reliq[o];
              if (radliq < 2.5 || radliq > 60.) { *W.err = 4; }
              else {
                int index = (int)(radliq - 1.5);
                if (index == 0) index = 1;
                if (index == 58) index = 57;
                const double fint = radliq - 1.5 - (double)index;
                const double k0 = tb[T.absliq1 + (index - 1) * 16 + ib], k1 = tb[T.absliq1 + index * 16 + ib];
                aliq = k0 + fint * (k1 - (k0));
              }
            }
            tau = ciwp * aice + clwp * aliq;
          }
        }
        W.cld[((size_t)ib * nlay + l) * ncc + c] = tau;
      }
    }
    // ---- cldprop for this layer, rrtmg_lw_cldprop.f90:163-270 (taucloud parked in W.cld slot 0)
    if (COLUMN_PART && clouds && !fl.mcica) {
      double taucloud[16];
      for (int ib = 0; ib < 16; ++ib) taucloud[ib] = 0.0;
      const double cldfrac = in.cldfr[o];
      const double ciwp = in.cicewp[o], clwp = in.cliqwp[o];
      double tauctot = 0.0;
      const double* tc = in.taucld ? in.taucld + 16 * ((size_t)l * ncol + gc) : nullptr;  // null = all zero
      if (tc) for (int ib = 0; ib < 16; ++ib) tauctot = tauctot + tc[ib];
      const double cwp = ciwp + clwp;
      const double cldmin = 1.e-20;
      if (cldfrac >= cldmin && (cwp >= cldmin || tauctot >= cldmin)) {
        if (fl.inflag == 0) {
          ncbands = 16;
          for (int ib = 0; ib < 16; ++ib) taucloud[ib] = tc ? tc[ib] : 0.0;
        } else if (fl.inflag == 1) {
          ncbands = 16;
          for (int ib = 0; ib < 16; ++ib) taucloud[ib] = T.abscld1 * cwp;
        } else if (fl.inflag == 2) {
          const double radice = in.reice[o];
          if (ciwp == 0.0) {
            abscoice[0] = 0.0;
            iceind = 0;
          } else if (fl.iceflag == 0) {
            if (radice < 10.0) *W.err = 1;
            abscoice[0] = tb[T.absice0] + tb[T.absice0 + 1] / radice;
            iceind = 0;
          } else if (fl.iceflag == 1) {
            if (radice < 13.0 || radice > 130.) *W.err = 2;
            ncbands = 5;
            for (int ib = 0; ib < 5; ++ib) abscoice[ib] = tb[T.absice1 + 2 * ib] + tb[T.absice1 + 2 * ib + 1] / radice;
            iceind = 1;
          } else if (fl.iceflag == 2) {
            if (radice < 5.0 || radice > 131.0) { *W.err = 2; }
            else {
              ncbands = 16;
              factor = (radice - 2.) / 3.;
              int index = (int)factor;
              if (index == 43) index = 42;
              const double fint = factor - (double)index;
              for (int ib = 0; ib < 16; ++ib) {
                const double k0 = tb[T.absice2 + (index - 1) * 16 + ib], k1 = tb[T.absice2 + index * 16 + ib];
                abscoice[ib] = k0 + fint * (k1 - (k0));
              }
              iceind = 2;
            }
          } else if (fl.iceflag == 3) {
            if (radice < 5.0 || radice > 140.0) { *W.err = 3; }
            else {
              ncbands = 16;
              factor = (radice - 2.) / 3.;
              int index = (int)factor;
              if (index == 46) index = 45;
              const double fint = factor - (double)index;
              for (int ib = 0; ib < 16; ++ib) {
                const double k0 = tb[T.absice3 + (index - 1) * 16 + ib], k1 = tb[T.absice3 + index * 16 + ib];
                abscoice[ib] = k0 + fint * (k1 - (k0));
              }
              iceind = 2;
            }
          }
          if (clwp == 0.0) {
            abscoliq[0] = 0.0;
            liqind = 0;
            if (iceind == 1) iceind = 2;
          } else if (fl.liqflag == 0) {
            abscoliq[0] = T.absliq0;
            liqind = 0;
            if (iceind == 1) iceind = 2;
          } else if (fl.liqflag == 1) {
            const double radliq = in.reliq[o];
            if (radliq < 2.5 || radliq > 60.) { *W.err = 4; }
            else {
              int index = (int)(radliq - 1.5);
              if (index == 0) index = 1;
              if (index == 58) index = 57;
              const double fint = radliq - 1.5 - (double)index;
              ncbands = 16;
              for (int ib = 0; ib < 16; ++ib) {
                const double k0 = tb[T.absliq1 + (index - 1) * 16 + ib], k1 = tb[T.absliq1 + index * 16 + ib];
                abscoliq[ib] = k0 + fint * (k1 - (k0));
              }
              liqind = 2;
            }
          }
          for (int ib = 0; ib < ncbands; ++ib) {
            // icb(ib, ind): ind 0 -> 1 ; ind 1 -> {1,2,3,3,3,4,4,4,5,...}; ind 2 -> ib
            const int pat5[16] = {0, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4};
            const int ii = iceind == 0 ? 0 : (iceind == 1 ? pat5[ib] : ib);
            const int il = liqind == 0 ? 0 : (liqind == 1 ? pat5[ib] : ib);
            taucloud[ib] = ciwp * abscoice[ii] + clwp * abscoliq[il];
          }
        }
      }
      if (cldfrac >= 1.e-6) anycld = true;
      for (int ib = 0; ib < 16; ++ib) W.cld[((size_t)ib * nlay + l) * ncc + c] = taucloud[ib];
    }
  }
#undef WS
  if (!COLUMN_PART) return;
  if (clouds && fl.mcica) { anycld = true; ncbands = 16; }  // `clouds` already means: some layer has cldfr >= 1e-20 (:247)
  const double wvsh = (amw * wvttl) / (amd * amttl);
  const double pwvcm = wvsh * (1.e3 * in.plev[gc]) / (1.e2 * T.grav);
  W.pwvcm[c] = pwvcm;
  W.laytrop[c] = laytrop;
  W.ncbands[c] = (clouds && anycld) ? ncbands : 0;
  if (clouds && anycld && !fl.mcica && fl.icld >= 2) prep_overlap(in, W, c0, c);
}

// Maximum-random overlap factors of one column (rrtmg_lw_rtrnmr.f90:322-479): functions of the cloud-fraction profile only,
// shared by every g-point.  Rows: up sweep (index lev+1 = 1..nlay+1) and down sweep (index lev-1 = 0..nlay).
enum OvlRow { OV_CLD1 = 0, OV_CLD2, OV_CLR1, OV_CLR2, OV_CMB1, OV_CMB2, OV_CLD1D, OV_CLD2D, OV_CLR1D, OV_CLR2D, OV_CMB1D, OV_CMB2D,
              OV_IST, OV_ISTD, OV_NROWS };
CB_HD void prep_overlap(const In& in, const Work& W, int c0, int c) {
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const size_t gc = (size_t)(c0 + c);
  const size_t rs = (size_t)(nlay + 2) * ncc;
#define OV(r, i) W.ovl[(size_t)(r) * rs + (size_t)(i) * ncc + c]
  // cldfrac(0) is read by the Fortran at :400-401 (one element before the array, always multiplied by zero): taken as 0
  auto cf = [&](int lev) { return lev >= 1 && lev <= nlay ? in.cldfr[(size_t)(lev - 1) * ncol + gc] : 0.0; };
  auto cloudy = [&](int lev) { return cf(lev) >= 1.e-6; };
  for (int r = 0; r < OV_NROWS; ++r)
    for (int i = 0; i <= nlay + 1; ++i) OV(r, i) = 0.0;
  double rat1 = 0., rat2 = 0.;
  OV(OV_IST, 1) = 1.;
  OV(OV_ISTD, nlay) = 1.;
  for (int lev = 1; lev <= nlay; ++lev) {
    if (cloudy(lev)) {
      OV(OV_IST, lev + 1) = 0.;
      double cld1 = 0., cld2 = 0., clr1 = 0., clr2 = 0.;
      if (lev == nlay) {
        // all six factors of lev+1 are zero
      } else if (cf(lev + 1) >= cf(lev)) {
        if (OV(OV_IST, lev) == 1.) {
          if (cf(lev) < 1.) clr2 = (cf(lev + 1) - cf(lev)) / (1. - cf(lev));
          OV(OV_CLR2, lev) = 0.;
          OV(OV_CLD2, lev) = 0.;
        } else {
          const double fmax = fmax2(cf(lev), cf(lev - 1));
          if (cf(lev + 1) > fmax) {
            clr1 = rat2;
            clr2 = (cf(lev + 1) - fmax) / (1. - fmax);
          } else if (cf(lev + 1) < fmax) {
            clr1 = (cf(lev + 1) - cf(lev)) / (cf(lev - 1) - cf(lev));
          } else {
            clr1 = rat2;
          }
        }
        if (clr1 > 0. || clr2 > 0.) { rat1 = 1.; rat2 = 0.; }
        else { rat1 = 0.; rat2 = 0.; }
      } else {
        if (OV(OV_IST, lev) == 1.) {
          cld2 = (cf(lev) - cf(lev + 1)) / cf(lev);
          OV(OV_CLR2, lev) = 0.;
          OV(OV_CLD2, lev) = 0.;
        } else {
          const double fmin = fmin2(cf(lev), cf(lev - 1));
          if (cf(lev + 1) <= fmin) {
            cld1 = rat1;
            cld2 = (fmin - cf(lev + 1)) / fmin;
          } else {
            cld1 = (cf(lev) - cf(lev + 1)) / (cf(lev) - fmin);
          }
        }
        if (cld1 > 0. || cld2 > 0.) { rat1 = 0.; rat2 = 1.; }
        else { rat1 = 0.; rat2 = 0.; }
      }
      OV(OV_CLD1, lev + 1) = cld1; OV(OV_CLD2, lev + 1) = cld2; OV(OV_CLR1, lev + 1) = clr1; OV(OV_CLR2, lev + 1) = clr2;
      OV(OV_CMB1, lev + 1) = clr1 * OV(OV_CLD2, lev) * cf(lev - 1);
      OV(OV_CMB2, lev + 1) = cld1 * OV(OV_CLR2, lev) * (1. - cf(lev - 1));
    } else {
      OV(OV_IST, lev + 1) = 1.;
    }
  }
  for (int lev = nlay; lev >= 1; --lev) {
    if (cloudy(lev)) {
      OV(OV_ISTD, lev - 1) = 0.;
      double cld1 = 0., cld2 = 0., clr1 = 0., clr2 = 0.;
      if (lev == 1) {
      } else if (cf(lev - 1) >= cf(lev)) {
        if (OV(OV_ISTD, lev) == 1.) {
          if (cf(lev) < 1.) clr2 = (cf(lev - 1) - cf(lev)) / (1. - cf(lev));
          OV(OV_CLR2D, lev) = 0.;
          OV(OV_CLD2D, lev) = 0.;
        } else {
          const double fmax = fmax2(cf(lev), cf(lev + 1));
          if (cf(lev - 1) > fmax) {
            clr1 = rat2;
            clr2 = (cf(lev - 1) - fmax) / (1. - fmax);
          } else if (cf(lev - 1) < fmax) {
            clr1 = (cf(lev - 1) - cf(lev)) / (cf(lev + 1) - cf(lev));
          } else {
            clr1 = rat2;
          }
        }
        if (clr1 > 0. || clr2 > 0.) { rat1 = 1.; rat2 = 0.; }
        else { rat1 = 0.; rat2 = 0.; }
      } else {
        if (OV(OV_ISTD, lev) == 1.) {
          cld2 = (cf(lev) - cf(lev - 1)) / cf(lev);
          OV(OV_CLR2D, lev) = 0.;
          OV(OV_CLD2D, lev) = 0.;
        } else {
          const double fmin = fmin2(cf(lev), cf(lev + 1));
          if (cf(lev - 1) <= fmin) {
            cld1 = rat1;
            cld2 = (fmin - cf(lev - 1)) / fmin;
          } else {
            cld1 = (cf(lev) - cf(lev - 1)) / (cf(lev) - fmin);
          }
        }
        if (cld1 > 0. || cld2 > 0.) { rat1 = 0.; rat2 = 1.; }
        else { rat1 = 0.; rat2 = 0.; }
      }
      OV(OV_CLD1D, lev - 1) = cld1; OV(OV_CLD2D, lev - 1) = cld2; OV(OV_CLR1D, lev - 1) = clr1; OV(OV_CLR2D, lev - 1) = clr2;
      OV(OV_CMB1D, lev - 1) = clr1 * OV(OV_CLD2D, lev) * cf(lev + 1);
      OV(OV_CMB2D, lev - 1) = cld1 * OV(OV_CLR2D, lev) * (1. - cf(lev + 1));
    } else {
      OV(OV_ISTD, lev - 1) = 1.;
    }
  }
#undef OV
}

// rtrn prologue for one (column, layer), rrtmg_lw_rtrn.f90:300-316: cloud optical depth along the diffusivity angle and
// effective cloud fraction per cloud band.  Runs after prep_column (needs pwvcm and the column's ncbands).
// Note: secdiff is indexed by the CLOUD band index there.
CB_HD void prep_cloud_scale(const In& in, const Flags& fl, const Work& W, int c0, int c, int l) {
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const int ncbands = W.ncbands[c];
  if (ncbands == 0) return;
  const double pwvcm = W.pwvcm[c];
  const double cldfrac = in.cldfr[(size_t)l * ncol + (c0 + c)];
  for (int ib = 0; ib < 16; ++ib) {
    const size_t o0 = ((size_t)ib * nlay + l) * ncc + c;
    const size_t o1 = ((size_t)(16 + ib) * nlay + l) * ncc + c;
    if (ib < ncbands && (fl.mcica || cldfrac >= 1.e-6)) {
      const double od = secdiff_band(pwvcm, ib) * W.cld[o0];
      const double transcld = exp(-od);
      const double abscld = 1. - transcld;
      W.cld[o0] = od;
      W.cld[o1] = fl.mcica ? abscld : abscld * cldfrac;  // McICA: efclfrac = abscld * 1 (rtrnmc.f90:307)
    } else {
      W.cld[o0] = 0.0;
      W.cld[o1] = 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Band descriptions.  Gas ids index WsField F_COLH2O + id.
enum Gas { H2O = 0, CO2 = 1, O3 = 2, N2O = 3, CO = 4, CH4 = 5, O2 = 6 };
enum RatId { R_H2OCO2 = 0, R_H2OO3 = 1, R_H2ON2O = 2, R_H2OCH4 = 3, R_N2OCO2 = 4, R_O3CO2 = 5 };
enum AmtKind { A_COL, A_ADJ, A_BRD_N2, A_BRD, A_O2SC };
enum RefR { RM_A, RM_B, RM_A3 };

struct Minor {
  int slot;     // BandOff::m[slot]
  bool binary;  // table is (nsp, 19, ng) and interpolated in the key-species ratio too
  int refr;     // which refrat_m_* drives the binary interpolation
  int amt;      // AmtKind
  int gas;      // for A_COL / A_ADJ
  double thr, base, expo;  // A_ADJ: if ratio > thr: adj = base + (ratio-base)**expo
  bool ref355;  // band 13: reference mixing ratio is the literal 3.55e-4 instead of chi_mls(gas, jp+1)
  bool e20f;    // `1.e20` written as a default-real literal in the Fortran
};
struct Region {
  int kind;  // 0: no key species, 1: one key species, 2: two key species
  int a, b, rat;
  bool self, forn;
  int nminor;
  Minor m[3];
  int nx;
  int xslot[2], xwx[2];
  int corr;    // 0 none, 1: band-1 lower, 2: band-1 upper, 3: band-2 lower
  int planck;  // 0: fixed fracref, 1: interpolated in the key-species ratio, 2: zero
  int scale;   // 0 none, 1: band-4 upper, 2: band-7 upper
};

constexpr Minor kNoMinor = {0, false, RM_A, A_COL, H2O, 0., 0., 0., false, false};
constexpr Minor minor_col(int slot, bool bin, int refr, int gas) { return {slot, bin, refr, A_COL, gas, 0., 0., 0., false, false}; }
constexpr Minor minor_adj(int slot, bool bin, int refr, int gas, double thr, double base, double expo, bool e20f = false,
                          bool ref355 = false) {
  return {slot, bin, refr, A_ADJ, gas, thr, base, expo, ref355, e20f};
}
constexpr Minor minor_amt(int slot, bool bin, int refr, int amt) { return {slot, bin, refr, amt, H2O, 0., 0., 0., false, false}; }
constexpr Region R0(int planck) { return {0, 0, 0, 0, false, false, 0, {kNoMinor, kNoMinor, kNoMinor}, 0, {0, 0}, {0, 0}, 0, planck, 0}; }

// rrtmg_lw_taumol.f90: taugb1 :280-376, taugb2 :379-476, taugb3 :479-763, taugb4 :766-1018, taugb5 :1021-1300,
// taugb6 :1303-1391, taugb7 :1394-1653, taugb8 :1656-1791, taugb9 :1794-2040, taugb10 :2043-2110,
// taugb11 :2113-2195, taugb12 :2198-2395, taugb13 :2398-2650, taugb14 :2653-2718, taugb15 :2721-2936,
// taugb16 :2939-3145.
template <int B, bool LOWER>
constexpr Region region() {
  Region r = R0(0);
  if (B == 1) {
    r.kind = 1; r.a = H2O; r.self = LOWER; r.forn = true;
    r.nminor = 1; r.m[0] = minor_amt(LOWER ? 0 : 1, false, RM_A, A_BRD_N2);
    r.corr = LOWER ? 1 : 2;
  } else if (B == 2) {
    r.kind = 1; r.a = H2O; r.self = LOWER; r.forn = true; r.corr = LOWER ? 3 : 0;
  } else if (B == 3) {
    r.kind = 2; r.a = H2O; r.b = CO2; r.rat = R_H2OCO2; r.self = LOWER; r.forn = true;
    r.nminor = 1; r.m[0] = minor_adj(LOWER ? 0 : 1, true, LOWER ? RM_A : RM_B, N2O, 1.5, 0.5, 0.65, !LOWER);
    r.planck = 1;
  } else if (B == 4) {
    r.kind = 2; r.a = LOWER ? H2O : O3; r.b = CO2; r.rat = LOWER ? R_H2OCO2 : R_O3CO2; r.self = LOWER; r.forn = LOWER;
    r.planck = 1; r.scale = LOWER ? 0 : 1;
  } else if (B == 5) {
    r.kind = 2; r.a = LOWER ? H2O : O3; r.b = CO2; r.rat = LOWER ? R_H2OCO2 : R_O3CO2; r.self = LOWER; r.forn = LOWER;
    if (LOWER) { r.nminor = 1; r.m[0] = minor_col(0, true, RM_A, O3); }
    r.nx = 1; r.xslot[0] = 0; r.xwx[0] = 0;
    r.planck = 1;
  } else if (B == 6) {
    if (LOWER) {
      r.kind = 1; r.a = H2O; r.self = true; r.forn = true;
      r.nminor = 1; r.m[0] = minor_adj(0, false, RM_A, CO2, 3.0, 2.0, 0.77);
    }
    r.nx = 2; r.xslot[0] = 0; r.xwx[0] = 1; r.xslot[1] = 1; r.xwx[1] = 2;
  } else if (B == 7) {
    if (LOWER) {
      r.kind = 2; r.a = H2O; r.b = O3; r.rat = R_H2OO3; r.self = true; r.forn = true;
      r.nminor = 1; r.m[0] = minor_adj(0, true, RM_A, CO2, 3.0, 3.0, 0.79, true);
      r.planck = 1;
    } else {
      r.kind = 1; r.a = O3;
      r.nminor = 1; r.m[0] = minor_adj(1, false, RM_A, CO2, 3.0, 2.0, 0.79, true);
      r.scale = 2;
    }
  } else if (B == 8) {
    r.kind = 1; r.a = LOWER ? H2O : O3; r.self = LOWER; r.forn = LOWER;
    if (LOWER) {
      r.nminor = 3;
      r.m[0] = minor_adj(0, false, RM_A, CO2, 3.0, 2.0, 0.65);
      r.m[1] = minor_col(1, false, RM_A, O3);
      r.m[2] = minor_col(2, false, RM_A, N2O);
    } else {
      r.nminor = 2;
      r.m[0] = minor_adj(3, false, RM_A, CO2, 3.0, 2.0, 0.65);
      r.m[1] = minor_col(4, false, RM_A, N2O);
    }
    r.nx = 2; r.xslot[0] = 0; r.xwx[0] = 2; r.xslot[1] = 1; r.xwx[1] = 3;
  } else if (B == 9) {
    if (LOWER) {
      r.kind = 2; r.a = H2O; r.b = CH4; r.rat = R_H2OCH4; r.self = true; r.forn = true;
      r.nminor = 1; r.m[0] = minor_adj(0, true, RM_A, N2O, 1.5, 0.5, 0.65);
      r.planck = 1;
    } else {
      r.kind = 1; r.a = CH4;
      r.nminor = 1; r.m[0] = minor_adj(1, false, RM_A, N2O, 1.5, 0.5, 0.65);
    }
  } else if (B == 10) {
    r.kind = 1; r.a = H2O; r.self = LOWER; r.forn = true;
  } else if (B == 11) {
    r.kind = 1; r.a = H2O; r.self = LOWER; r.forn = true;
    r.nminor = 1; r.m[0] = minor_amt(LOWER ? 0 : 1, false, RM_A, A_O2SC);
  } else if (B == 12) {
    if (LOWER) { r.kind = 2; r.a = H2O; r.b = CO2; r.rat = R_H2OCO2; r.self = true; r.forn = true; r.planck = 1; }
    else r.planck = 2;
  } else if (B == 13) {
    if (LOWER) {
      r.kind = 2; r.a = H2O; r.b = N2O; r.rat = R_H2ON2O; r.self = true; r.forn = true;
      r.nminor = 2;
      r.m[0] = minor_adj(0, true, RM_A, CO2, 3.0, 2.0, 0.68, false, true);
      r.m[1] = minor_col(1, true, RM_A3, CO);
      r.planck = 1;
    } else {
      r.nminor = 1; r.m[0] = minor_col(2, false, RM_A, O3);
    }
  } else if (B == 14) {
    r.kind = 1; r.a = CO2; r.self = LOWER; r.forn = LOWER;
  } else if (B == 15) {
    if (LOWER) {
      r.kind = 2; r.a = N2O; r.b = CO2; r.rat = R_N2OCO2; r.self = true; r.forn = true;
      r.nminor = 1; r.m[0] = minor_amt(0, true, RM_A, A_BRD);
      r.planck = 1;
    } else r.planck = 2;
  } else if (B == 16) {
    if (LOWER) { r.kind = 2; r.a = H2O; r.b = CH4; r.rat = R_H2OCH4; r.self = true; r.forn = true; r.planck = 1; }
    else { r.kind = 1; r.a = CH4; }
  }
  return r;
}

constexpr int kNG[16] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
constexpr int kGS[16] = {0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138};
constexpr int kNSPA[16] = {1, 1, 9, 9, 9, 1, 9, 1, 9, 1, 1, 9, 9, 1, 9, 9};
constexpr int kNSPB[16] = {1, 1, 5, 5, 5, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0, 0};

struct BinSpec {
  double speccomb, specparm, fs;
  int js;  // 1-based like the Fortran
};
CB_HD BinSpec binspec(double cola, double rat, double colb, double n, double oneminus) {
  BinSpec s;
  s.speccomb = cola + rat * colb;
  s.specparm = cola / s.speccomb;
  if (s.specparm >= oneminus) s.specparm = oneminus;
  const double specmult = n * s.specparm;
  s.js = 1 + (int)specmult;
  s.fs = fmod(specmult, 1.0);
  return s;
}
// weights/row offsets of the 2x(2|3)-point stencil in (species ratio) x (temperature) for one pressure level
struct Stencil {
  double w[6];
  int off[6];
  int n;
};
template <bool LOWER>
CB_HD Stencil make_stencil(double specparm, double fs, double fa, double fb) {
  constexpr int nsp = LOWER ? 9 : 5;
  Stencil s;
  if (LOWER && specparm < 0.125) {
    const double p = fs - 1, p2 = p * p, p4 = p2 * p2, fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
    s.n = 6;
    s.w[0] = fk0 * fa; s.off[0] = 0;
    s.w[1] = fk1 * fa; s.off[1] = 1;
    s.w[2] = fk2 * fa; s.off[2] = 2;
    s.w[3] = fk0 * fb; s.off[3] = nsp;
    s.w[4] = fk1 * fb; s.off[4] = nsp + 1;
    s.w[5] = fk2 * fb; s.off[5] = nsp + 2;
  } else if (LOWER && specparm > 0.875) {
    const double p = -fs, p2 = p * p, p4 = p2 * p2, fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
    s.n = 6;
    s.w[0] = fk2 * fa; s.off[0] = -1;
    s.w[1] = fk1 * fa; s.off[1] = 0;
    s.w[2] = fk0 * fa; s.off[2] = 1;
    s.w[3] = fk2 * fb; s.off[3] = nsp - 1;
    s.w[4] = fk1 * fb; s.off[4] = nsp;
    s.w[5] = fk0 * fb; s.off[5] = nsp + 1;
  } else {
    s.n = 4;
    s.w[0] = (1. - fs) * fa; s.off[0] = 0;
    s.w[1] = fs * fa;        s.off[1] = 1;
    s.w[2] = (1. - fs) * fb; s.off[2] = nsp;
    s.w[3] = fs * fb;        s.off[3] = nsp + 1;
    s.w[4] = 0.; s.off[4] = 0; s.w[5] = 0.; s.off[5] = 0;
  }
  return s;
}

// Gas optical depth and Planck fraction for U consecutive g-points [g0, g0+U) of band B in ONE layer.
// gb: the (band, g-point group) block of band tables holding g0 (+ (g0 % 4)), in HBM or -- STAGED -- staged in shared memory.
template <int B, bool LOWER, int U, bool STAGED = false>
CB_HD void eval_band(const Tables& T, const double* __restrict__ gb, const double* __restrict__ ws, size_t wstride /* = nlay*ncc */,
                     int idx, double pavel, int g0, double* __restrict__ tau, double* __restrict__ frac) {
  constexpr Region R = region<B, LOWER>();
  constexpr int ng = TGW;  // row stride of the band tables (the g-points of one group)
  const BandOff& O = T.b[B - 1];
  const double* __restrict__ tb = T.base;
  auto ldtab = [](const double* __restrict__ p) { return STAGED ? ldrow_plain<U>(p) : ldrow<U>(p); };
#define WSF(f) CB_LDG(ws + (size_t)(f) * wstride)
  const int jp = idx & 63, jt = (idx >> 6) & 7, jt1 = (idx >> 9) & 7;
  const int inds = (idx >> 12) & 15, indf = (idx >> 16) & 3, indm = (idx >> 18) & 31;
  double acc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = 0.0;
  double cola = 0., colb = 0.;
  if (R.kind >= 1) cola = WSF(F_COLH2O + R.a);
  if (R.kind == 2) colb = WSF(F_COLH2O + R.b);
  CB_COV(0, B, LOWER, COV_REGION);

  // ---- key species
  if (R.kind == 1) {
    const double fac00 = WSF(F_FAC00), fac01 = WSF(F_FAC01), fac10 = WSF(F_FAC10), fac11 = WSF(F_FAC11);
    constexpr int nsp = LOWER ? kNSPA[B - 1] : kNSPB[B - 1];
    const int row0 = LOWER ? ((jp - 1) * 5 + (jt - 1)) * nsp : ((jp - 13) * 5 + (jt - 1)) * nsp;
    const int row1 = LOWER ? (jp * 5 + (jt1 - 1)) * nsp : ((jp - 12) * 5 + (jt1 - 1)) * nsp;
    const double* __restrict__ a0 = gb + (size_t)((LOWER ? O.absa : O.absb) + row0) * ng;
    const double* __restrict__ a1 = gb + (size_t)((LOWER ? O.absa : O.absb) + row1) * ng;
    const Row<U> k00 = ldtab(a0), k10 = ldtab(a0 + ng), k01 = ldtab(a1), k11 = ldtab(a1 + ng);
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = cola * (fac00 * k00[u] + fac10 * k10[u] + fac01 * k01[u] + fac11 * k11[u]);
  } else if (R.kind == 2) {
    const double fac00 = WSF(F_FAC00), fac01 = WSF(F_FAC01), fac10 = WSF(F_FAC10), fac11 = WSF(F_FAC11);
    constexpr double n = LOWER ? 8. : 4.;
    constexpr int nsp = LOWER ? 9 : 5;
    const double rat0 = CB_LDG(tb + T.rat + R.rat * 59 + jp - 1), rat1 = CB_LDG(tb + T.rat + R.rat * 59 + jp);
    const BinSpec s0 = binspec(cola, rat0, colb, n, T.oneminus);
    const BinSpec s1 = binspec(cola, rat1, colb, n, T.oneminus);
    const int row0 = (LOWER ? ((jp - 1) * 5 + (jt - 1)) * nsp : ((jp - 13) * 5 + (jt - 1)) * nsp) + s0.js - 1;
    const int row1 = (LOWER ? (jp * 5 + (jt1 - 1)) * nsp : ((jp - 12) * 5 + (jt1 - 1)) * nsp) + s1.js - 1;
    const Stencil t0 = make_stencil<LOWER>(s0.specparm, s0.fs, fac00, fac10);
    const Stencil t1 = make_stencil<LOWER>(s1.specparm, s1.fs, fac01, fac11);
    const double* __restrict__ a0 = gb + (size_t)((LOWER ? O.absa : O.absb) + row0) * ng;
    const double* __restrict__ a1 = gb + (size_t)((LOWER ? O.absa : O.absb) + row1) * ng;
    double d0[U], d1[U];
    {
      const Row<U> r0 = ldtab(a0 + t0.off[0] * ng), r1 = ldtab(a0 + t0.off[1] * ng), r2 = ldtab(a0 + t0.off[2] * ng),
                   r3 = ldtab(a0 + t0.off[3] * ng);
#pragma unroll
      for (int u = 0; u < U; ++u) d0[u] = ((t0.w[0] * r0[u] + t0.w[1] * r1[u]) + t0.w[2] * r2[u]) + t0.w[3] * r3[u];
      if (LOWER && t0.n == 6) {
        const Row<U> r4 = ldtab(a0 + t0.off[4] * ng), r5 = ldtab(a0 + t0.off[5] * ng);
#pragma unroll
        for (int u = 0; u < U; ++u) d0[u] = (d0[u] + t0.w[4] * r4[u]) + t0.w[5] * r5[u];
      }
    }
    {
      const Row<U> r0 = ldtab(a1 + t1.off[0] * ng), r1 = ldtab(a1 + t1.off[1] * ng), r2 = ldtab(a1 + t1.off[2] * ng),
                   r3 = ldtab(a1 + t1.off[3] * ng);
#pragma unroll
      for (int u = 0; u < U; ++u) d1[u] = ((t1.w[0] * r0[u] + t1.w[1] * r1[u]) + t1.w[2] * r2[u]) + t1.w[3] * r3[u];
      if (LOWER && t1.n == 6) {
        const Row<U> r4 = ldtab(a1 + t1.off[4] * ng), r5 = ldtab(a1 + t1.off[5] * ng);
#pragma unroll
        for (int u = 0; u < U; ++u) d1[u] = (d1[u] + t1.w[4] * r4[u]) + t1.w[5] * r5[u];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = s0.speccomb * d0[u] + s1.speccomb * d1[u];
    CB_COV(0, B, LOWER, !LOWER ? COV_S0_MID : (s0.specparm < 0.125 ? COV_S0_LOW : (s0.specparm > 0.875 ? COV_S0_HIGH : COV_S0_MID)));
    CB_COV(0, B, LOWER, !LOWER ? COV_S1_MID : (s1.specparm < 0.125 ? COV_S1_LOW : (s1.specparm > 0.875 ? COV_S1_HIGH : COV_S1_MID)));
  }
  if (R.kind >= 1 && acc[0] > 0.) CB_COV(0, B, LOWER, COV_KEY_NONZERO);
  // ---- water-vapour self and foreign continua
  if (R.self) {
    const double selffac = WSF(F_SELFFAC), selffrac = WSF(F_SELFFRAC);
    const double* __restrict__ s = gb + (size_t)(O.selfref + inds - 1) * ng;
    const Row<U> k0 = ldtab(s), k1 = ldtab(s + ng);
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = acc[u] + selffac * (k0[u] + selffrac * (k1[u] - k0[u]));
    if (selffac * (k0[0] + selffrac * (k1[0] - k0[0])) > 0.) CB_COV(0, B, LOWER, COV_SELF_NONZERO);
  }
  if (R.forn) {
    const double forfac = WSF(F_FORFAC), forfrac = WSF(F_FORFRAC);
    const double* __restrict__ s = gb + (size_t)(O.forref + indf - 1) * ng;
    const Row<U> k0 = ldtab(s), k1 = ldtab(s + ng);
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = acc[u] + forfac * (k0[u] + forfrac * (k1[u] - k0[u]));
    if (forfac * (k0[0] + forfrac * (k1[0] - k0[0])) > 0.) CB_COV(0, B, LOWER, COV_FOR_NONZERO);
  }
  // ---- minor gases
  if (R.nminor > 0) {
    const double minorfrac = WSF(F_MINORFRAC);
#pragma unroll
    for (int k = 0; k < R.nminor; ++k) {
      const Minor M = R.m[k];
      double amount;
      if (M.amt == A_COL) {
        amount = WSF(F_COLH2O + M.gas);
      } else if (M.amt == A_BRD_N2) {
        amount = WSF(F_COLBRD) * WSF(F_SCALEMINORN2);
      } else if (M.amt == A_BRD) {
        amount = WSF(F_COLBRD) * WSF(F_SCALEMINOR);
      } else if (M.amt == A_O2SC) {
        amount = WSF(F_COLO2) * WSF(F_SCALEMINOR);
      } else {  // A_ADJ: in atmospheres where the gas is too abundant to be "minor" (e.g. taumol.f90:529-535)
        const double colx = WSF(F_COLH2O + M.gas), coldry = WSF(F_COLDRY);
        const double chi_x = colx / coldry;
        const double e20 = M.e20f ? (double)1.e20f : 1.e20;
        const double ref = M.ref355 ? 3.55e-4 : CB_LDG(tb + T.chi_mls + M.gas * 59 + jp);  // chi_mls(gas, jp+1)
        const double ratio = e20 * chi_x / ref;
        if (ratio > M.thr) {
          const double adjfac = M.base + pow(ratio - M.base, M.expo);
          const double ref2 = M.ref355 ? (double)3.55e-4f : ref;
          amount = adjfac * ref2 * coldry * 1.e-20;
          CB_COV(0, B, LOWER, COV_MINOR0_ADJ + k);
        } else {
          amount = colx;
        }
      }
      if (amount > 0.) CB_COV(0, B, LOWER, COV_MINOR0_NONZERO + k);
      if (M.binary) {
        constexpr double n = LOWER ? 8. : 4.;
        constexpr int nsp = LOWER ? 9 : 5;
        const double refr = M.refr == RM_A ? O.refrat_m_a : (M.refr == RM_B ? O.refrat_m_b : O.refrat_m_a3);
        const BinSpec sm = binspec(cola, refr, colb, n, T.oneminus);
        const double* __restrict__ t = gb + ((size_t)O.m[M.slot] + (size_t)(sm.js - 1) + nsp * (size_t)(indm - 1)) * ng;
        const Row<U> k00 = ldtab(t), k10 = ldtab(t + ng), k01 = ldtab(t + nsp * ng), k11 = ldtab(t + (nsp + 1) * ng);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double m1 = k00[u] + sm.fs * (k10[u] - k00[u]);
          const double m2 = k01[u] + sm.fs * (k11[u] - k01[u]);
          acc[u] = acc[u] + amount * (m1 + minorfrac * (m2 - m1));
        }
      } else {
        const double* __restrict__ t = gb + (size_t)(O.m[M.slot] + indm - 1) * ng;
        const Row<U> k0 = ldtab(t), k1 = ldtab(t + ng);
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = acc[u] + amount * (k0[u] + minorfrac * (k1[u] - k0[u]));
      }
    }
  }
  // ---- halocarbon cross sections
#pragma unroll
  for (int k = 0; k < R.nx; ++k) {
    const double wx = WSF(F_WX1 + R.xwx[k]);
    if (wx > 0.) CB_COV(0, B, LOWER, COV_XSEC0_NONZERO + k);
    const double* __restrict__ t = gb + (size_t)O.x[R.xslot[k]] * ng;
    const Row<U> xr = ldtab(t);
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = acc[u] + wx * xr[u];
  }
  // ---- empirical corrections
  if (R.corr != 0) {
    double corradj;
    if (R.corr == 1) {
      corradj = 1.;
      if (pavel < 250.) corradj = 1. - 0.15 * (250. - pavel) / 154.4;
    } else if (R.corr == 2) {
      corradj = 1. - 0.15 * (pavel / 95.6);
    } else {
      corradj = 1. - .05 * (pavel - 100.) / 900.;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = corradj * acc[u];
  }
  if (R.scale == 1) {  // taumol.f90:1009-1015, default-real literals; g-points 8..14 of band 4
    const double sc[7] = {(double)0.92f, (double)0.88f, (double)1.07f, (double)1.1f, (double)0.99f, (double)0.88f, (double)0.943f};
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int g = g0 + u;
      if (g >= 7) acc[u] = acc[u] * sc[g - 7];
    }
  } else if (R.scale == 2) {  // :1642-1650; g-points 6..11 of band 7
    const double sc[6] = {0.92, 0.88, 1.07, 1.1, 0.99, 0.855};
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int g = g0 + u;
      if (g >= 5 && g <= 10) acc[u] = acc[u] * sc[g - 5];
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) tau[u] = acc[u];
  // ---- Planck fractions
  if (R.planck == 2) {
#pragma unroll
    for (int u = 0; u < U; ++u) frac[u] = 0.0;
  } else if (R.planck == 1) {
    constexpr double n = LOWER ? 8. : 4.;
    const BinSpec sp = binspec(cola, LOWER ? O.refrat_planck_a : O.refrat_planck_b, colb, n, T.oneminus);
    CB_COV(0, B, LOWER, COV_PLANCK_INTERP);
    const double* __restrict__ f = gb + (size_t)((LOWER ? O.fracrefa : O.fracrefb) + sp.js - 1) * ng;
    const Row<U> f0 = ldtab(f), f1 = ldtab(f + ng);
#pragma unroll
    for (int u = 0; u < U; ++u) frac[u] = f0[u] + sp.fs * (f1[u] - f0[u]);
  } else {
    // band 6 and 12/15 have no "b" table; band 6 upper uses fracrefa (taumol.f90:1388)
    constexpr bool use_a = LOWER || B == 6;
    const double* __restrict__ f = gb + (size_t)(use_a ? O.fracrefa : O.fracrefb) * ng;
    const Row<U> f0 = ldtab(f);
#pragma unroll
    for (int u = 0; u < U; ++u) frac[u] = f0[u];
  }
#undef WSF
}

// Planck function integrated over band B at temperature t: 1-K table, linear interpolation
// (rrtmg_lw_setcoef.f90:154-253).
CB_HD double planck_band(const double* __restrict__ tp /* totplnk row of band */, double t) {
  int ind = f2i(t - 159.);
  if (ind < 1) ind = 1; else if (ind > 180) ind = 180;
  const double fr = t - 159. - (double)ind;
  const double p0 = CB_LDG(tp + ind - 1), p1 = CB_LDG(tp + ind);
  return p0 + fr * (p1 - p0);
}

// ---------------------------------------------------------------------------------------------
// lw_unit: taumol + rtrn for U g-points of band B in one column (rrtmg_lw_rtrn.f90:320-526).
// The per-g-point work is split in two kernels (r01 ncu: 49 % of the fused kernel's stalls were memory latency with 16 warps
// per SM, 14 % instruction fetch across 32 band variants):
//   lw_taumol_unit<B,U>    band-specialised, layers independent: taug and Planck fractions of U g-points (taumol) for a
//                          chunk of layers -> two scratch rows per g-point
//   lw_transfer_unit<U,MC> ONE code body for every band: rtrn / rtrnmc (down sweep, surface, up sweep)
constexpr int NSCR = 6;               // scratch rows per g-point
constexpr int R_TAU = 4, R_FRAC = 5;  // written by lw_taumol_unit (rows 0-3: atrans, bbugas, atot, bbutot of the down sweep)

CB_HD int band_gstart(int ib) {  // first g-point of band ib (0-based), for code that is generic in the band
  switch (ib) {
    case 0: return 0; case 1: return 10; case 2: return 22; case 3: return 38; case 4: return 52; case 5: return 68;
    case 6: return 76; case 7: return 88; case 8: return 96; case 9: return 108; case 10: return 114; case 11: return 122;
    case 12: return 130; case 13: return 134; case 14: return 136; default: return 138;
  }
}

// staged: the (band, group) block of band tables already copied to shared memory (CUDA kernel only), else null -> read from HBM
template <int B, int U, bool STAGED = false>
CB_HD void lw_taumol_unit(const Tables& T, const In& in, const Work& W, int c0, int c, int g0, int l0, int l1,
                          const double* __restrict__ staged = nullptr) {
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const size_t gc = (size_t)(c0 + c);
  const size_t wstride = (size_t)nlay * ncc;
  const int laytrop = W.laytrop[c];
  const int gabs = kGS[B - 1] + g0;
  const BandOff& O = T.b[B - 1];
  const double* __restrict__ gb = (STAGED ? staged : T.base + O.base + (size_t)(g0 / TGW) * O.rows * TGW) + (g0 % TGW);
  for (int l = l0; l < l1; ++l) {
    const int lev = l + 1;
    const int idx = W.idx[(size_t)l * ncc + c];
    const double pavel = in.play[(size_t)l * ncol + gc];
    double tau[U], frac[U];
    const double* ws = W.ws + (size_t)l * ncc + c;
    if (lev <= laytrop) eval_band<B, true, U, STAGED>(T, gb, ws, wstride, idx, pavel, g0, tau, frac);
    else eval_band<B, false, U, STAGED>(T, gb, ws, wstride, idx, pavel, g0, tau, frac);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double* __restrict__ scr = W.scr + (((size_t)(gabs + u) * NSCR) * nlay + l) * ncc + c;
      scr[R_TAU * wstride] = tau[u];
      scr[R_FRAC * wstride] = frac[u];
    }
  }
}

// rtrn / rtrnmc for U consecutive g-points of band ib (0-based) -- generic in the band (rrtmg_lw_rtrn.f90:300-557)
// MR: maximum-random overlap of the fractional clouds (rtrnmr, rrtmg_lw_rtrnmr.f90:481-700) instead of rtrn's random overlap.
// Where the two sweeps put the radiance sums of their g-points, one call per interface (down sweep: from the top down, up sweep:
// from the surface up), already weighted by the band's wtdiff * delwave (rtrn.f90:529-543) so that units of different bands can
// share rows.  The units of a GROUP of CB_LW_GROUP units share one set of rows `part[group][npart][nlay+1][ncc]`:
//   * CUDA kernel (lw_engine.cu, LwPartSmem): the units of a group are the warps of one block; values are staged in shared memory a
//     few interfaces at a time and summed over the warps in unit order before they reach HBM (r01: the per-unit rows and the
//     kernel that re-read them were ~19 % of a step's DRAM traffic);
//   * host emulation (LwPartDirect): the units of a group run one after the other and accumulate into the zeroed rows in the same
//     order -- the same bits.
// Rows: 0 up, 1 down, 2 up clear, 3 down clear [, 4 d(up)/dTs, 5 d(up clear)/dTs].  The clear-sky rows of a cloud-free column are
// not stored (equal to the total ones bit for bit); the downward flux at the top interface is zero and not stored either.
#ifndef CB_LW_GROUP
#define CB_LW_GROUP 4  // 70 units of 2 g-points -> 18 groups (the last one half empty): 128-thread blocks, 6 per SM at 80 registers
#endif
struct LwPartDirect {
  double* part;  // rows of this unit's group, at this column
  size_t pstride;
  int ncc;
  CB_HD void put_dn(size_t lev, bool cloudy_col, double dn, double dnc) {
    part[1 * pstride + lev * ncc] += dn;
    if (cloudy_col) part[3 * pstride + lev * ncc] += dnc;
  }
  CB_HD void put_up(size_t lev, bool cloudy_col, double up, double upc, bool drv, double dup, double dupc) {
    part[0 * pstride + lev * ncc] += up;
    if (cloudy_col) part[2 * pstride + lev * ncc] += upc;
    if (drv) {
      part[4 * pstride + lev * ncc] += dup;
      if (cloudy_col) part[5 * pstride + lev * ncc] += dupc;
    }
  }
  CB_HD void end_sweep() {}
};

// DRV: also the derivative of the upward flux with respect to the surface temperature (idrv = 1, rtrn.f90:458-461,473-476,
// 492-512; the same lines in rtrnmc.f90:447-510 and rtrnmr.f90:629-711): one more multiplicative recurrence in the up sweep.
// Where the down sweep parks what the up sweep needs again (rows 0-3 of NSCR).  Default (p == nullptr): the unit's own rows of
// W.scr.  The slab form of the CUDA kernel (lw_engine.cu, k_units_slab) points it at a per-warp slab reused for every unit the
// warp processes: written top-down, read back bottom-up (last in, first out), so a slab that fits the warp's share of the L2
// never reaches HBM.
struct Carry {
  double* p;          // at this thread's lane
  size_t rs, ls, us;  // strides between rows, layers and the g-points of a unit
};
// SLAB: the carried rows live in an L2-resident slab (cy given): cache-residency hints of cb_common.h on every access, the taumol
// rows fetched kSlabAhead layers ahead into the L2.
constexpr int kSlabAhead = 4;
template <int U, bool MC, bool MR, bool DRV, class Sink, bool SLAB = false>
CB_HD void lw_transfer_unit(const Tables& T, const In& in, const Work& W, int c0, int c, int ib, int g0, Sink& sink,
                            Carry cy = Carry{nullptr, 0, 0, 0}) {
  const unsigned long long pol_keep = SLAB ? policy_keep() : 0ull, pol_drop = SLAB ? policy_drop() : 0ull;
  auto cst = [&](double* p, double v) { if (SLAB) st_policy(p, v, pol_keep); else *p = v; };
  auto cld = [&](const double* p) { return SLAB ? ld_policy(p, pol_drop) : *p; };
  auto sld = [&](const double* p) { return SLAB ? ld_stream(p) : *p; };
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const size_t gc = (size_t)(c0 + c);
  const double* __restrict__ tb = T.base;
  const double rec_6 = 0.166667, tblint = 10000.0, bpade = T.bpade;
  const double* __restrict__ tau_tbl = tb + T.tau_tbl;
  const double* __restrict__ exp_tbl = tb + T.exp_tbl;
  const double* __restrict__ tfn_tbl = tb + T.tfn_tbl;
  const double* __restrict__ et_tbl = tb + T.et_tbl;
  (void)exp_tbl; (void)tfn_tbl;
  const double* __restrict__ tp = tb + T.totplnk + (size_t)ib * 181;
  const size_t wstride = (size_t)nlay * ncc;
  const int ncb = W.ncbands[c];
  const double pwvcm = W.pwvcm[c];
  const double secdiff = secdiff_band(pwvcm, ib);
  // cloud band feeding this LW band: ipat(B, 0|1|2) for ncbands = 1|5|16 (rtrn.f90:233-235,324-330)
  int ibc = 0;
  if (ncb == 5) {
    ibc = ib <= 1 ? ib : (ib <= 4 ? 2 : (ib <= 7 ? 3 : 4));  // ipat(:,1) = 1,2,3,3,3,4,4,4,5,...,5
  } else if (ncb == 16) {
    ibc = ib;
  }
  const int gabs = band_gstart(ib) + g0;  // absolute g-point of u = 0
  if (!cy.p) cy = Carry{W.scr + ((size_t)gabs * NSCR) * wstride + c, wstride, (size_t)ncc, (size_t)NSCR * wstride};
  double radld[U], radclrd[U], frac1[U];
  double mr_cld[U], mr_clr[U], mr_rad[U];  // MR: cloudy / clear parts of the radiance and the overlap correction (cldradd, clrradd, rad)
#pragma unroll
  for (int u = 0; u < U; ++u) { mr_cld[u] = 0.; mr_clr[u] = 0.; mr_rad[u] = 0.; }
  const size_t ovs = (size_t)(nlay + 2) * ncc;  // row stride of W.ovl
#pragma unroll
  for (int u = 0; u < U; ++u) { radld[u] = 0.; radclrd[u] = 0.; frac1[u] = 0.; }
  int iclddn = 0;
  const double wband = 0.5 * CB_LDG(tb + T.delwave + ib);  // wtdiff * delwave(iband)
  const bool cloudy_col = ncb > 0;
  double plev_up = planck_band(tp, in.tlev[(size_t)nlay * ncol + gc]);  // planklev(nlay)
  // The loads of a layer do not depend on the recurrence: fetch layer lev-1 while layer lev is computed (the kernel's top
  // stall is memory latency at 24 warps per SM).
  struct LayerIn {
    double taua, tlay, tlev, tau[U], frac[U];
  };
  auto fetch = [&](int l) {
    LayerIn v;
    const size_t o = (size_t)l * ncol + gc;
    v.taua = in.tauaer[((size_t)ib * nlay + l) * ncol + gc];
    v.tlay = in.tlay[o];
    v.tlev = in.tlev[o];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* __restrict__ scr = W.scr + (((size_t)(gabs + u) * NSCR) * nlay + l) * ncc + c;
      if (SLAB && l >= kSlabAhead) {
        prefetch_l2(scr + R_TAU * wstride - (size_t)kSlabAhead * ncc);
        prefetch_l2(scr + R_FRAC * wstride - (size_t)kSlabAhead * ncc);
      }
      v.tau[u] = sld(scr + R_TAU * wstride);
      v.frac[u] = sld(scr + R_FRAC * wstride);
    }
    return v;
  };
  LayerIn nxt = fetch(nlay - 1);
  for (int lev = nlay; lev >= 1; --lev) {
    const int l = lev - 1;
    const size_t o = (size_t)l * ncol + gc;
    const LayerIn cur = nxt;
    if (lev > 1) nxt = fetch(l - 1);
    const double taua = cur.taua;
    const double blay = planck_band(tp, cur.tlay);
    const double plev_dn = planck_band(tp, cur.tlev);  // planklev(lev-1)
    const double dplankup = plev_up - blay;
    const double dplankdn = plev_dn - blay;
    plev_up = plev_dn;
    // non-McICA: a layer is cloudy when its cloud fraction is >= 1e-6 (rtrn.f90:302); McICA: when ANY sub-column
    // of the layer is cloudy (rtrnmc.f90:298-309), each g-point then sees cloud fraction 0 or 1.
    bool cloudy;
    unsigned mbits = 0u;
    double odcld_l = 0., efclfrac_l = 0., cldfrac_l = 0.;
    if (MC) {
      cloudy = false;
      if (ncb > 0) {
        const size_t ms = (size_t)W.mstride;
        const unsigned* mw = W.mask + ((size_t)l * 5) * ms + W.moff + c;
        const unsigned w0 = mw[0], w1 = mw[ms], w2 = mw[2 * ms], w3 = mw[3 * ms], w4 = mw[4 * ms];
        cloudy = (w0 | w1 | w2 | w3 | w4) != 0u;
        // the unit's <= 4 g-points never straddle a 32-bit word boundary twice; fetch their bits
        for (int u = 0; u < U; ++u) {
          const int g = gabs + u;
          const unsigned w = (g >> 5) == 0 ? w0 : ((g >> 5) == 1 ? w1 : ((g >> 5) == 2 ? w2 : ((g >> 5) == 3 ? w3 : w4)));
          mbits |= ((w >> (g & 31)) & 1u) << u;
        }
      }
    } else {
      cloudy = ncb > 0 && in.cldfr[o] >= 1.e-6;
    }
    if (cloudy) {
      iclddn = 1;
      cldfrac_l = MC ? 1.0 : in.cldfr[o];
      odcld_l = W.cld[((size_t)ibc * nlay + l) * ncc + c];
      efclfrac_l = W.cld[((size_t)(16 + ibc) * nlay + l) * ncc + c];
    }
    double ov_ist = 0., ov_cld1 = 0., ov_cld2 = 0., ov_clr1 = 0., ov_clr2 = 0., ov_cmb1 = 0., ov_cmb2 = 0.;
    if (MR && cloudy) {  // factors of the downward path at index lev-1, istcldd(lev)
      const double* ov = W.ovl + (size_t)(lev - 1) * ncc + c;
      ov_ist = W.ovl[OV_ISTD * ovs + (size_t)lev * ncc + c];
      ov_cld1 = ov[OV_CLD1D * ovs]; ov_cld2 = ov[OV_CLD2D * ovs]; ov_clr1 = ov[OV_CLR1D * ovs]; ov_clr2 = ov[OV_CLR2D * ovs];
      ov_cmb1 = ov[OV_CMB1D * ovs]; ov_cmb2 = ov[OV_CMB2D * ovs];
    }
    double sum_d = 0., sum_dc = 0.;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool on = !MC || ((mbits >> u) & 1u);
      const double odcld = on ? odcld_l : 0., efclfrac = on ? efclfrac_l : 0., cldfrac = on ? cldfrac_l : 0.;
      double* __restrict__ cr = cy.p + (size_t)u * cy.us + (size_t)l * cy.ls;
      const double plfrac = cur.frac[u];
      double odepth = secdiff * (cur.tau[u] + taua);
      if (odepth < 0.0) odepth = 0.0;
      double atrans, bbd, bbugas;
      if (cloudy) {
        double odtot = odepth + odcld;
        double gassrc, bbdtot, atot, bbutot;
        if (odtot < 0.06) {
          atrans = odepth - 0.5 * odepth * odepth;
          const double odepth_rec = rec_6 * odepth;
          gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans;
          atot = odtot - 0.5 * odtot * odtot;
          const double odtot_rec = rec_6 * odtot;
          bbdtot = plfrac * (blay + dplankdn * odtot_rec);
          bbd = plfrac * (blay + dplankdn * odepth_rec);
          bbugas = plfrac * (blay + dplankup * odepth_rec);
          bbutot = plfrac * (blay + dplankup * odtot_rec);
        } else if (odepth <= 0.06) {
          atrans = odepth - 0.5 * odepth * odepth;
          const double odepth_rec = rec_6 * odepth;
          gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans;
          odtot = odepth + odcld;
          const int ittot = tbl_slot(odtot, bpade, tblint);
          const Row<2> ett = ldrow<2>(et_tbl + 2 * ittot);
          const double tfactot = ett[1];
          bbdtot = plfrac * (blay + tfactot * dplankdn);
          bbd = plfrac * (blay + dplankdn * odepth_rec);
          atot = 1. - ett[0];
          bbugas = plfrac * (blay + dplankup * odepth_rec);
          bbutot = plfrac * (blay + tfactot * dplankup);
        } else {
          const int itgas = tbl_slot(odepth, bpade, tblint);
          odepth = CB_LDG(tau_tbl + itgas);
          const Row<2> etg = ldrow<2>(et_tbl + 2 * itgas);
          atrans = 1. - etg[0];
          const double tfacgas = etg[1];
          gassrc = atrans * plfrac * (blay + tfacgas * dplankdn);
          odtot = odepth + odcld;
          const int ittot = tbl_slot(odtot, bpade, tblint);
          const Row<2> ett = ldrow<2>(et_tbl + 2 * ittot);
          const double tfactot = ett[1];
          bbdtot = plfrac * (blay + tfactot * dplankdn);
          bbd = plfrac * (blay + tfacgas * dplankdn);
          atot = 1. - ett[0];
          bbugas = plfrac * (blay + tfacgas * dplankup);
          bbutot = plfrac * (blay + tfactot * dplankup);
        }
        if (MR) {
          // rtrnmr.f90:591-616
          if (ov_ist == 1.) {
            mr_cld[u] = cldfrac * radld[u];
            mr_clr[u] = radld[u] - mr_cld[u];
            mr_rad[u] = 0.;
          }
          const double ttot = 1. - atot;
          const double cldsrc = bbdtot * atot;
          mr_cld[u] = mr_cld[u] * ttot + cldfrac * cldsrc;
          mr_clr[u] = mr_clr[u] * (1. - atrans) + (1. - cldfrac) * gassrc;
          radld[u] = mr_cld[u] + mr_clr[u];
          const double radmod = mr_rad[u] * (ov_clr1 * (1. - atrans) + ov_cld1 * ttot) - ov_cmb1 * gassrc + ov_cmb2 * cldsrc;
          const double oldcld = mr_cld[u] - radmod;
          const double oldclr = mr_clr[u] + radmod;
          mr_rad[u] = -radmod + ov_clr2 * oldclr - ov_cld2 * oldcld;
          mr_cld[u] = mr_cld[u] + mr_rad[u];
          mr_clr[u] = mr_clr[u] - mr_rad[u];
        } else {
          radld[u] = radld[u] - radld[u] * (atrans + efclfrac * (1. - atrans)) + gassrc + cldfrac * (bbdtot * atot - gassrc);
        }
        cst(cr + 2 * cy.rs, atot);
        cst(cr + 3 * cy.rs, bbutot);
      } else {
        if (odepth <= 0.06) {
          atrans = odepth - 0.5 * odepth * odepth;
          odepth = rec_6 * odepth;
          bbd = plfrac * (blay + dplankdn * odepth);
          bbugas = plfrac * (blay + dplankup * odepth);
        } else {
          const int itr = tbl_slot(odepth, bpade, tblint);
          const Row<2> et = ldrow<2>(et_tbl + 2 * itr);
          const double transc = et[0];
          atrans = 1. - transc;
          const double tausfac = et[1];
          bbd = plfrac * (blay + tausfac * dplankdn);
          bbugas = plfrac * (blay + tausfac * dplankup);
        }
        radld[u] = radld[u] + (bbd - radld[u]) * atrans;
      }
      cst(cr, atrans);
      cst(cr + cy.rs, bbugas);
      sum_d = sum_d + radld[u];
      if (iclddn == 1) {
        radclrd[u] = radclrd[u] + (bbd - radclrd[u]) * atrans;
      } else {
        radclrd[u] = radld[u];
      }
      sum_dc = sum_dc + radclrd[u];
      if (lev == 1) frac1[u] = plfrac;
    }
    sink.put_dn((size_t)(lev - 1), cloudy_col, sum_d * wband, sum_dc * wband);
  }
  sink.end_sweep();  // (top of atmosphere: no downward flux -- lw_reduce_level knows)
  // surface (rtrn.f90:455-470)
  const double tbound = in.tsfc[gc];
  const double semiss = in.emis[(size_t)ib * ncol + gc];
  const double plankbnd = semiss * planck_band(tp, tbound);
  const double reflect = 1. - semiss;
  double radlu[U], radclru[U];
  double d_radlu[U], d_radclru[U];  // DRV
  {
    double s = 0., sc = 0., sd = 0.;
    // dplankbnd_dt (setcoef.f90:197-200): the same 1-K table interpolation on totplnkderiv
    const double dplankbnd_dt = DRV ? semiss * planck_band(tb + T.totplnkderiv + (size_t)ib * 181, tbound) : 0.;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double rad0 = frac1[u] * plankbnd;
      radlu[u] = rad0 + reflect * radld[u];
      radclru[u] = rad0 + reflect * radclrd[u];
      s = s + radlu[u];
      sc = sc + radclru[u];
      d_radlu[u] = frac1[u] * dplankbnd_dt;
      d_radclru[u] = d_radlu[u];
      sd = sd + d_radlu[u];
    }
    sink.put_up(0, cloudy_col, s * wband, sc * wband, DRV, sd * wband, sd * wband);
  }
  // upward sweep (rtrn.f90:478-521), the (atrans, bbugas) rows of the next layer fetched one layer ahead
  struct UpIn {
    double atrans[U], bbugas[U];
  };
  auto fetch_up = [&](int l) {
    UpIn v;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* __restrict__ cr = cy.p + (size_t)u * cy.us + (size_t)l * cy.ls;
      v.atrans[u] = cld(cr);
      v.bbugas[u] = cld(cr + cy.rs);
    }
    return v;
  };
  UpIn unx = fetch_up(0);
  for (int lev = 1; lev <= nlay; ++lev) {
    const int l = lev - 1;
    const size_t o = (size_t)l * ncol + gc;
    const UpIn ucur = unx;
    if (lev < nlay) unx = fetch_up(l + 1);
    bool cloudy;
    unsigned mbits = 0u;
    if (MC) {
      cloudy = false;
      if (ncb > 0) {
        const size_t ms = (size_t)W.mstride;
        const unsigned* mw = W.mask + ((size_t)l * 5) * ms + W.moff + c;
        const unsigned w0 = mw[0], w1 = mw[ms], w2 = mw[2 * ms], w3 = mw[3 * ms], w4 = mw[4 * ms];
        cloudy = (w0 | w1 | w2 | w3 | w4) != 0u;
        for (int u = 0; u < U; ++u) {
          const int g = gabs + u;
          const unsigned w = (g >> 5) == 0 ? w0 : ((g >> 5) == 1 ? w1 : ((g >> 5) == 2 ? w2 : ((g >> 5) == 3 ? w3 : w4)));
          mbits |= ((w >> (g & 31)) & 1u) << u;
        }
      }
    } else {
      cloudy = ncb > 0 && in.cldfr[o] >= 1.e-6;
    }
    double efclfrac_l = 0., cldfrac_l = 0.;
    if (cloudy) {
      cldfrac_l = MC ? 1.0 : in.cldfr[o];
      efclfrac_l = W.cld[((size_t)(16 + ibc) * nlay + l) * ncc + c];
    }
    double ov_ist = 0., ov_cld1 = 0., ov_cld2 = 0., ov_clr1 = 0., ov_clr2 = 0., ov_cmb1 = 0., ov_cmb2 = 0.;
    if (MR && cloudy) {  // factors of the upward path at index lev+1, istcld(lev)
      const double* ov = W.ovl + (size_t)(lev + 1) * ncc + c;
      ov_ist = W.ovl[OV_IST * ovs + (size_t)lev * ncc + c];
      ov_cld1 = ov[OV_CLD1 * ovs]; ov_cld2 = ov[OV_CLD2 * ovs]; ov_clr1 = ov[OV_CLR1 * ovs]; ov_clr2 = ov[OV_CLR2 * ovs];
      ov_cmb1 = ov[OV_CMB1 * ovs]; ov_cmb2 = ov[OV_CMB2 * ovs];
    }
    double s = 0., sc = 0., sd = 0., sdc = 0.;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool on = !MC || ((mbits >> u) & 1u);
      const double efclfrac = on ? efclfrac_l : 0., cldfrac = on ? cldfrac_l : 0.;
      const double* __restrict__ cr = cy.p + (size_t)u * cy.us + (size_t)l * cy.ls;
      const double atrans = ucur.atrans[u], bbugas = ucur.bbugas[u];
      if (cloudy) {
        const double atot = cld(cr + 2 * cy.rs), bbutot = cld(cr + 3 * cy.rs);
        const double gassrc = bbugas * atrans;
        if (DRV) d_radlu[u] = d_radlu[u] * cldfrac * (1.0 - atot) + d_radlu[u] * (1.0 - cldfrac) * (1.0 - atrans);
        if (MR) {
          // rtrnmr.f90:655-677
          if (ov_ist == 1.) {
            mr_cld[u] = cldfrac * radlu[u];
            mr_clr[u] = radlu[u] - mr_cld[u];
            mr_rad[u] = 0.;
          }
          const double ttot = 1. - atot;
          const double cldsrc = bbutot * atot;
          mr_cld[u] = mr_cld[u] * ttot + cldfrac * cldsrc;
          mr_clr[u] = mr_clr[u] * (1.0 - atrans) + (1. - cldfrac) * gassrc;
          radlu[u] = mr_cld[u] + mr_clr[u];
          const double radmod = mr_rad[u] * (ov_clr1 * (1.0 - atrans) + ov_cld1 * ttot) - ov_cmb1 * gassrc + ov_cmb2 * cldsrc;
          const double oldcld = mr_cld[u] - radmod;
          const double oldclr = mr_clr[u] + radmod;
          mr_rad[u] = -radmod + ov_clr2 * oldclr - ov_cld2 * oldcld;
          mr_cld[u] = mr_cld[u] + mr_rad[u];
          mr_clr[u] = mr_clr[u] - mr_rad[u];
        } else {
          radlu[u] = radlu[u] - radlu[u] * (atrans + efclfrac * (1. - atrans)) + gassrc + cldfrac * (bbutot * atot - gassrc);
        }
      } else {
        radlu[u] = radlu[u] + (bbugas - radlu[u]) * atrans;
        if (DRV) d_radlu[u] = d_radlu[u] * (1.0 - atrans);
      }
      s = s + radlu[u];
      if (iclddn == 1) {
        radclru[u] = radclru[u] + (bbugas - radclru[u]) * atrans;
        if (DRV) d_radclru[u] = d_radclru[u] * (1.0 - atrans);
      } else {
        radclru[u] = radlu[u];
        if (DRV) d_radclru[u] = d_radlu[u];
      }
      sc = sc + radclru[u];
      if (DRV) { sd = sd + d_radlu[u]; sdc = sdc + d_radclru[u]; }
    }
    sink.put_up((size_t)lev, cloudy_col, s * wband, sc * wband, DRV, sd * wband, sdc * wband);
  }
  sink.end_sweep();
}

// ---------------------------------------------------------------------------------------------
// Column-tile form of the transfer (lw_engine.cu: k_lw_tile; host emulation: tests/emul/lw_emul.cpp).
// The recurrence over the levels is cheap and strictly serial; what costs time in lw_transfer_unit is everything around it, which
// does NOT depend on the recurrence: the optical-depth -> transmittance / Planck source conversion of every (layer, column) cell.
// The tile form separates the two:
//   lw_tile_cell    one (layer, column) cell of one g-point: the layer's rows -- any thread of the block can evaluate any cell, so
//                   the cells of a tile of TW adjacent columns are spread over all producer warps, their rows parked in SHARED memory;
//   lw_tile_sweeps  the down sweep, surface and up sweep of one radiance stream of one column over those rows, adding the
//                   band-weighted radiances of the g-point to per-level sums that also live in shared memory.
// Nothing the two sweeps exchange reaches HBM; the sums over the g-points of a band group are written once per tile.
// Same physics as lw_transfer_unit<U, MC, MR = false, DRV = false> with the per-level update brought to ONE dependent operation:
//   clear layer   rad + (bb - rad) atrans                                  = rad (1 - atrans) + bb atrans       (rtrn.f90:408, 517)
//   cloudy layer  rad - rad (atrans + efclfrac (1 - atrans)) + gassrc + cldfrac (bbtot atot - gassrc)            (rtrn.f90:386-393)
//                                                                          = rad (1 - A) + S
// i.e. rad <- fma(rad, T, S) with the layer's T and S formed per cell (a convex combination either way; the rounding differs from
// the reference's association in the last bits, tests: <= 1e-12 of the unit form).
// Rows of a cell: the clear-sky stream's T, S_down, S_up and -- cloudy form -- the total-sky stream's (equal to them in a clear layer).
constexpr int TR_T = 0, TR_SD = 1, TR_SU = 2, TR_TT = 3, TR_SDT = 4, TR_SUT = 5;
constexpr int kTileRowsClear = 3, kTileRowsCloudy = 6;

// secdiff and cloud band of band ib for one column: what lw_transfer_unit derives in its prologue
struct TileBandCol {
  int ibc, ncb;
  double secdiff;
};
CB_HD TileBandCol lw_tile_band_col(const Work& W, int c, int ib) {
  TileBandCol b;
  b.ncb = W.ncbands[c];
  b.secdiff = secdiff_band(W.pwvcm[c], ib);
  b.ibc = 0;
  if (b.ncb == 5) b.ibc = ib <= 1 ? ib : (ib <= 4 ? 2 : (ib <= 7 ? 3 : 4));
  else if (b.ncb == 16) b.ibc = ib;
  return b;
}
// what a cell needs that does not depend on the band: is the layer cloudy, and for which g-points (McICA)
struct TileCellCol {
  bool cloudy;         // non-McICA: cloud fraction >= 1e-6 (rtrn.f90:302); McICA: ANY sub-column of the layer cloudy (rtrnmc.f90:298-309)
  double cldfrac;      // non-McICA cloud fraction
  unsigned mask[5];    // McICA: the layer's sub-column bits
};
template <bool MC>
CB_HD TileCellCol lw_tile_cell_col(const In& in, const Work& W, int c0, int c, int l) {
  TileCellCol k;
  k.cloudy = false; k.cldfrac = 0.;
  for (int i = 0; i < 5; ++i) k.mask[i] = 0u;
  if (W.ncbands[c] > 0) {
    if (MC) {
      const size_t ms = (size_t)W.mstride;
      const unsigned* mw = W.mask + ((size_t)l * 5) * ms + W.moff + c;
      unsigned any = 0u;
      for (int i = 0; i < 5; ++i) { k.mask[i] = mw[(size_t)i * ms]; any |= k.mask[i]; }
      k.cloudy = any != 0u;
      k.cldfrac = 1.0;
    } else {
      k.cldfrac = in.cldfr[(size_t)l * in.ncol + (size_t)(c0 + c)];
      k.cloudy = k.cldfrac >= 1.e-6;
    }
  }
  return k;
}
// what a cell needs that is the same for every g-point of its band: aerosol optical depth, the Planck terms of the layer, the
// band's cloud optical depth and effective cloud fraction factor
struct TileCellBand {
  double taua, blay, dplankdn, dplankup, odcld, efclfrac;
};
template <bool CLOUDY>
CB_HD TileCellBand lw_tile_cell_band(const Tables& T, const In& in, const Work& W, int c0, int c, int l, int ib, const TileBandCol& bc,
                                     bool cloudy) {
  const int nlay = in.nlay, ncol = in.ncol;
  const size_t gc = (size_t)(c0 + c);
  const double* __restrict__ tp = T.base + T.totplnk + (size_t)ib * 181;
  const size_t o = (size_t)l * ncol + gc;
  TileCellBand b;
  b.taua = in.tauaer[((size_t)ib * nlay + l) * ncol + gc];
  b.blay = planck_band(tp, in.tlay[o]);
  b.dplankdn = planck_band(tp, in.tlev[o]) - b.blay;         // planklev(lev-1) - planklay(lev)
  b.dplankup = planck_band(tp, in.tlev[o + ncol]) - b.blay;  // planklev(lev)   - planklay(lev)
  b.odcld = 0.; b.efclfrac = 0.;
  if (CLOUDY && cloudy) {
    b.odcld = W.cld[((size_t)bc.ibc * nlay + l) * W.ncc + c];
    b.efclfrac = W.cld[((size_t)(16 + bc.ibc) * nlay + l) * W.ncc + c];
  }
  return b;
}
// the two taumol rows of cell (l, c) for g-point gabs: optical depth and Planck fraction (streamed: read once)
CB_HD void lw_tile_cell_load(const In& in, const Work& W, int c, int l, int gabs, double& tau, double& plfrac) {
  const size_t wstride = (size_t)in.nlay * W.ncc;
  const double* __restrict__ scr = W.scr + (((size_t)gabs * NSCR) * in.nlay + l) * W.ncc + c;
  tau = ld_stream(scr + R_TAU * wstride);
  plfrac = ld_stream(scr + R_FRAC * wstride);
}

// rows of one cell for g-point gabs -> out[r * rs], r = TR_*
template <bool MC, bool CLOUDY>
CB_HD void lw_tile_cell(const Tables& T, int gabs, const TileBandCol& bc, const TileCellCol& cc, const TileCellBand& cb_, double tau,
                        double plfrac, double* __restrict__ out, size_t rs) {
  const double* __restrict__ tb = T.base;
  const double rec_6 = 0.166667, tblint = 10000.0, bpade = T.bpade;
  const double* __restrict__ tau_tbl = tb + T.tau_tbl;
  const double* __restrict__ et_tbl = tb + T.et_tbl;
  const double taua = cb_.taua, blay = cb_.blay, dplankdn = cb_.dplankdn, dplankup = cb_.dplankup;
  double odepth = bc.secdiff * (tau + taua);
  if (odepth < 0.0) odepth = 0.0;
  double atrans, bbd, bbugas;
  if (CLOUDY && cc.cloudy) {
    bool on = true;
    if (MC) {  // this g-point's sub-column: cloud fraction 0 or 1
      const int wi = gabs >> 5;
      const unsigned w = wi == 0 ? cc.mask[0] : (wi == 1 ? cc.mask[1] : (wi == 2 ? cc.mask[2] : (wi == 3 ? cc.mask[3] : cc.mask[4])));
      on = ((w >> (gabs & 31)) & 1u) != 0u;
    }
    const double odcld = on ? cb_.odcld : 0., efclfrac = on ? cb_.efclfrac : 0., cldfrac = on ? cc.cldfrac : 0.;
    double odtot = odepth + odcld;
    double gassrc, bbdtot, atot, bbutot;
    if (odtot < 0.06) {
      atrans = odepth - 0.5 * odepth * odepth;
      const double odepth_rec = rec_6 * odepth;
      gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans;
      atot = odtot - 0.5 * odtot * odtot;
      const double odtot_rec = rec_6 * odtot;
      bbdtot = plfrac * (blay + dplankdn * odtot_rec);
      bbd = plfrac * (blay + dplankdn * odepth_rec);
      bbugas = plfrac * (blay + dplankup * odepth_rec);
      bbutot = plfrac * (blay + dplankup * odtot_rec);
    } else if (odepth <= 0.06) {
      atrans = odepth - 0.5 * odepth * odepth;
      const double odepth_rec = rec_6 * odepth;
      gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans;
      odtot = odepth + odcld;
      const int ittot = tbl_slot(odtot, bpade, tblint);
      const Row<2> ett = ldrow<2>(et_tbl + 2 * ittot);
      const double tfactot = ett[1];
      bbdtot = plfrac * (blay + tfactot * dplankdn);
      bbd = plfrac * (blay + dplankdn * odepth_rec);
      atot = 1. - ett[0];
      bbugas = plfrac * (blay + dplankup * odepth_rec);
      bbutot = plfrac * (blay + tfactot * dplankup);
    } else {
      const int itgas = tbl_slot(odepth, bpade, tblint);
      odepth = CB_LDG(tau_tbl + itgas);
      const Row<2> etg = ldrow<2>(et_tbl + 2 * itgas);
      atrans = 1. - etg[0];
      const double tfacgas = etg[1];
      gassrc = atrans * plfrac * (blay + tfacgas * dplankdn);
      odtot = odepth + odcld;
      const int ittot = tbl_slot(odtot, bpade, tblint);
      const Row<2> ett = ldrow<2>(et_tbl + 2 * ittot);
      const double tfactot = ett[1];
      bbdtot = plfrac * (blay + tfactot * dplankdn);
      bbd = plfrac * (blay + tfacgas * dplankdn);
      atot = 1. - ett[0];
      bbugas = plfrac * (blay + tfacgas * dplankup);
      bbutot = plfrac * (blay + tfactot * dplankup);
    }
    // rtrn.f90:386-393 / 503-508 (rtrnmc.f90:377-384 / 490-495), with the upward gas source bbugas * atrans
    const double gassrc_up = bbugas * atrans;
    out[TR_TT * rs] = 1. - (atrans + efclfrac * (1. - atrans));
    out[TR_SDT * rs] = gassrc + cldfrac * (bbdtot * atot - gassrc);
    out[TR_SUT * rs] = gassrc_up + cldfrac * (bbutot * atot - gassrc_up);
    out[TR_T * rs] = 1. - atrans;
    out[TR_SD * rs] = bbd * atrans;
    out[TR_SU * rs] = gassrc_up;
  } else {
    if (odepth <= 0.06) {
      atrans = odepth - 0.5 * odepth * odepth;
      odepth = rec_6 * odepth;
      bbd = plfrac * (blay + dplankdn * odepth);
      bbugas = plfrac * (blay + dplankup * odepth);
    } else {
      const int itr = tbl_slot(odepth, bpade, tblint);
      const Row<2> et = ldrow<2>(et_tbl + 2 * itr);
      atrans = 1. - et[0];
      const double tausfac = et[1];
      bbd = plfrac * (blay + tausfac * dplankdn);
      bbugas = plfrac * (blay + tausfac * dplankup);
    }
    const double t = 1. - atrans, sd = bbd * atrans, su = bbugas * atrans;
    out[TR_T * rs] = t; out[TR_SD * rs] = sd; out[TR_SU * rs] = su;
    if (CLOUDY) { out[TR_TT * rs] = t; out[TR_SDT * rs] = sd; out[TR_SUT * rs] = su; }
  }
}

// The three sweeps of ONE radiance stream of one column over the parked rows of one g-point (rtrn.f90:320-526): P = the stream's
// three rows of this column (T, S_down, S_up; row stride rs, layer stride ls); acc_dn / acc_up = the stream's per-level sums of
// this column (level stride als).  The clear-sky stream of a column with clouds reads the clear rows from the top (lw_transfer_unit
// copies the total-sky radiance above the highest cloud: the same update of the same inputs, the same bits).  The levels go in
// chunks of four whose rows and sums are loaded before the chain -- one fma per level -- runs over them.
// rad0 = plfrac(layer 1) * emissivity * B(T_sfc).
CB_HD void lw_tile_sweeps(const double* __restrict__ P, size_t rs, size_t ls, int nlay, double rad0, double reflect, double wband,
                          double* __restrict__ acc_dn, double* __restrict__ acc_up, size_t als) {
  constexpr int NB = 4;
  double rad = 0.;
  int l = nlay - 1;
  for (; l >= NB - 1; l -= NB) {
    double t[NB], s[NB], a[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const double* __restrict__ r = P + (size_t)(l - k) * ls;
      t[k] = r[0]; s[k] = r[rs];
      a[k] = acc_dn[(size_t)(l - k) * als];
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      rad = rad * t[k] + s[k];
      acc_dn[(size_t)(l - k) * als] = a[k] + rad * wband;
    }
  }
  for (; l >= 0; --l) {
    const double* __restrict__ r = P + (size_t)l * ls;
    rad = rad * r[0] + r[rs];
    acc_dn[(size_t)l * als] += rad * wband;
  }
  // surface (rtrn.f90:455-470)
  rad = rad0 + reflect * rad;
  acc_up[0] += rad * wband;
  l = 0;
  for (; l + NB <= nlay; l += NB) {
    double t[NB], s[NB], a[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const double* __restrict__ r = P + (size_t)(l + k) * ls;
      t[k] = r[0]; s[k] = r[2 * rs];
      a[k] = acc_up[(size_t)(l + k + 1) * als];
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      rad = rad * t[k] + s[k];
      acc_up[(size_t)(l + k + 1) * als] = a[k] + rad * wband;
    }
  }
  for (; l < nlay; ++l) {
    const double* __restrict__ r = P + (size_t)l * ls;
    rad = rad * r[0] + r[2 * rs];
    acc_up[(size_t)(l + 1) * als] += rad * wband;
  }
}

// band groups of the tile form: every tile of columns is processed once per group, the group's g-points one after the other;
// the groups' sums are reduced by lw_reduce_level like the unit groups of the other form.  4 groups of 38 + 38 + 32 + 32 g-points.
constexpr int kTileGroups = 4;
CB_HD void lw_tile_group_bands(int group, int& ib0, int& ib1) {  // 0-based bands [ib0, ib1)
  switch (group) {
    case 0: ib0 = 0; ib1 = 3; break;
    case 1: ib0 = 3; ib1 = 6; break;
    case 2: ib0 = 6; ib1 = 9; break;
    default: ib0 = 9; ib1 = 16; break;
  }
}
CB_HD int band_ngpt(int ib) {  // g-points of band ib (0-based): band_gstart differences
  return (ib == 15 ? 140 : band_gstart(ib + 1)) - band_gstart(ib);
}

// ---------------------------------------------------------------------------------------------
// Units: static list shared by host and device.  Bands are split in chunks of <= 4 g-points.
struct Unit {
  int band, g0, u;
};
#ifndef CB_LW_UMAX
#define CB_LW_UMAX 2      // g-points per thread of the transfer kernel (2 or 4)
#endif
#ifndef CB_LW_TAU_UMAX
#define CB_LW_TAU_UMAX 4  // g-points per thread of the taumol kernel (2 or 4)
#endif
constexpr int kMaxUnits = 140;  // one g-point per unit at most
inline int build_units(Unit* out, int umax) {  // host only
  int n = 0;
  // heavy (two-key-species, lower-atmosphere-rich) work first so the tail of the grid is made of light blocks
  for (int pass = 0; pass < 2; ++pass)
    for (int b = 1; b <= 16; ++b) {
      const bool heavy = kNSPA[b - 1] == 9;
      if ((pass == 0) != heavy) continue;
      const int ng = kNG[b - 1];
      for (int g0 = 0; g0 < ng; g0 += umax) {
        out[n].band = b;
        out[n].g0 = g0;
        out[n].u = (ng - g0) >= umax ? umax : (ng - g0);
        ++n;
      }
    }
  return n;
}

// lw_reduce: sum over the groups of units in a fixed order (deterministic; the band weights wtdiff * delwave of rtrn.f90:529-557
// were applied by the units) -> fluxes in W m-2.
CB_HD void lw_reduce_level(const Tables& T, const Work& W, int ngroups, int nlay, int c0, int c, int lev, int ncol, const Out& out) {
  const int ncc = W.ncc;
  const size_t pstride = (size_t)(nlay + 1) * ncc;
  // cloud-free column: the clear-sky streams equal the total ones bit for bit and were not stored (lw_transfer_unit)
  const bool cloudy_col = W.ncbands[c] > 0;
  const bool drv = out.duflx_dt != nullptr;
  const bool top = lev == nlay;  // no downward flux at the top interface (not stored)
  double tot[6] = {0., 0., 0., 0., 0., 0.};
  for (int k = 0; k < ngroups; ++k) {
    const double* p = W.part + (size_t)k * W.npart * pstride + (size_t)lev * ncc + c;
    tot[0] = tot[0] + p[0];
    if (!top) tot[1] = tot[1] + p[pstride];
    if (cloudy_col) {
      tot[2] = tot[2] + p[2 * pstride];
      if (!top) tot[3] = tot[3] + p[3 * pstride];
    }
    if (drv) {
      tot[4] = tot[4] + p[4 * pstride];
      if (cloudy_col) tot[5] = tot[5] + p[5 * pstride];
    }
  }
  if (!cloudy_col) { tot[2] = tot[0]; tot[3] = tot[1]; tot[5] = tot[4]; }
  const size_t o = (size_t)lev * ncol + (c0 + c);
  out.uflx[o] = tot[0] * T.fluxfac;
  out.dflx[o] = tot[1] * T.fluxfac;
  out.uflxc[o] = tot[2] * T.fluxfac;
  out.dflxc[o] = tot[3] * T.fluxfac;
  if (drv) { out.duflx_dt[o] = tot[4] * T.fluxfac; out.duflxc_dt[o] = tot[5] * T.fluxfac; }
}
// heating rates (rtrn.f90:569-581)
CB_HD void lw_heating(const Tables& T, const In& in, const Out& out, int gcol, int l) {
  const int ncol = in.ncol;
  const size_t o0 = (size_t)l * ncol + gcol, o1 = o0 + ncol;
  const double dp = in.plev[o0] - in.plev[o1];
  const double fnet0 = out.uflx[o0] - out.dflx[o0], fnet1 = out.uflx[o1] - out.dflx[o1];
  const double fnetc0 = out.uflxc[o0] - out.dflxc[o0], fnetc1 = out.uflxc[o1] - out.dflxc[o1];
  out.hr[o0] = T.heatfac * (fnet0 - fnet1) / dp;
  out.hrc[o0] = T.heatfac * (fnetc0 - fnetc1) / dp;
}

}  // namespace lw
}  // namespace cb
