// climt_b200 -- RRTMG shortwave engine, per-thread core (sm_100a device code; host-compilable for the
// test-only emulation).  Same decomposition as the longwave engine (lw_core.cuh):
//   sw_prep_column<LAYER,COLUMN>  inatm_sw + setcoef_sw + cldprop_sw (+ ECMWF aerosol mix) per (column, layer); laytrop and the
//                           solar-source layers per column
//   sw_taumol_unit<B,U>     one thread per (column, unit of <=4 g-points of band B, chunk of layers): taumol_sw
//   sw_transfer_unit<U,MC>  one thread per (column, unit of <=2 g-points), band-generic: delta-scaling + reftra_sw for the clear
//                           and cloudy paths, upward adding pass of vrtqdr_sw (surface -> top), then the downward pass that
//                           turns the stored layer properties into fluxes.  Lanes = adjacent columns.
//   sw_reduce_level / sw_heating   fixed-order sum of the per-unit partial fluxes, heating rates.
//
// Reference being replaced: climt/_lib/rrtmg_sw/rrtmg_sw_rad.nomcica.f90 (driver, inatm_sw), rrtmg_sw_setcoef.f90,
// rrtmg_sw_cldprop.f90, rrtmg_sw_taumol.f90, rrtmg_sw_spcvrt.f90, rrtmg_sw_reftra.f90, rrtmg_sw_vrtqdr.f90.
#pragma once
#include <cmath>

#include "cb_common.h"

namespace cb {
namespace sw {

constexpr int NBND = 14, NGPT = 112, NTBL = 10000;
enum WsField {
  F_FAC00 = 0, F_FAC01, F_FAC10, F_FAC11,
  F_COLH2O, F_COLCO2, F_COLO3, F_COLCH4, F_COLO2, F_COLMOL,
  F_SELFFAC, F_SELFFRAC, F_FORFAC, F_FORFRAC,
  NF
};
constexpr int NSCR = 16;  // per g-point scratch rows: 7 (ref, refd, tra, trad, dbt, rup, rupd) x {clear, total} + taug, taur

// absa / absb / selfref / forref are stored like the longwave band tables (lw_core.cuh, BandOff): per group of 4 consecutive
// g-points a contiguous block of rows x 4 doubles -- element (row r, g-point g) at X + (g / 4) * gs_X + r * 4 + (g % 4) -- so the
// 2-8 table rows a taumol thread gathers per layer are neighbouring 32-byte sectors.  The per-g-point vectors keep (rows, ng).
struct BandOff {
  int absa, absb, selfref, forref, sfluxref, irradnce, facbrght, snsptdrk;
  int gs_absa, gs_absb, gs_selfref, gs_forref;  // group strides (rows * 4) of the four regrouped tables
  int raylv;   // per-g Rayleigh vector (bands 23, 25, 26, 27), rayla (band 24, (9, ng)) else -1
  int raylb;   // band 24 upper
  int x0, x1;  // absch4 | abso3a, abso3b | absco2, absh2o
  double rayl; // scalar Rayleigh coefficient
};
struct Tables {
  const double* base;
  BandOff b[NBND];
  int preflog, tref, exp_tbl;
  int extliq1, ssaliq1, asyliq1, extice2, ssaice2, asyice2, extice3, ssaice3, asyice3, fdlice3;  // (n, 14)
  int abari, bbari, cbari, dbari, ebari, fbari;
  int rsrtaua, rsrpiza, rsrasya;  // (14, 6)
  double bpade, heatfac, oneminus, avogad, grav;
};
// Column-independent part of inatm_sw (solar constant / variability, earth-sun distance): computed on the host.
struct Solar {
  int isolvar;
  double adjflux[NBND];
  double svar_f, svar_s, svar_i;
  double svar_f_bnd[NBND], svar_s_bnd[NBND], svar_i_bnd[NBND];
};
struct In {  // reference ABI layout (rrtmg_sw_c_binder.f90:203-270)
  int ncol, nlay;
  const double *play, *plev, *tlay, *tlev, *tsfc, *h2o, *o3, *co2, *ch4, *n2o, *o2, *asdir, *asdif, *aldir, *aldif,
      *coszen, *cldfr, *taucld, *ssacld, *asmcld, *fsfcld, *cicewp, *cliqwp, *reice, *reliq, *tauaer, *ssaaer, *asmaer,
      *ecaer;
};
struct Out {
  double *uflx, *dflx, *hr, *uflxc, *dflxc, *hrc;
};
struct Flags {
  int icld, iaer, inflag, iceflag, liqflag;
  int mcica;  // 0: spcvrt (cloud fraction 0/1 per layer); 1: McICA (spcvmc: per-g-point 0/1 cloud mask)
};
struct Work {
  int ncc;
  double* ws;     // [NF][nlay][ncc]
  int* idx;       // [nlay][ncc] packed jp|jt|jt1|indself|indfor
  int* laytrop;   // [ncc]
  int* laysolfr;  // [14][ncc]   1-based layer that provides the solar source function of each band
  int* anycld;    // [ncc]
  double* cld;    // [3][14][nlay][ncc]  delta-scaled cloud tau, ssa, asym
  double* aer;    // [3][14][nlay][ncc]  aerosol tau, ssa, asym (iaer = 6 only)
  double* scr;    // [112][NSCR][nlay][ncc]
  double* src;    // [112][ncc]  solar source function of each g-point (taken at layer laysolfr of its band)
  double* part;   // [ngroups][4][nlay+1][ncc]   fu, fd, cu, cd  (already weighted by the incoming flux), summed over a group's units
  unsigned* mask; // [nlay][4][mstride] (+ moff) McICA cloud mask, bit (g & 31) of word (g >> 5)
  int mstride, moff;
  int* err;
};

CB_HD int pack_idx(int jp, int jt, int jt1, int inds, int indf) { return jp | (jt << 6) | (jt1 << 9) | (inds << 12) | (indf << 16); }

constexpr int kNG[14] = {6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12};
constexpr int kGS[14] = {0, 6, 18, 26, 34, 44, 54, 56, 66, 74, 80, 86, 94, 100};
// first g-point of band ib, for code that is generic in the band (a constexpr array cannot be indexed at run time on the device)
CB_HD int band_gstart(int ib) {
  switch (ib) {
    case 0: return 0; case 1: return 6; case 2: return 18; case 3: return 26; case 4: return 34; case 5: return 44;
    case 6: return 54; case 7: return 56; case 8: return 66; case 9: return 74; case 10: return 80; case 11: return 86;
    case 12: return 94; default: return 100;
  }
}
constexpr int kNSPA[14] = {9, 9, 9, 9, 1, 9, 9, 1, 9, 1, 0, 1, 9, 1};
constexpr int kNSPB[14] = {1, 5, 1, 1, 1, 5, 1, 0, 1, 0, 0, 1, 5, 1};

// ---------------------------------------------------------------------------------------------
// sw_prep_column: inatm_sw (rrtmg_sw_rad.nomcica.f90:1414-1537), setcoef_sw (rrtmg_sw_setcoef.f90:49-305),
// cldprop_sw (rrtmg_sw_cldprop.f90:53-365), ECMWF aerosol mix (rad.nomcica.f90:693-727), and the layer that supplies
// each band's solar source (the `laysolfr` logic of rrtmg_sw_taumol.f90, e.g. :334-337 and :586-590).
// Two launch modes (see lw_core.cuh prep_column): LAYER_PART = inatm_sw + setcoef_sw + the ECMWF aerosol mix of layers
// [l0, l1), independent per layer; COLUMN_PART = laytrop, cldprop_sw (routine-locals persist from layer to layer in the
// Fortran), the solar-source layers.  <true, true> over [0, nlay) is the original single pass.
// ECMWF aerosols (iaer = 6), rad.nomcica.f90:693-727: band optical properties of one (column, layer) from the six aerosol types.
// A function (and, on the device, a kernel of its own, launched only when iaer = 6) so that the option costs nothing when it is
// off.  (It is NOT what makes k_sw_prep_layer a 255-register kernel -- that is the per-band cloud-optics arrays of cldprop_sw
// in the same pass: ptxas reports the same 255 registers / 536-byte frame without this block.  The kernel is ~1 % of a step.)
CB_HD void sw_aerosol_mix(const Tables& T, const In& in, const Work& W, int c0, int c, int l) {
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const size_t gc = (size_t)(c0 + c);
  const double* tb = T.base;
  for (int ib = 0; ib < 14; ++ib) {
    double ta = 0., om = 0., as = 0.;
    for (int ia = 0; ia < 6; ++ia) {
      const double ec = in.ecaer[((size_t)ia * nlay + l) * ncol + gc];
      const double rt = tb[T.rsrtaua + ib * 6 + ia], rp = tb[T.rsrpiza + ib * 6 + ia], ra = tb[T.rsrasya + ib * 6 + ia];
      ta = ta + rt * ec;
      om = om + rt * ec * rp;
      as = as + rt * ec * rp * ra;
    }
    if (ta == 0.) { ta = 0.; as = 0.; om = 1.; }
    else {
      if (om != 0.) as = as / om;
      if (ta != 0.) om = om / ta;
    }
    W.aer[((size_t)(0 * 14 + ib) * nlay + l) * ncc + c] = ta;
    W.aer[((size_t)(1 * 14 + ib) * nlay + l) * ncc + c] = om;
    W.aer[((size_t)(2 * 14 + ib) * nlay + l) * ncc + c] = as;
  }
}

template <bool LAYER_PART, bool COLUMN_PART, bool AER6 = true>
CB_HD void sw_prep_column(const Tables& T, const In& in, const Flags& fl, const Work& W, int c0, int c, int l0, int l1) {
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const size_t gc = (size_t)(c0 + c);
  const double* tb = T.base;
  const double amd = 28.9660, amw = 18.0160;
  const double stpfac = 296. / 1013.;
  bool clouds = fl.icld >= 1;
  int laytrop = 0;
  bool anycld = false;
  if (COLUMN_PART && clouds) {
    // no cloudy layer -> nothing downstream reads the cloud optics (anycld = 0): skip cldprop_sw altogether
    int any = 0;  // no short circuit: the loads of different layers stay independent of each other
#pragma unroll 8
    for (int l = 0; l < nlay; ++l) any |= (int)(CB_LDG(in.cldfr + (size_t)l * ncol + gc) > 1.e-12);
    clouds = any != 0;
  }
  // cldprop_sw work arrays (each is reassigned for all 14 bands in every layer that enters the cloud block)
  double extcoice[14], gice[14], ssacoice[14], forwice[14], extcoliq[14], gliq[14], ssacoliq[14], forwliq[14];
  for (int i = 0; i < 14; ++i) { extcoice[i] = gice[i] = ssacoice[i] = forwice[i] = extcoliq[i] = gliq[i] = ssacoliq[i] = forwliq[i] = 0.; }
#define WS(f, l) W.ws[((size_t)(f) * nlay + (l)) * ncc + c]
  if (COLUMN_PART && !LAYER_PART && !clouds) {
    // cloud-free column: all that is left of the layer loop is the tropopause count; pressures of 8 layers are loaded together
    constexpr int NB = 8;
    for (int l = 0; l < nlay; l += NB) {
      double pm[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) pm[j] = CB_LDG(in.play + (size_t)(l + j < nlay ? l + j : nlay - 1) * ncol + gc);
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (l + j < nlay && !(log(pm[j]) <= 4.56)) laytrop = laytrop + 1;
    }
    l0 = l1;  // skip the general loop
  }
  for (int l = l0; l < l1; ++l) {
    const size_t o = (size_t)l * ncol + gc;
    const double pavel = in.play[o], tavel = in.tlay[o];
    const double pz_below = in.plev[o];
    const double pz = in.plev[o + ncol];
    double wkl[7];
    wkl[0] = in.h2o[o]; wkl[1] = in.co2[o]; wkl[2] = in.o3[o]; wkl[3] = in.n2o[o]; wkl[4] = 0.; wkl[5] = in.ch4[o]; wkl[6] = in.o2[o];
    const double amm = (1. - wkl[0]) * amd + wkl[0] * amw;
    const double coldry = (pz_below - pz) * 1.e3 * T.avogad / (1.e2 * T.grav * amm * (1. + wkl[0]));
    for (int i = 0; i < 7; ++i) wkl[i] = coldry * wkl[i];
    const double plog = log(pavel);
    if (COLUMN_PART && !(plog <= 4.56)) laytrop = laytrop + 1;  // rrtmg_sw_setcoef.f90:218
    double factor = 0.;
    if (LAYER_PART) {
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
      const double forfac = scalefac / (1. + water);
      double forfrac, selffac = 0., selffrac = 0.;
      int indfor, indself = 0;
      const bool lower = !(plog <= 4.56);
      if (lower) {
        factor = (332.0 - tavel) / 36.0;
        indfor = imin(2, imax(1, (int)factor));
        forfrac = factor - (double)indfor;
        selffac = water * forfac;
        factor = (tavel - 188.0) / 7.2;
        indself = imin(9, imax(1, (int)factor - 7));
        selffrac = factor - (double)(indself + 7);
      } else {
        factor = (tavel - 188.0) / 36.0;
        indfor = 3;
        forfrac = factor - 1.0;
      }
      const double colh2o = 1.e-20 * wkl[0];
      double colco2 = 1.e-20 * wkl[1];
      const double colo3 = 1.e-20 * wkl[2];
      double colch4 = 1.e-20 * wkl[5], colo2 = 1.e-20 * wkl[6];
      const double colmol = 1.e-20 * coldry + colh2o;
      if (colco2 == 0.) colco2 = 1.e-32 * coldry;
      if (colch4 == 0.) colch4 = 1.e-32 * coldry;
      if (colo2 == 0.) colo2 = 1.e-32 * coldry;
      const double compfp = 1. - fp;
      WS(F_FAC10, l) = compfp * ft;
      WS(F_FAC00, l) = compfp * (1. - ft);
      WS(F_FAC11, l) = fp * ft1;
      WS(F_FAC01, l) = fp * (1. - ft1);
      WS(F_COLH2O, l) = colh2o; WS(F_COLCO2, l) = colco2; WS(F_COLO3, l) = colo3; WS(F_COLCH4, l) = colch4;
      WS(F_COLO2, l) = colo2; WS(F_COLMOL, l) = colmol;
      WS(F_SELFFAC, l) = selffac; WS(F_SELFFRAC, l) = selffrac; WS(F_FORFAC, l) = forfac; WS(F_FORFRAC, l) = forfrac;
      W.idx[(size_t)l * ncc + c] = pack_idx(jp, jt, jt1, indself, indfor);
    }
    // ---- cloud optics (cldprop_sw / cldprmc_sw): every routine-local is (re)assigned for all 14 bands in each layer that is
    // entered, for every supported flag combination -> layer-independent
    if (LAYER_PART && fl.icld >= 1) {
      const double eps = 1.e-06, cldmin = 1.e-20;
      const double cldfrac = in.cldfr[o];
      if (!fl.mcica && cldfrac > 1.e-06 && cldfrac < T.oneminus) *W.err = 10;  // 'PARTIAL CLOUD NOT ALLOWED' (rad.nomcica.f90:616-620)
      const double ciwp = in.cicewp[o], clwp = in.cliqwp[o];
      // direct-input cloud optics; a null taucld means "all four arrays are zero" (host path with inflag != 0)
      const double zero14[14] = {0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.};
      const bool direct = in.taucld != nullptr;
      const double* tc = direct ? in.taucld + 14 * ((size_t)l * ncol + gc) : zero14;
      const double* sc = direct ? in.ssacld + 14 * ((size_t)l * ncol + gc) : zero14;
      const double* ac = direct ? in.asmcld + 14 * ((size_t)l * ncol + gc) : zero14;
      const double* fc = direct ? in.fsfcld + 14 * ((size_t)l * ncol + gc) : zero14;
      double tauctot = 0.;
      for (int ib = 0; ib < 14; ++ib) tauctot = tauctot + tc[ib];
      double taucloud[14], ssacloud[14], asmcloud[14];
      for (int ib = 0; ib < 14; ++ib) { taucloud[ib] = 0.; ssacloud[ib] = 1.; asmcloud[ib] = 0.; }
      const double cwp = ciwp + clwp;
      // McICA (cldprmc_sw, rrtmg_sw_cldprmc.f90:155-161): a cloudy sub-column carries the layer's water paths and its
      // band's direct-input optics, so the result only depends on the band; the entry test is per band there.
      bool enter = cldfrac >= cldmin && (cwp >= cldmin || tauctot >= cldmin);
      bool band_ok[14];
      for (int ib = 0; ib < 14; ++ib) band_ok[ib] = true;
      if (fl.mcica) {
        enter = false;
        for (int ib = 0; ib < 14; ++ib) {
          band_ok[ib] = cwp >= cldmin || tc[ib] >= cldmin;
          enter = enter || band_ok[ib];
          taucloud[ib] = tc[ib]; ssacloud[ib] = sc[ib]; asmcloud[ib] = ac[ib];
        }
        if (fl.inflag == 1 && enter) *W.err = 8;
      }
      if (enter) {
        if (fl.inflag == 0) {
          for (int ib = 0; ib < 14; ++ib) {
            if (!band_ok[ib]) continue;
            const double ffp = fc[ib], ffp1 = 1.0 - ffp, ffpssa = 1.0 - ffp * sc[ib];
            ssacloud[ib] = ffp1 * sc[ib] / ffpssa;
            taucloud[ib] = ffpssa * tc[ib];
            asmcloud[ib] = (ac[ib] - ffp) / (ffp1);
          }
        } else if (fl.inflag == 2) {
          const double radice = in.reice[o];
          if (ciwp == 0.0) {
            for (int ib = 0; ib < 14; ++ib) { extcoice[ib] = 0.; ssacoice[ib] = 0.; gice[ib] = 0.; forwice[ib] = 0.; }
          } else if (fl.iceflag == 1) {
            if (radice < 13.0 || radice > 130.) { *W.err = 2; }
            else {
              const double wavenum2[14] = {3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000., 38000., 50000., 2600.};
              for (int ib = 0; ib < 14; ++ib) {
                int icx = 5;
                if (wavenum2[ib] > 1.43e04) icx = 1;
                else if (wavenum2[ib] > 7.7e03) icx = 2;
                else if (wavenum2[ib] > 5.3e03) icx = 3;
                else if (wavenum2[ib] > 4.0e03) icx = 4;
                extcoice[ib] = tb[T.abari + icx - 1] + tb[T.bbari + icx - 1] / radice;
                ssacoice[ib] = 1. - tb[T.cbari + icx - 1] - tb[T.dbari + icx - 1] * radice;
                gice[ib] = tb[T.ebari + icx - 1] + tb[T.fbari + icx - 1] * radice;
                if (gice[ib] >= 1.0) gice[ib] = 1.0 - eps;
                forwice[ib] = gice[ib] * gice[ib];
                if (extcoice[ib] < 0.0 || ssacoice[ib] > 1.0 || ssacoice[ib] < 0.0 || gice[ib] > 1.0 || gice[ib] < 0.0) *W.err = 5;
              }
            }
          } else if (fl.iceflag == 2 || fl.iceflag == 3) {
            const double rmax = fl.iceflag == 2 ? 131.0 : 140.0;
            if (radice < 5.0 || radice > rmax) { *W.err = fl.iceflag == 2 ? 2 : 3; }
            else {
              factor = (radice - 2.) / 3.;
              int index = (int)factor;
              const int top = fl.iceflag == 2 ? 43 : 46;
              if (index == top) index = top - 1;
              const double fint = factor - (double)index;
              const int te = fl.iceflag == 2 ? T.extice2 : T.extice3, tsa = fl.iceflag == 2 ? T.ssaice2 : T.ssaice3,
                        tg = fl.iceflag == 2 ? T.asyice2 : T.asyice3;
              for (int ib = 0; ib < 14; ++ib) {
                const size_t r0 = (size_t)(index - 1) * 14 + ib, r1 = (size_t)index * 14 + ib;
                extcoice[ib] = tb[te + r0] + fint * (tb[te + r1] - tb[te + r0]);
                ssacoice[ib] = tb[tsa + r0] + fint * (tb[tsa + r1] - tb[tsa + r0]);
                gice[ib] = tb[tg + r0] + fint * (tb[tg + r1] - tb[tg + r0]);
                if (fl.iceflag == 2) {
                  forwice[ib] = gice[ib] * gice[ib];
                } else {
                  const double fdelta = tb[T.fdlice3 + r0] + fint * (tb[T.fdlice3 + r1] - tb[T.fdlice3 + r0]);
                  if (fdelta < 0.0 || fdelta > 1.0) *W.err = 6;
                  forwice[ib] = fdelta + 0.5 / ssacoice[ib];
                  if (forwice[ib] > gice[ib]) forwice[ib] = gice[ib];
                }
                if (extcoice[ib] < 0.0 || ssacoice[ib] > 1.0 || ssacoice[ib] < 0.0 || gice[ib] > 1.0 || gice[ib] < 0.0) *W.err = 5;
              }
            }
          }
          if (clwp == 0.0) {
            for (int ib = 0; ib < 14; ++ib) { extcoliq[ib] = 0.; ssacoliq[ib] = 0.; gliq[ib] = 0.; forwliq[ib] = 0.; }
          } else if (fl.liqflag == 1) {
            const double radliq = in.reliq[o];
            if (radliq < 2.5 || radliq > 60.) { *W.err = 4; }
            else {
              int index = (int)(radliq - 1.5);
              if (index == 0) index = 1;
              if (index == 58) index = 57;
              const double fint = radliq - 1.5 - (double)index;
              for (int ib = 0; ib < 14; ++ib) {
                const size_t r0 = (size_t)(index - 1) * 14 + ib, r1 = (size_t)index * 14 + ib;
                extcoliq[ib] = tb[T.extliq1 + r0] + fint * (tb[T.extliq1 + r1] - tb[T.extliq1 + r0]);
                ssacoliq[ib] = tb[T.ssaliq1 + r0] + fint * (tb[T.ssaliq1 + r1] - tb[T.ssaliq1 + r0]);
                if (fint < 0. && ssacoliq[ib] > 1.) ssacoliq[ib] = tb[T.ssaliq1 + r0];
                gliq[ib] = tb[T.asyliq1 + r0] + fint * (tb[T.asyliq1 + r1] - tb[T.asyliq1 + r0]);
                forwliq[ib] = gliq[ib] * gliq[ib];
                if (extcoliq[ib] < 0.0 || ssacoliq[ib] > 1.0 || ssacoliq[ib] < 0.0 || gliq[ib] > 1.0 || gliq[ib] < 0.0) *W.err = 7;
              }
            }
          }
          for (int ib = 0; ib < 14; ++ib) {
            if (!band_ok[ib]) continue;
            const double tauliqorig = clwp * extcoliq[ib], tauiceorig = ciwp * extcoice[ib];
            const double ssaliq = ssacoliq[ib] * (1.0 - forwliq[ib]) / (1.0 - forwliq[ib] * ssacoliq[ib]);
            const double tauliq = (1.0 - forwliq[ib] * ssacoliq[ib]) * tauliqorig;
            const double ssaice = ssacoice[ib] * (1.0 - forwice[ib]) / (1.0 - forwice[ib] * ssacoice[ib]);
            const double tauice = (1.0 - forwice[ib] * ssacoice[ib]) * tauiceorig;
            const double scatliq = ssaliq * tauliq;
            double scatice = ssaice * tauice;
            taucloud[ib] = tauliq + tauice;
            if (taucloud[ib] == 0.0) taucloud[ib] = cldmin;
            if (scatice == 0.0) scatice = cldmin;
            ssacloud[ib] = (scatliq + scatice) / taucloud[ib];
            if (fl.iceflag == 3) {
              asmcloud[ib] = (1.0 / (scatliq + scatice)) * (scatliq * (gliq[ib] - forwliq[ib]) / (1.0 - forwliq[ib]) +
                                                           scatice * ((gice[ib] - forwice[ib]) / (1.0 - forwice[ib])));
            } else {
              asmcloud[ib] = (scatliq * (gliq[ib] - forwliq[ib]) / (1.0 - forwliq[ib]) +
                              scatice * (gice[ib] - forwice[ib]) / (1.0 - forwice[ib])) / (scatliq + scatice);
            }
          }
        }
      }
      for (int ib = 0; ib < 14; ++ib) {
        W.cld[((size_t)(0 * 14 + ib) * nlay + l) * ncc + c] = taucloud[ib];
        W.cld[((size_t)(1 * 14 + ib) * nlay + l) * ncc + c] = ssacloud[ib];
        W.cld[((size_t)(2 * 14 + ib) * nlay + l) * ncc + c] = asmcloud[ib];
      }
    }
    if (LAYER_PART && AER6 && fl.iaer == 6) sw_aerosol_mix(T, in, W, c0, c, l);
  }
#undef WS
  if (!COLUMN_PART) return;
  anycld = clouds;  // `clouds` already means: some layer has cldfr > 1e-12
  W.laytrop[c] = laytrop;
  W.anycld[c] = (clouds && anycld) ? 1 : 0;
  // Layer whose key-species ratio selects each band's solar source function.  The Fortran updates `laysolfr`
  // while looping over layers and assigns the source when `lay == laysolfr`; the last assignment wins
  // (lower-atmosphere form e.g. taumol18 :586-590, upper form e.g. taumol16 :334-337; band 26 :1455 has no trigger).
  // One pass over the layers serves all 14 bands (jp of layers lay-1, lay, lay+1 in registers): 60 independent loads instead of the
  // 14 x 60 dependent ones of a band-by-band search (r01 launch list: this kernel was 92 us of a 2.7 ms step).
  {
    const int layreffr[14] = {18, 30, 6, 3, 3, 8, 2, 6, 1, 2, 0, 32, 58, 49};
    const bool lower_src[14] = {false, false, true, true, true, true, true, true, true, true, true, false, false, false};
    int ls[14];
#pragma unroll
    for (int b = 0; b < 14; ++b) ls[b] = lower_src[b] ? laytrop : nlay;
    int jpm = 0;                                   // jp(lay - 1); 0 below the first layer
    int jp0 = W.idx[c] & 63;                       // jp(lay)
    constexpr int NB = 8;
    for (int base = 1; base <= nlay; base += NB) {
      int nxt[NB];                                 // jp(lay + 1) of the next 8 layers, loaded together; 0 above the last layer
#pragma unroll
      for (int j = 0; j < NB; ++j) nxt[j] = base + j < nlay ? (W.idx[(size_t)(base + j) * ncc + c] & 63) : 0;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int lay = base + j;
        if (lay > nlay) break;
        const int jp1 = nxt[j];
#pragma unroll
        for (int b = 0; b < 14; ++b) {
          if (lower_src[b]) {
            if (lay <= laytrop && b != 10 && jp0 < layreffr[b] && jp1 >= layreffr[b]) ls[b] = imin(lay + 1, laytrop);
          } else {
            if (lay > laytrop && jpm < layreffr[b] && jp0 >= layreffr[b]) ls[b] = lay;
          }
        }
        jpm = jp0;
        jp0 = jp1;
      }
    }
#pragma unroll
    for (int b = 0; b < 14; ++b) W.laysolfr[(size_t)b * ncc + c] = ls[b];
  }
}

// ---------------------------------------------------------------------------------------------
// Band descriptions (rrtmg_sw_taumol.f90: taumol16 :275-389 ... taumol29 :1695-1787).
enum Gas { H2O = 0, CO2 = 1, O3 = 2, CH4 = 3, O2 = 4 };  // F_COLH2O + id
enum Extra { X_NONE, X_CH4, X_O2CONT, X_O3, X_CO2, X_H2O };
enum RaylKind { RAY_SCALAR, RAY_VEC, RAY_A_INTERP, RAY_B };
struct Region {
  int kind;        // 0 none, 1 one key species, 2 two key species
  int a, b;
  double strrat;   // binary ratio (band 22: o2adj*strrat)
  double keyscale; // band 23: givfac on the key term; band 22 upper: o2adj
  bool self, forn;
  bool inside;     // self/foreign continuum inside the colh2o*( ) bracket together with the key term
  int extra;       // additional absorber
  int xslot;       // BandOff x0/x1
  int rayl;
};
constexpr Region mk(int kind, int a, int b, double strrat, double keyscale, bool self, bool forn, bool inside, int extra,
                    int xslot, int rayl) {
  return {kind, a, b, strrat, keyscale, self, forn, inside, extra, xslot, rayl};
}
template <int B, bool LOWER>
constexpr Region region() {
  // B = 16..29
  return B == 16 ? (LOWER ? mk(2, H2O, CH4, 252.131, 1., true, true, false, X_NONE, 0, RAY_SCALAR)
                          : mk(1, CH4, 0, 0., 1., false, false, false, X_NONE, 0, RAY_SCALAR))
       : B == 17 ? (LOWER ? mk(2, H2O, CO2, 0.364641, 1., true, true, false, X_NONE, 0, RAY_SCALAR)
                          : mk(2, H2O, CO2, 0.364641, 1., false, true, false, X_NONE, 0, RAY_SCALAR))
       : B == 18 ? (LOWER ? mk(2, H2O, CH4, 38.9589, 1., true, true, false, X_NONE, 0, RAY_SCALAR)
                          : mk(1, CH4, 0, 0., 1., false, false, false, X_NONE, 0, RAY_SCALAR))
       : B == 19 ? (LOWER ? mk(2, H2O, CO2, 5.49281, 1., true, true, false, X_NONE, 0, RAY_SCALAR)
                          : mk(1, CO2, 0, 0., 1., false, false, false, X_NONE, 0, RAY_SCALAR))
       : B == 20 ? (LOWER ? mk(1, H2O, 0, 0., 1., true, true, true, X_CH4, 0, RAY_SCALAR)
                          : mk(1, H2O, 0, 0., 1., false, true, true, X_CH4, 0, RAY_SCALAR))
       : B == 21 ? (LOWER ? mk(2, H2O, CO2, 0.0045321, 1., true, true, false, X_NONE, 0, RAY_SCALAR)
                          : mk(2, H2O, CO2, 0.0045321, 1., false, true, false, X_NONE, 0, RAY_SCALAR))
       : B == 22 ? (LOWER ? mk(2, H2O, O2, 1.6 * 0.022708, 1., true, true, false, X_O2CONT, 0, RAY_SCALAR)
                          : mk(1, O2, 0, 0., 1.6, false, false, false, X_O2CONT, 0, RAY_SCALAR))
       : B == 23 ? (LOWER ? mk(1, H2O, 0, 0., 1.029, true, true, true, X_NONE, 0, RAY_VEC)
                          : mk(0, 0, 0, 0., 1., false, false, false, X_NONE, 0, RAY_VEC))
       : B == 24 ? (LOWER ? mk(2, H2O, O2, 0.124692, 1., true, true, false, X_O3, 0, RAY_A_INTERP)
                          : mk(1, O2, 0, 0., 1., false, false, false, X_O3, 1, RAY_B))
       : B == 25 ? (LOWER ? mk(1, H2O, 0, 0., 1., false, false, false, X_O3, 0, RAY_VEC)
                          : mk(0, 0, 0, 0., 1., false, false, false, X_O3, 1, RAY_VEC))
       : B == 26 ? mk(0, 0, 0, 0., 1., false, false, false, X_NONE, 0, RAY_VEC)
       : B == 27 ? mk(1, O3, 0, 0., 1., false, false, false, X_NONE, 0, RAY_VEC)
       : B == 28 ? mk(2, O3, O2, 6.67029e-07, 1., false, false, false, X_NONE, 0, RAY_SCALAR)
       :           (LOWER ? mk(1, H2O, 0, 0., 1., true, true, true, X_CO2, 0, RAY_SCALAR)
                          : mk(1, CO2, 0, 0., 1., false, false, false, X_H2O, 1, RAY_SCALAR));
}
// does the band's solar source depend on the key-species ratio at layer laysolfr?
template <int B> constexpr bool src_interp() { return B == 17 || B == 18 || B == 19 || B == 21 || B == 22 || B == 24 || B == 28; }
template <int B> constexpr bool src_lower() { return B >= 18 && B <= 26; }

// Optical depths (gas, Rayleigh) for U consecutive g-points of band B in one layer; optionally the solar source.
template <int B, bool LOWER, int U>
CB_HD void eval_band(const Tables& T, const double* __restrict__ ws, size_t wstride, int idx, int g0,
                     double* __restrict__ taug, double* __restrict__ taur, bool want_src, const Solar& sol,
                     double* __restrict__ src) {
  constexpr Region R = region<B, LOWER>();
  constexpr int ib = B - 16;
  constexpr int ng = kNG[ib];
  constexpr int RS = 4;  // row stride of the regrouped tables (see BandOff)
  const BandOff& O = T.b[ib];
  const double* __restrict__ tb = T.base;
#define WSF(f) CB_LDG(ws + (size_t)(f) * wstride)
  const int jp = idx & 63, jt = (idx >> 6) & 7, jt1 = (idx >> 9) & 7, inds = (idx >> 12) & 15, indf = (idx >> 16) & 3;
  double acc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = 0.;
  const double colh2o = WSF(F_COLH2O);
  int js = 0;
  double fs = 0.;
  CB_COV(1, B, LOWER, COV_REGION);
  // water-vapour continua (SW: un-premultiplied selffac/forfac, rrtmg_sw_setcoef.f90:232,238)
  double cont[U];
#pragma unroll
  for (int u = 0; u < U; ++u) cont[u] = 0.;
  if (R.self || R.forn) {
    const double forfac = WSF(F_FORFAC), forfrac = WSF(F_FORFRAC);
    const double* __restrict__ f = tb + O.forref + (size_t)(g0 / RS) * O.gs_forref + (size_t)(indf - 1) * RS + (g0 % RS);
    if (R.self) {
      const double selffac = WSF(F_SELFFAC), selffrac = WSF(F_SELFFRAC);
      const double* __restrict__ s = tb + O.selfref + (size_t)(g0 / RS) * O.gs_selfref + (size_t)(inds - 1) * RS + (g0 % RS);
      const Row<U> s0 = ldrow<U>(s), s1 = ldrow<U>(s + RS), f0 = ldrow<U>(f), f1 = ldrow<U>(f + RS);
#pragma unroll
      for (int u = 0; u < U; ++u)
        cont[u] = selffac * (s0[u] + selffrac * (s1[u] - s0[u])) + forfac * (f0[u] + forfrac * (f1[u] - f0[u]));
      if (colh2o * selffac * (s0[0] + selffrac * (s1[0] - s0[0])) > 0.) CB_COV(1, B, LOWER, COV_SELF_NONZERO);
      if (colh2o * forfac * (f0[0] + forfrac * (f1[0] - f0[0])) > 0.) CB_COV(1, B, LOWER, COV_FOR_NONZERO);
    } else {
      const Row<U> f0 = ldrow<U>(f), f1 = ldrow<U>(f + RS);
#pragma unroll
      for (int u = 0; u < U; ++u) cont[u] = forfac * (f0[u] + forfrac * (f1[u] - f0[u]));
      if (colh2o * cont[0] > 0.) CB_COV(1, B, LOWER, COV_FOR_NONZERO);
    }
  }
  if (R.kind == 2) {
    const double fac00 = WSF(F_FAC00), fac01 = WSF(F_FAC01), fac10 = WSF(F_FAC10), fac11 = WSF(F_FAC11);
    constexpr double n = LOWER ? 8. : 4.;
    constexpr int nsp = LOWER ? 9 : 5;
    const double cola = WSF(F_COLH2O + R.a), colb = WSF(F_COLH2O + R.b);
    const double speccomb = cola + R.strrat * colb;
    double specparm = cola / speccomb;
    if (specparm >= T.oneminus) specparm = T.oneminus;
    const double specmult = n * specparm;
    js = 1 + (int)specmult;
    fs = fmod(specmult, 1.);
    CB_COV(1, B, LOWER, specparm < 0.125 ? COV_S0_LOW : (specparm > 0.875 ? COV_S0_HIGH : COV_S0_MID));  // (SW has one stencil: bins only)
    if (speccomb > 0.) CB_COV(1, B, LOWER, COV_KEY_NONZERO);
    const double f000 = (1. - fs) * fac00, f010 = (1. - fs) * fac10, f100 = fs * fac00, f110 = fs * fac10;
    const double f001 = (1. - fs) * fac01, f011 = (1. - fs) * fac11, f101 = fs * fac01, f111 = fs * fac11;
    const int row0 = (LOWER ? ((jp - 1) * 5 + (jt - 1)) * nsp : ((jp - 13) * 5 + (jt - 1)) * nsp) + js - 1;
    const int row1 = (LOWER ? (jp * 5 + (jt1 - 1)) * nsp : ((jp - 12) * 5 + (jt1 - 1)) * nsp) + js - 1;
    const double* __restrict__ ab = tb + (LOWER ? O.absa : O.absb) + (size_t)(g0 / RS) * (LOWER ? O.gs_absa : O.gs_absb) + (g0 % RS);
    const double* __restrict__ a0 = ab + (size_t)row0 * RS;
    const double* __restrict__ a1 = ab + (size_t)row1 * RS;
    const Row<U> k000 = ldrow<U>(a0), k100 = ldrow<U>(a0 + RS), k010 = ldrow<U>(a0 + nsp * RS), k110 = ldrow<U>(a0 + (nsp + 1) * RS);
    const Row<U> k001 = ldrow<U>(a1), k101 = ldrow<U>(a1 + RS), k011 = ldrow<U>(a1 + nsp * RS), k111 = ldrow<U>(a1 + (nsp + 1) * RS);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double d = f000 * k000[u];
      d = d + f100 * k100[u];
      d = d + f010 * k010[u];
      d = d + f110 * k110[u];
      d = d + f001 * k001[u];
      d = d + f101 * k101[u];
      d = d + f011 * k011[u];
      d = d + f111 * k111[u];
      acc[u] = speccomb * d;
      if (R.self || R.forn) acc[u] = acc[u] + colh2o * cont[u];
    }
  } else if (R.kind == 1) {
    const double fac00 = WSF(F_FAC00), fac01 = WSF(F_FAC01), fac10 = WSF(F_FAC10), fac11 = WSF(F_FAC11);
    constexpr int nsp = LOWER ? kNSPA[ib] : kNSPB[ib];
    const int row0 = LOWER ? ((jp - 1) * 5 + (jt - 1)) * nsp : ((jp - 13) * 5 + (jt - 1)) * nsp;
    const int row1 = LOWER ? (jp * 5 + (jt1 - 1)) * nsp : ((jp - 12) * 5 + (jt1 - 1)) * nsp;
    const double cola = WSF(F_COLH2O + R.a);
    if (cola > 0.) CB_COV(1, B, LOWER, COV_KEY_NONZERO);
    const double* __restrict__ ab = tb + (LOWER ? O.absa : O.absb) + (size_t)(g0 / RS) * (LOWER ? O.gs_absa : O.gs_absb) + (g0 % RS);
    const double* __restrict__ a0 = ab + (size_t)row0 * RS;
    const double* __restrict__ a1 = ab + (size_t)row1 * RS;
    const Row<U> k00 = ldrow<U>(a0), k10 = ldrow<U>(a0 + RS), k01 = ldrow<U>(a1), k11 = ldrow<U>(a1 + RS);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double k4 = fac00 * k00[u] + fac10 * k10[u] + fac01 * k01[u] + fac11 * k11[u];
      if (R.inside) {
        // e.g. taumol20 :830-842 / taumol23 :1179-1189 / taumol29 :1735-1746 ; upper 20: :857-866
        if (R.keyscale != 1.) acc[u] = cola * (R.keyscale * k4 + cont[u]);
        else acc[u] = cola * (k4 + cont[u]);
      } else {
        acc[u] = (R.keyscale != 1.) ? cola * R.keyscale * k4 : cola * k4;
      }
    }
  }
  // additional absorbers
  if (R.extra == X_CH4 || R.extra == X_O3 || R.extra == X_CO2 || R.extra == X_H2O) {
    const int gas = R.extra == X_CH4 ? CH4 : (R.extra == X_O3 ? O3 : (R.extra == X_CO2 ? CO2 : H2O));
    const double colx = WSF(F_COLH2O + gas);
    if (colx > 0.) CB_COV(1, B, LOWER, COV_XSEC0_NONZERO);
    const double* __restrict__ x = tb + (R.xslot == 0 ? O.x0 : O.x1) + g0;
    const Row<U> xr = ldrow<U>(x);
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = acc[u] + colx * xr[u];
  } else if (R.extra == X_O2CONT) {
    const double o2cont = 4.35e-4 * WSF(F_COLO2) / (350.0 * 2.0);
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = acc[u] + o2cont;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) taug[u] = acc[u];
  // Rayleigh
  const double colmol = WSF(F_COLMOL);
  if (R.rayl == RAY_SCALAR) {
#pragma unroll
    for (int u = 0; u < U; ++u) taur[u] = colmol * O.rayl;
  } else if (R.rayl == RAY_VEC) {
#pragma unroll
    for (int u = 0; u < U; ++u) taur[u] = colmol * CB_LDG(tb + O.raylv + g0 + u);
  } else if (R.rayl == RAY_B) {
#pragma unroll
    for (int u = 0; u < U; ++u) taur[u] = colmol * CB_LDG(tb + O.raylb + g0 + u);
  } else {
    const double* __restrict__ r = tb + O.raylv + (size_t)(js - 1) * ng + g0;
    const Row<U> r0 = ldrow<U>(r), r1 = ldrow<U>(r + ng);
#pragma unroll
    for (int u = 0; u < U; ++u) taur[u] = colmol * (r0[u] + fs * (r1[u] - r0[u]));
  }
  // solar source function at this layer (only evaluated at layer laysolfr)
  if (want_src) {
    constexpr bool interp = src_interp<B>();
    if (interp) CB_COV(1, B, LOWER, COV_PLANCK_INTERP);
    auto val = [&](int off, int u) {
      if (interp) {
        const double* __restrict__ t = tb + off + (size_t)(js - 1) * ng + g0 + u;
        const double v0 = CB_LDG(t), v1 = CB_LDG(t + ng);
        return v0 + fs * (v1 - v0);
      }
      return CB_LDG(tb + off + g0 + u);
    };
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (sol.isolvar < 0) {
        src[u] = (B == 27) ? (50.15 / 48.37) * val(O.sfluxref, u) : val(O.sfluxref, u);
      } else if (sol.isolvar <= 2) {
        src[u] = sol.svar_f * val(O.facbrght, u) + sol.svar_s * val(O.snsptdrk, u) + sol.svar_i * val(O.irradnce, u);
      } else {
        src[u] = sol.svar_f_bnd[ib] * val(O.facbrght, u) + sol.svar_s_bnd[ib] * val(O.snsptdrk, u) +
                 sol.svar_i_bnd[ib] * val(O.irradnce, u);
      }
    }
  }
#undef WSF
}

// exp(-x) through the reference's 10 001-entry table or its small-argument series (spcvrt.f90:463-470 etc.)
CB_HD double exp_neg(const double* __restrict__ exp_tbl, double bpade, double ze1) {
  if (ze1 <= 0.06) return 1. - ze1 + 0.5 * ze1 * ze1;
  const int itind = tbl_slot(ze1, bpade, 10000.0);  // int(10000 ze1 / (bpade + ze1) + 0.5), the reference's slot (cb_common.h)
  return CB_LDG(exp_tbl + itind);
}

// reftra_sw for one layer (rrtmg_sw_reftra.f90:148-317, kmodts = 2)
CB_HD void reftra(const double* __restrict__ exp_tbl, double bpade, double zg, double prmuz, double zto1, double zw,
                  double& pref, double& prefd, double& ptra, double& ptrad) {
  const double eps = 1.e-08, zwcrit = 0.9999995;
  const double zg3 = 3. * zg;
  const double zgamma1 = (8. - zw * (5. + zg3)) * 0.25;
  const double zgamma2 = 3. * (zw * (1. - zg)) * 0.25;
  const double zgamma3 = (2. - zg3 * prmuz) * 0.25;
  const double zgamma4 = 1. - zgamma3;
  const double r = fdiv(zg, 1. - zg);
  const double zwo = fdiv(zw, 1. - (1. - zw) * (r * r));
  if (zwo >= zwcrit) {
    const double za = zgamma1 * prmuz;
    const double za1 = za - zgamma3;
    const double zgt = zgamma1 * zto1;
    const double ze1 = fmin(fdiv(zto1, prmuz), 500.);
    const double ze2 = exp_neg(exp_tbl, bpade, ze1);
    pref = fdiv(zgt - za1 * (1. - ze2), 1. + zgt);
    ptra = 1. - pref;
    prefd = fdiv(zgt, 1. + zgt);
    ptrad = 1. - prefd;
    if (ze2 == 1.0) { pref = 0.0; ptra = 1.0; prefd = 0.0; ptrad = 1.0; }
  } else {
    const double za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3;
    const double za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4;
    const double zrk = sqrt(zgamma1 * zgamma1 - zgamma2 * zgamma2);
    const double zrp = zrk * prmuz;
    const double zrp1 = 1. + zrp, zrm1 = 1. - zrp, zrk2 = 2. * zrk, zrpp = 1. - zrp * zrp, zrkg = zrk + zgamma1;
    const double zr1 = zrm1 * (za2 + zrk * zgamma3);
    const double zr2 = zrp1 * (za2 - zrk * zgamma3);
    const double zr3 = zrk2 * (zgamma3 - za2 * prmuz);
    const double zr4 = zrpp * zrkg;
    const double zr5 = zrpp * (zrk - zgamma1);
    const double zt1 = zrp1 * (za1 + zrk * zgamma4);
    const double zt2 = zrm1 * (za1 - zrk * zgamma4);
    const double zt3 = zrk2 * (zgamma4 + za1 * prmuz);
    const double zbeta = fdiv(zgamma1 - zrk, zrkg);
    const double ze1 = fmin(zrk * zto1, 500.);
    const double ze2 = fmin(fdiv(zto1, prmuz), 500.);
    const double zem1 = exp_neg(exp_tbl, bpade, ze1), zep1 = frcp(zem1);
    const double zem2 = exp_neg(exp_tbl, bpade, ze2), zep2 = frcp(zem2);
    const double zdenr = zr4 * zep1 + zr5 * zem1;
    const double zdent = zr4 * zep1 + zr5 * zem1;  // zt4 = zr4, zt5 = zr5
    if (zdenr >= -eps && zdenr <= eps) {
      pref = eps;
      ptra = zem2;
    } else {
      pref = fdiv(zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2), zdenr);
      ptra = zem2 - fdiv(zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2), zdent);
    }
    const double zemm = zem1 * zem1;
    const double zdend = frcp((1. - zbeta * zemm) * zrkg);
    prefd = zgamma2 * (1. - zemm) * zdend;
    ptrad = zrk2 * zem1 * zdend;
  }
}

// ---------------------------------------------------------------------------------------------
// The per-g-point work is split in two kernels (r01 ncu: the fused version stalled 26 % of its issue slots on
// instruction fetch -- 28 band variants of ~55 KB each -- and ran 12 warps per SM at 168 registers):
//   sw_taumol_unit<B,U>   band-specialised, layers independent: gas optical depth + Rayleigh depth of U g-points
//                         (taumol_sw) for a chunk of layers -> two scratch rows per g-point; solar source at laysolfr
//   sw_transfer_unit<U,MC> ONE code body for every band: delta-scaling, reftra_sw, adding (spcvrt_sw / spcvmc_sw)
constexpr int R_TAUG = 14, R_TAUR = 15;  // scratch rows written by sw_taumol_unit (rows 0-13: sw_transfer_unit)

template <int B, int U>
CB_HD void sw_taumol_unit(const Tables& T, const Solar& sol, const In& in, const Work& W, int c0, int c, int g0, int l0, int l1) {
  constexpr int ib = B - 16;
  const int nlay = in.nlay, ncc = W.ncc;
  const size_t wstride = (size_t)nlay * ncc;
  const int laytrop = W.laytrop[c];
  const int laysolfr = W.laysolfr[(size_t)ib * ncc + c];
  const int gabs = kGS[ib] + g0;
  for (int l = l0; l < l1; ++l) {
    const int lay = l + 1;
    const int idx = W.idx[(size_t)l * ncc + c];
    const double* ws = W.ws + (size_t)l * ncc + c;
    double taug[U], taur[U], src[U];
    const bool want = lay == laysolfr;
    if (lay <= laytrop) eval_band<B, true, U>(T, ws, wstride, idx, g0, taug, taur, want, sol, src);
    else eval_band<B, false, U>(T, ws, wstride, idx, g0, taug, taur, want, sol, src);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double* __restrict__ scr = W.scr + (((size_t)(gabs + u) * NSCR) * nlay + l) * ncc + c;
      scr[R_TAUG * wstride] = taug[u];
      scr[R_TAUR * wstride] = taur[u];
      if (want) W.src[(size_t)(gabs + u) * ncc + c] = src[u];
    }
  }
}

// Layer optical properties of one g-point in one layer: delta-scaled clear-sky mix (gas + Rayleigh + aerosol), two-stream R/T and
// direct-beam transmittance (spcvrt.f90:430-514), and -- cloudy column -- the same for the layer's overcast part mixed by the
// cloud fraction (:516-588).  Used by both sweeps of sw_transfer_unit when the layer properties are recomputed instead of carried.
struct SwLayer {
  double refc, refdc, trac, tradc, dbtc;  // clear
  double ref, refd, tra, trad, dbt;       // total (cloud-fraction mix); only set for cloudy columns
};
CB_HD SwLayer sw_layer_props(const double* __restrict__ exp_tbl, double bpade, double prmu0, double taug, double taur,
                             double ptaua, double pomga, double pasya, bool cloudy_col, double pclfr, double ptauc,
                             double pomgc, double pasyc) {
  SwLayer L;
  double ztauc = taur + taug + ptaua;
  double zomcc = taur * 1.0 + ptaua * pomga;
  double zgcc = fdiv(pasya * pomga * ptaua, zomcc);
  zomcc = fdiv(zomcc, ztauc);
  const double zf = zgcc * zgcc;
  const double zwf = zomcc * zf;
  ztauc = (1.0 - zwf) * ztauc;
  zomcc = fdiv(zomcc - zwf, 1.0 - zwf);
  zgcc = fdiv(zgcc - zf, 1.0 - zf);
  reftra(exp_tbl, bpade, zgcc, prmu0, ztauc, zomcc, L.refc, L.refdc, L.trac, L.tradc);
  L.dbtc = exp_neg(exp_tbl, bpade, fdiv(ztauc, prmu0));
  L.ref = L.refd = L.tra = L.trad = L.dbt = 0.;
  if (cloudy_col) {
    // cloudy (overcast) two-stream of this layer and the cloud-fraction mix (spcvrt.f90:516-588)
    const double ztauo = ztauc + ptauc;
    double zomco = ztauc * zomcc + ptauc * pomgc;
    const double zgco = fdiv(ptauc * pomgc * pasyc + ztauc * zomcc * zgcc, zomco);
    zomco = fdiv(zomco, ztauo);
    double refo = 0., refdo = 0., trao = 1., trado = 1.;
    if (pclfr > 1.e-12) reftra(exp_tbl, bpade, zgco, prmu0, ztauo, zomco, refo, refdo, trao, trado);
    const double zclear = 1.0 - pclfr, zcloud = pclfr;
    L.ref = zclear * L.refc + zcloud * refo;
    L.refd = zclear * L.refdc + zcloud * refdo;
    L.tra = zclear * L.trac + zcloud * trao;
    L.trad = zclear * L.tradc + zcloud * trado;
    const double dbtmo = exp_neg(exp_tbl, bpade, fdiv(ztauo, prmu0));
    L.dbt = zclear * L.dbtc + zcloud * dbtmo;
  }
  return L;
}

#ifndef CB_SW_DISCARD
#define CB_SW_DISCARD 0  // 1: the downward sweep discards the L2 lines of the rows it has consumed (discard.global.L2, SASS CCTL.RML2).
                         // r02 B200, 8192 x 60: DRAM traffic 7.18 -> 6.63 GB as hoped, kernel time 1.83 -> 11.1 ms: the cache-control
                         // operations serialise.  Off.
#endif
#ifndef CB_SW_RECOMPUTE
#define CB_SW_RECOMPUTE 0  // 1: the downward sweep recomputes the layer properties from (taug, taur) instead of reading the five rows
                           // the upward sweep stored: 4 instead of 14 scratch rows carried per g-point (clear sky), reftra twice
#endif

// Where the downward sweep puts the flux contributions of its g-points, one call per interface from the top down.  The partial
// sums of a GROUP of CB_SW_GROUP units share one set of rows `part[group][4][nlay+1][ncc]`:
//   * CUDA kernel (sw_engine.cu, SwPartSmem): the units of a group are the warps of one block; values are staged in shared memory a
//     few levels at a time and summed over the warps in unit order before they reach HBM -- r01: the per-unit rows and the
//     kernel that re-read them were 19 % of a step's DRAM traffic;
//   * host emulation (SwPartDirect): the units of a group run one after the other and accumulate into the zeroed rows in the
//     same order, so both give the same bits.
#ifndef CB_SW_GROUP
#define CB_SW_GROUP 4
#endif
struct SwPartDirect {
  double* part;  // rows of this unit's group, at this column
  size_t pstride;
  int ncc;
  CB_HD void put(size_t lev, bool cloudy_col, double sfu, double sfd, double scu, double scd) {
    if (cloudy_col) {  // cloud-free column: total == clear, not stored (sw_reduce_level copies)
      part[0 * pstride + lev * ncc] += sfu;
      part[1 * pstride + lev * ncc] += sfd;
    }
    part[2 * pstride + lev * ncc] += scu;
    part[3 * pstride + lev * ncc] += scd;
  }
  CB_HD void finish() {}
};

// Where the upward sweep parks what the downward sweep needs again (rows 0-13 of NSCR: five layer properties + rup, rupd, clear
// and total).  Default (p == nullptr): the unit's own rows of W.scr.  The slab form of the CUDA kernel (sw_engine.cu,
// k_sw_transfer_slab) points it at a per-warp slab that the warp reuses for every unit it processes: the rows are written bottom-up
// and read back top-down (last in, first out), so a slab that fits the L2 share of its warp never reaches HBM.
struct Carry {
  double* p;          // at this thread's lane
  size_t rs, ls, us;  // strides between rows, layers and the g-points of a unit
};

// spcvrt_sw / spcvmc_sw for U consecutive g-points of band ib (rrtmg_sw_spcvrt.f90:329-661) -- generic in the band.
// SLAB: the carried rows live in an L2-resident slab (cy given): cache-residency hints of cb_common.h on every access, the taumol
// rows fetched kSlabAhead layers ahead into the L2.
constexpr int kSlabAhead = 4;
template <int U, bool MC, class Sink, bool RECOMPUTE = (CB_SW_RECOMPUTE != 0), bool SLAB = false>
CB_HD void sw_transfer_unit(const Tables& T, const Solar& sol, const In& in, const Flags& fl, const Work& W, int c0, int c,
                            int ib, int g0, Sink& sink, Carry cy = Carry{nullptr, 0, 0, 0}) {
  const unsigned long long pol_keep = SLAB ? policy_keep() : 0ull, pol_drop = SLAB ? policy_drop() : 0ull;
  auto cst = [&](double* p, double v) { if (SLAB) st_policy(p, v, pol_keep); else *p = v; };
  auto cld = [&](const double* p) { return SLAB ? ld_policy(p, pol_drop) : *p; };
  auto sld = [&](const double* p) { return SLAB ? ld_stream(p) : *p; };
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const size_t gc = (size_t)(c0 + c);
  const double* __restrict__ tb = T.base;
  const double* __restrict__ exp_tbl = tb + T.exp_tbl;
  const double bpade = T.bpade;
  const size_t wstride = (size_t)nlay * ncc;
  const bool cloudy_col = W.anycld[c] != 0;
  double prmu0 = in.coszen[gc];
  if (prmu0 < 1.e-10) prmu0 = 1.e-10;
  // albedo by band (rad.nomcica.f90:648-659): bands 16-24 and 29 near-IR, 25-28 UV/visible
  const bool nir = (ib <= 8) || ib == 13;
  const double albdir = nir ? in.aldir[gc] : in.asdir[gc];
  const double albdif = nir ? in.aldif[gc] : in.asdif[gc];
  const int gabs = band_gstart(ib) + g0;
  const size_t growstride = (size_t)NSCR * wstride;  // scratch stride between consecutive g-points
  double* __restrict__ scr0 = W.scr + ((size_t)gabs * NSCR) * wstride + c;
  if (!cy.p) cy = Carry{scr0, wstride, (size_t)ncc, growstride};
  // aerosol / cloud optical properties of this band in layer l, and the McICA bits of the unit's g-points
  struct BandLayer {
    double ptaua, pomga, pasya, pclfr, ptauc, pomgc, pasyc;
    unsigned mbits;
  };
  auto band_layer = [&](int l) {
    BandLayer b{0., 1., 0., 0., 0., 1., 0., 0u};
    if (fl.iaer == 10) {
      const size_t oa = ((size_t)ib * nlay + l) * ncol + gc;
      b.ptaua = in.tauaer[oa]; b.pomga = in.ssaaer[oa]; b.pasya = in.asmaer[oa];
    } else if (fl.iaer == 6) {
      b.ptaua = W.aer[((size_t)(0 * 14 + ib) * nlay + l) * ncc + c];
      b.pomga = W.aer[((size_t)(1 * 14 + ib) * nlay + l) * ncc + c];
      b.pasya = W.aer[((size_t)(2 * 14 + ib) * nlay + l) * ncc + c];
    }
    if (cloudy_col) {
      b.pclfr = MC ? 1.0 : in.cldfr[(size_t)l * ncol + gc];
      b.ptauc = W.cld[((size_t)(0 * 14 + ib) * nlay + l) * ncc + c];
      b.pomgc = W.cld[((size_t)(1 * 14 + ib) * nlay + l) * ncc + c];
      b.pasyc = W.cld[((size_t)(2 * 14 + ib) * nlay + l) * ncc + c];
      if (MC) {
        const size_t ms = (size_t)W.mstride;
        const unsigned* mw = W.mask + ((size_t)l * 4) * ms + W.moff + c;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int g = gabs + u;
          b.mbits |= ((mw[(size_t)(g >> 5) * ms] >> (g & 31)) & 1u) << u;
        }
      }
    }
    return b;
  };
  auto layer = [&](const BandLayer& b, int u, double taug, double taur) {
    // spcvmc: the sub-column is either overcast with its band's optics or clear (mcica_subcol_gen_sw.f90:523-548)
    const bool on = !MC || ((b.mbits >> u) & 1u);
    return sw_layer_props(exp_tbl, bpade, prmu0, taug, taur, b.ptaua, b.pomga, b.pasya, cloudy_col, on ? b.pclfr : 0.,
                          on ? b.ptauc : 0., on ? b.pomgc : 1., on ? b.pasyc : 0.);
  };
  // ---- pass A: surface -> top.  layer optical properties, two-stream R/T, upward adding
  double rupc[U], rupdc[U], rup[U], rupd[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { rupc[u] = albdir; rupdc[u] = albdif; rup[u] = albdir; rupd[u] = albdif; }
  for (int l = 0; l < nlay; ++l) {
    const BandLayer b = band_layer(l);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* __restrict__ scr = scr0 + (size_t)u * growstride + (size_t)l * ncc;
      double* __restrict__ cr = cy.p + (size_t)u * cy.us + (size_t)l * cy.ls;
      if (SLAB && l + kSlabAhead < nlay) {
        prefetch_l2(scr + R_TAUG * wstride + (size_t)kSlabAhead * ncc);
        prefetch_l2(scr + R_TAUR * wstride + (size_t)kSlabAhead * ncc);
      }
      const double taug = sld(scr + R_TAUG * wstride), taur = sld(scr + R_TAUR * wstride);
      const SwLayer L = layer(b, u, taug, taur);
      {
        const double zreflect = frcp(1. - rupdc[u] * L.refdc);
        const double rn = L.refc + (L.tradc * ((L.trac - L.dbtc) * rupdc[u] + L.dbtc * rupc[u])) * zreflect;
        const double rdn = L.refdc + L.tradc * L.tradc * rupdc[u] * zreflect;
        rupc[u] = rn; rupdc[u] = rdn;
      }
      if (!RECOMPUTE) {
        cst(cr + 0 * cy.rs, L.refc); cst(cr + 1 * cy.rs, L.refdc); cst(cr + 2 * cy.rs, L.trac); cst(cr + 3 * cy.rs, L.tradc);
        cst(cr + 4 * cy.rs, L.dbtc);
      }
      cst(cr + 5 * cy.rs, rupc[u]); cst(cr + 6 * cy.rs, rupdc[u]);
      if (cloudy_col) {
        const double zreflect = frcp(1. - rupd[u] * L.refd);
        const double rn = L.ref + (L.trad * ((L.tra - L.dbt) * rupd[u] + L.dbt * rup[u])) * zreflect;
        const double rdn = L.refd + L.trad * L.trad * rupd[u] * zreflect;
        rup[u] = rn; rupd[u] = rdn;
        if (!RECOMPUTE) {
          cst(cr + 7 * cy.rs, L.ref); cst(cr + 8 * cy.rs, L.refd); cst(cr + 9 * cy.rs, L.tra); cst(cr + 10 * cy.rs, L.trad);
          cst(cr + 11 * cy.rs, L.dbt);
        }
        cst(cr + 12 * cy.rs, rup[u]); cst(cr + 13 * cy.rs, rupd[u]);
      }
    }
  }
  // ---- pass B: top -> surface.  downward adding and fluxes (vrtqdr.f90:146-169), weighted by the incoming flux
  double zinc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) zinc[u] = sol.adjflux[ib] * W.src[(size_t)(gabs + u) * ncc + c] * prmu0;
  double tdnc[U], rdndc[U], tdbtc[U], tdn[U], rdnd[U], tdbt[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { tdnc[u] = 1.; rdndc[u] = 0.; tdbtc[u] = 1.; tdn[u] = 1.; rdnd[u] = 0.; tdbt[u] = 1.; }
  for (int l = nlay - 1; l >= -1; --l) {
    // interface above layer l (level index l+1); l = -1 is the surface interface
    double sfu = 0., sfd = 0., scu = 0., scd = 0.;
    BandLayer b{0., 1., 0., 0., 0., 1., 0., 0u};
    if (RECOMPUTE && l >= 0) b = band_layer(l);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double refc = 0., refdc = 0., trac = 0., tradc = 0., dbtc = 0., prupc = albdir, prupdc = albdif;
      double ref = 0., refd = 0., tra = 0., trad = 0., dbt = 0., prup = albdir, prupd = albdif;
      if (l >= 0) {
        const double* __restrict__ scr = scr0 + (size_t)u * growstride + (size_t)l * ncc;
        const double* __restrict__ cr = cy.p + (size_t)u * cy.us + (size_t)l * cy.ls;
        if (RECOMPUTE) {
          const SwLayer L = layer(b, u, scr[R_TAUG * wstride], scr[R_TAUR * wstride]);
          refc = L.refc; refdc = L.refdc; trac = L.trac; tradc = L.tradc; dbtc = L.dbtc;
          ref = L.ref; refd = L.refd; tra = L.tra; trad = L.trad; dbt = L.dbt;
        } else {
          refc = cld(cr + 0 * cy.rs); refdc = cld(cr + 1 * cy.rs); trac = cld(cr + 2 * cy.rs); tradc = cld(cr + 3 * cy.rs);
          dbtc = cld(cr + 4 * cy.rs);
        }
        prupc = cld(cr + 5 * cy.rs); prupdc = cld(cr + 6 * cy.rs);
        if (cloudy_col) {
          if (!RECOMPUTE) {
            ref = cld(cr + 7 * cy.rs); refd = cld(cr + 8 * cy.rs); tra = cld(cr + 9 * cy.rs); trad = cld(cr + 10 * cy.rs);
            dbt = cld(cr + 11 * cy.rs);
          }
          prup = cld(cr + 12 * cy.rs); prupd = cld(cr + 13 * cy.rs);
        }
#if defined(__CUDA_ARCH__) && CB_SW_DISCARD
        // These rows were read for the last time.  A row of 32 lanes is two 128-byte lines: the lanes that start a line tell the
        // L2 that it need not write them back (r02: every carried row was being written to HBM even when its read had hit the L2).
        if (!SLAB && !RECOMPUTE && (threadIdx.x & 15) == 0) {
#pragma unroll
          for (int q = 0; q < 7; ++q) discard_l2(cr + q * cy.rs);
          if (cloudy_col) {
#pragma unroll
            for (int q = 7; q < 14; ++q) discard_l2(cr + q * cy.rs);
          }
        }
#endif
      }
      {
        const double zreflect = frcp(1. - rdndc[u] * prupdc);
        const double fu = (tdbtc[u] * prupc + (tdnc[u] - tdbtc[u]) * prupdc) * zreflect;
        const double fd = tdbtc[u] + (tdnc[u] - tdbtc[u] + tdbtc[u] * prupc * rdndc[u]) * zreflect;
        scu = scu + zinc[u] * fu;
        scd = scd + zinc[u] * fd;
        if (!cloudy_col) { sfu = sfu + zinc[u] * fu; sfd = sfd + zinc[u] * fd; }
        if (l >= 0) {
          if (l == nlay - 1) {  // jk = 1: ztdn(2) = ptra(1), prdnd(2) = prefd(1)
            tdnc[u] = trac;
            rdndc[u] = refdc;
          } else {
            const double zr = frcp(1. - refdc * rdndc[u]);
            const double t = tdbtc[u] * trac + (tradc * ((tdnc[u] - tdbtc[u]) + tdbtc[u] * refc * rdndc[u])) * zr;
            const double rd = refdc + tradc * tradc * rdndc[u] * zr;
            tdnc[u] = t; rdndc[u] = rd;
          }
          tdbtc[u] = dbtc * tdbtc[u];
        }
      }
      if (cloudy_col) {
        const double zreflect = frcp(1. - rdnd[u] * prupd);
        const double fu = (tdbt[u] * prup + (tdn[u] - tdbt[u]) * prupd) * zreflect;
        const double fd = tdbt[u] + (tdn[u] - tdbt[u] + tdbt[u] * prup * rdnd[u]) * zreflect;
        sfu = sfu + zinc[u] * fu;
        sfd = sfd + zinc[u] * fd;
        if (l >= 0) {
          if (l == nlay - 1) {
            tdn[u] = tra;
            rdnd[u] = refd;
          } else {
            const double zr = frcp(1. - refd * rdnd[u]);
            const double t = tdbt[u] * tra + (trad * ((tdn[u] - tdbt[u]) + tdbt[u] * ref * rdnd[u])) * zr;
            const double rd = refd + trad * trad * rdnd[u] * zr;
            tdn[u] = t; rdnd[u] = rd;
          }
          tdbt[u] = dbt * tdbt[u];
        }
      }
    }
    sink.put((size_t)(l + 1), cloudy_col, sfu, sfd, scu, scd);
  }
  sink.finish();
}

// ---------------------------------------------------------------------------------------------
// Column-tile form of the transfer (sw_engine.cu: k_sw_tile; host emulation: tests/emul/sw_emul.cpp) -- see lw_core.cuh for the
// idea.  Here the cell is where the arithmetic is (delta scaling, reftra_sw: two exponentials, a square root, a dozen divisions)
// and it does not depend on the adding recurrences at all:
//   sw_tile_cell    one (layer, column) cell of one g-point: the five layer properties of the clear-sky stream (ref, refd, tra,
//                   trad, dbt) and -- cloudy form -- of the total-sky stream; any thread of the block evaluates any cell, rows parked
//                   in SHARED memory;
//   sw_tile_sweeps  the upward adding sweep (rup, rupd per level, also in shared memory) and the downward sweep that turns them into
//                   the g-point's fluxes, added to per-level sums in shared memory -- one stream of one column, the reference's own
//                   formulas (vrtqdr.f90:107-169) in the reference's order.
// The 14 rows per (g-point, layer, column) that sw_transfer_unit carries through HBM (7.2 GB per 8192 x 60 launch against 0.49 GB
// of algorithmic bytes) never leave the SM.
constexpr int SR_REF = 0, SR_REFD = 1, SR_TRA = 2, SR_TRAD = 3, SR_DBT = 4;  // clear stream; the total stream's rows follow (+5)
constexpr int kSwTileRowsClear = 5, kSwTileRowsCloudy = 10;

// the two taumol rows of cell (l, c) for g-point gabs (streamed: read once)
CB_HD void sw_tile_cell_load(const In& in, const Work& W, int c, int l, int gabs, double& taug, double& taur) {
  const size_t wstride = (size_t)in.nlay * W.ncc;
  const double* __restrict__ scr = W.scr + (((size_t)gabs * NSCR) * in.nlay + l) * W.ncc + c;
  taug = ld_stream(scr + R_TAUG * wstride);
  taur = ld_stream(scr + R_TAUR * wstride);
}

// rows of cell (layer l, column c) for g-point gabs of band ib -> out[r * rs]: the band's aerosol and cloud properties of the layer
// and the g-point's sub-column bit exactly as sw_transfer_unit reads them
template <bool MC, bool CLOUDY>
CB_HD void sw_tile_cell(const Tables& T, const In& in, const Flags& fl, const Work& W, int c0, int c, int l, int ib, int gabs,
                        double prmu0, bool cloudy_col, double taug, double taur, double* __restrict__ out, size_t rs) {
  const int nlay = in.nlay, ncol = in.ncol, ncc = W.ncc;
  const size_t gc = (size_t)(c0 + c);
  const double* __restrict__ exp_tbl = T.base + T.exp_tbl;
  double ptaua = 0., pomga = 1., pasya = 0., pclfr = 0., ptauc = 0., pomgc = 1., pasyc = 0.;
  if (fl.iaer == 10) {
    const size_t oa = ((size_t)ib * nlay + l) * ncol + gc;
    ptaua = in.tauaer[oa]; pomga = in.ssaaer[oa]; pasya = in.asmaer[oa];
  } else if (fl.iaer == 6) {
    ptaua = W.aer[((size_t)(0 * 14 + ib) * nlay + l) * ncc + c];
    pomga = W.aer[((size_t)(1 * 14 + ib) * nlay + l) * ncc + c];
    pasya = W.aer[((size_t)(2 * 14 + ib) * nlay + l) * ncc + c];
  }
  const bool cl = CLOUDY && cloudy_col;
  if (cl) {
    bool on = true;
    if (MC) {  // spcvmc: the sub-column is either overcast with its band's optics or clear (mcica_subcol_gen_sw.f90:523-548)
      const unsigned w = W.mask[((size_t)l * 4 + (gabs >> 5)) * (size_t)W.mstride + W.moff + c];
      on = ((w >> (gabs & 31)) & 1u) != 0u;
    }
    if (on) {
      pclfr = MC ? 1.0 : in.cldfr[(size_t)l * ncol + gc];
      ptauc = W.cld[((size_t)(0 * 14 + ib) * nlay + l) * ncc + c];
      pomgc = W.cld[((size_t)(1 * 14 + ib) * nlay + l) * ncc + c];
      pasyc = W.cld[((size_t)(2 * 14 + ib) * nlay + l) * ncc + c];
    }
  }
  const SwLayer L = sw_layer_props(exp_tbl, T.bpade, prmu0, taug, taur, ptaua, pomga, pasya, cl, pclfr, ptauc, pomgc, pasyc);
  out[SR_REF * rs] = L.refc; out[SR_REFD * rs] = L.refdc; out[SR_TRA * rs] = L.trac; out[SR_TRAD * rs] = L.tradc;
  out[SR_DBT * rs] = L.dbtc;
  if (CLOUDY) {
    // a cloud-free column of a cloudy tile: its total-sky stream is its clear-sky one
    out[(5 + SR_REF) * rs] = cl ? L.ref : L.refc; out[(5 + SR_REFD) * rs] = cl ? L.refd : L.refdc;
    out[(5 + SR_TRA) * rs] = cl ? L.tra : L.trac; out[(5 + SR_TRAD) * rs] = cl ? L.trad : L.tradc;
    out[(5 + SR_DBT) * rs] = cl ? L.dbt : L.dbtc;
  }
}

// The two adding sweeps of ONE stream of one column over the parked rows of one g-point (spcvrt.f90:590-612, vrtqdr.f90:107-169):
// P = the stream's five rows of this column (row stride rs, layer stride ls); R = two rows (rup, rupd at the interface ABOVE layer l,
// row stride rrs, layer stride rls) written by the upward sweep and read by the downward one; acc_up / acc_dn = the stream's
// per-level flux sums of this column (level stride als).  zinc = adjflux * solar source * mu0 of the g-point.
//
// The reference's recurrences divide inside the serial chain -- rupd' = refd + trad^2 rupd / (1 - rupd refd) and its three
// siblings -- and a reciprocal plus its dependants is ~150 cycles of fp64 latency per level and sweep, which is what one warp per
// tile can least afford (r02 B200: 2.8 ms for 8192 x 60 with the chain written that way).  The maps are linear-fractional, so the
// sweeps carry numerators and a common denominator instead:
//   upward    rupd = p / q, rup = r / q:   q' = q - p refd;    p' = refd q' + trad^2 p;    r' = ref q' + trad ((tra - dbt) p + dbt r)
//   downward  rdnd = pd / qd, tdn = rd / qd, e = rd - tdbt qd (the diffuse part times qd):
//             qd' = qd - refd pd;   pd' = refd qd' + trad^2 pd;   rd' = tdbt tra qd' + trad (e + tdbt ref pd);   tdbt' = dbt tdbt
//             fu = (tdbt prup qd + e prupd) D,   fd = tdbt + (e + tdbt prup pd) D,   D = 1 / (qd - pd prupd)
// -- the same rational functions of the same layer properties (substitute and cancel q / q'), two dependent fma per level in the
// chain, the divisions (one per level and sweep) off it.  The denominators only shrink (by 1 - rupd refd <= 1 per level), so the
// triple is renormalised every kSwRenorm levels.  Differences from the reference's operation order: last bits (tests: <= 1e-12 of
// the unit form, <= 1e-10 of the oracle).
constexpr int kSwRenorm = 8;
CB_HD void sw_tile_sweeps(const double* __restrict__ P, size_t rs, size_t ls, int nlay, double albdir, double albdif, double zinc,
                          double* __restrict__ R, size_t rrs, size_t rls, double* __restrict__ acc_up, double* __restrict__ acc_dn,
                          size_t als) {
  // ---- surface -> top: reflectances of everything below each interface
  double p = albdif, q = 1., r = albdir;
#pragma unroll 4
  for (int l = 0; l < nlay; ++l) {
    const double* __restrict__ c = P + (size_t)l * ls;
    const double ref = c[SR_REF * rs], refd = c[SR_REFD * rs], tra = c[SR_TRA * rs], trad = c[SR_TRAD * rs], dbt = c[SR_DBT * rs];
    const double t2p = trad * trad * p, dir = trad * ((tra - dbt) * p + dbt * r);
    q = q - p * refd;
    p = refd * q + t2p;
    r = ref * q + dir;
    const double inv = frcp(q);
    const double rupd = p * inv, rup = r * inv;
    R[(size_t)l * rls] = rup; R[rrs + (size_t)l * rls] = rupd;
    if ((l & (kSwRenorm - 1)) == kSwRenorm - 1) { p = rupd; q = 1.; r = rup; }
  }
  // ---- top -> surface: downward adding and the fluxes at every interface
  double pd = 0., qd = 1., rd = 1., tdbt = 1.;
#pragma unroll 4
  for (int l = nlay - 1; l >= -1; --l) {
    double prup = albdir, prupd = albdif, ref = 0., refd = 0., tra = 0., trad = 0., dbt = 0.;
    if (l >= 0) {
      const double* __restrict__ c = P + (size_t)l * ls;
      ref = c[SR_REF * rs]; refd = c[SR_REFD * rs]; tra = c[SR_TRA * rs]; trad = c[SR_TRAD * rs]; dbt = c[SR_DBT * rs];
      prup = R[(size_t)l * rls]; prupd = R[rrs + (size_t)l * rls];
    }
    const double a_up = acc_up[(size_t)(l + 1) * als], a_dn = acc_dn[(size_t)(l + 1) * als];
    const double e = rd - tdbt * qd;
    const double D = frcp(qd - pd * prupd);
    const double fu = (tdbt * prup * qd + e * prupd) * D;
    const double fd = tdbt + (e + tdbt * prup * pd) * D;
    acc_up[(size_t)(l + 1) * als] = a_up + zinc * fu;
    acc_dn[(size_t)(l + 1) * als] = a_dn + zinc * fd;
    if (l >= 0) {
      const double t2p = trad * trad * pd, dif = trad * (e + tdbt * ref * pd), tt = tdbt * tra;
      qd = qd - refd * pd;
      pd = refd * qd + t2p;
      rd = tt * qd + dif;
      tdbt = dbt * tdbt;
      if (((nlay - 1 - l) & (kSwRenorm - 1)) == kSwRenorm - 1) {
        const double inv = frcp(qd);
        pd = pd * inv; rd = rd * inv; qd = 1.;
      }
    }
  }
}

// ---- scan form of the two sweeps (sw_engine.cu: k_sw_scan; host emulation: tests/emul/sw_emul.cpp) ----------------------------------
// sw_tile_sweeps is serial over the levels, so only one warp of a tile can run it, and its ~60 fp64 instructions per level on one
// scheduler are what bound k_sw_tile.  But the numerator / denominator maps above are LINEAR: a layer acts on the upward state
// (p, q, r) and on the downward state (pd, qd, b, e) -- b = tdbt, e = rd - tdbt qd -- as a structured matrix, matrices compose
// associatively, and the state at every interface is a prefix (from the surface) or suffix (from the top) product applied to a fixed
// start vector.  A group of kScanLanes lanes shares one column: every lane owns KL consecutive interfaces and the layers above them,
// multiplies its own layers' matrices, a log2(kScanLanes)-step shuffle scan gives each lane the product of everything below / above
// its block, and the lane then walks its own KL layers with the direct recurrences.  Columns are independent, so every warp of the
// block sweeps (32 / kScanLanes columns at a time): the sweeps are spread over all schedulers, nothing is parked between them (rup
// and rupd of a lane's interfaces stay in its registers until its downward walk needs them), and the flux sums of a lane's
// interfaces are register accumulators for the whole band group.
//   upward    [p'; q'] = A [p; q],  r' = c . [p; q] + a33 r        A = [[trad^2 - refd^2, refd], [-refd, 1]],
//                                                                   c = (trad (tra - dbt) - ref refd, ref),  a33 = trad dbt
//   downward  [pd'; qd'] = A [pd; qd],  b' = dbt b,  e' = trad e + b g . [pd; qd]      g = (trad ref - (tra - dbt) refd, tra - dbt)
// (expand sw_tile_sweeps' updates).  tools/adding_scan_study.py (r01) measured the reordering at <= 1.9e-15 of the serial sweep.
constexpr int kScanLanes = 8;
struct UpMap { double a11, a12, a21, a22, c1, c2, a33; };
struct DnMap { double a11, a12, a21, a22, beta, t, g1, g2; };
CB_HD UpMap up_identity() { return UpMap{1., 0., 0., 1., 0., 0., 1.}; }
CB_HD DnMap dn_identity() { return DnMap{1., 0., 0., 1., 1., 1., 0., 0.}; }
CB_HD UpMap up_layer(double ref, double refd, double tra, double trad, double dbt) {
  return UpMap{trad * trad - refd * refd, refd, -refd, 1., trad * (tra - dbt) - ref * refd, ref, trad * dbt};
}
CB_HD DnMap dn_layer(double ref, double refd, double tra, double trad, double dbt) {
  return DnMap{trad * trad - refd * refd, refd, -refd, 1., dbt, trad, trad * ref - (tra - dbt) * refd, tra - dbt};
}
// the map "m first, then n"
CB_HD UpMap up_compose(const UpMap& n, const UpMap& m) {
  UpMap o;
  o.a11 = n.a11 * m.a11 + n.a12 * m.a21; o.a12 = n.a11 * m.a12 + n.a12 * m.a22;
  o.a21 = n.a21 * m.a11 + n.a22 * m.a21; o.a22 = n.a21 * m.a12 + n.a22 * m.a22;
  o.c1 = n.c1 * m.a11 + n.c2 * m.a21 + n.a33 * m.c1;
  o.c2 = n.c1 * m.a12 + n.c2 * m.a22 + n.a33 * m.c2;
  o.a33 = n.a33 * m.a33;
  return o;
}
CB_HD DnMap dn_compose(const DnMap& n, const DnMap& m) {
  DnMap o;
  o.a11 = n.a11 * m.a11 + n.a12 * m.a21; o.a12 = n.a11 * m.a12 + n.a12 * m.a22;
  o.a21 = n.a21 * m.a11 + n.a22 * m.a21; o.a22 = n.a21 * m.a12 + n.a22 * m.a22;
  o.beta = n.beta * m.beta;
  o.t = n.t * m.t;
  o.g1 = n.t * m.g1 + m.beta * (n.g1 * m.a11 + n.g2 * m.a21);
  o.g2 = n.t * m.g2 + m.beta * (n.g1 * m.a12 + n.g2 * m.a22);
  return o;
}
// One lane's share of the two sweeps of one g-point.  i = lane of the column's group, KL = interfaces per lane (<= KLMAX); the lane
// owns interfaces i KL + k and the layers with the same indices, k = 0 .. KL-1.  row(r, k) reads property r of the lane's k-th layer.
// Part 1 (before the scans): the lane's own two products in one pass over its layers, lowest first -- the upward product takes each
// new layer on the left (applied after the layers below), the downward one on the right (applied before them).
template <int KLMAX, class RowFn>
CB_HD void scan_local(const RowFn& row, int nlay, int i, int KL, UpMap& um, DnMap& dm) {
  um = up_identity();
  dm = dn_identity();
#pragma unroll
  for (int k = 0; k < KLMAX; ++k) {
    if (k < KL && i * KL + k < nlay) {
      const double ref = row(SR_REF, k), refd = row(SR_REFD, k), tra = row(SR_TRA, k), trad = row(SR_TRAD, k), dbt = row(SR_DBT, k);
      um = up_compose(up_layer(ref, refd, tra, trad, dbt), um);
      dm = dn_compose(dm, dn_layer(ref, refd, tra, trad, dbt));
    }
  }
}
// Part 2 (after the scans): below = product of all layers under the lane's block, above = of all layers over it.  The lane walks
// its layers upward from the state at its lowest interface, keeping rup / rupd of its interfaces, then downward from the state at
// the top of its block, and adds zinc x (fu, fd) of its interfaces to acc_up / acc_dn[KLMAX] (k-th interface of the lane).
template <int KLMAX, class RowFn>
CB_HD void scan_walk(const RowFn& row, int nlay, int i, int KL, const UpMap& below, const DnMap& above, double albdir, double albdif,
                     double zinc, double* acc_up, double* acc_dn) {
  double rup[KLMAX], rupd[KLMAX];
  {
    // state at the surface (albdif, 1, albdir) carried through everything below the block
    double p = below.a11 * albdif + below.a12, q = below.a21 * albdif + below.a22;
    double r = below.c1 * albdif + below.c2 + below.a33 * albdir;
#pragma unroll
    for (int k = 0; k < KLMAX; ++k) {
      const int l = i * KL + k;  // interface l, then layer l above it
      if (k < KL && l <= nlay) {
        const double inv = frcp(q);
        rup[k] = r * inv; rupd[k] = p * inv;
        if (l < nlay) {
          const double ref = row(SR_REF, k), refd = row(SR_REFD, k), tra = row(SR_TRA, k), trad = row(SR_TRAD, k), dbt = row(SR_DBT, k);
          const double t2p = trad * trad * p, dir = trad * ((tra - dbt) * p + dbt * r);
          q = q - p * refd;
          p = refd * q + t2p;
          r = ref * q + dir;
        }
      }
    }
  }
  // state at the top of the atmosphere (pd, qd, b, e) = (0, 1, 1, 0) carried through everything above the block
  double pd = above.a12, qd = above.a22, b = above.beta, e = above.g2;
#pragma unroll
  for (int k = KLMAX - 1; k >= 0; --k) {
    const int l = i * KL + k;  // layer l (if any) first, then interface l below it
    if (k < KL && l <= nlay) {
      if (l < nlay) {
        const double ref = row(SR_REF, k), refd = row(SR_REFD, k), tra = row(SR_TRA, k), trad = row(SR_TRAD, k), dbt = row(SR_DBT, k);
        const double t2p = trad * trad * pd, g = trad * ref * pd;
        const double qn = qd - refd * pd;
        e = trad * e + b * ((tra - dbt) * qn + g);
        pd = refd * qn + t2p;
        qd = qn;
        b = dbt * b;
      }
      const double D = frcp(qd - pd * rupd[k]);
      const double fu = (b * rup[k] * qd + e * rupd[k]) * D;
      const double fd = b + (e + b * rup[k] * pd) * D;
      acc_up[k] = acc_up[k] + zinc * fu;
      acc_dn[k] = acc_dn[k] + zinc * fd;
    }
  }
}

// band groups of the tile form (see lw_core.cuh): 26 + 28 + 26 + 32 g-points
constexpr int kTileGroups = 4;
CB_HD void sw_tile_group_bands(int group, int& ib0, int& ib1) {  // 0-based bands [ib0, ib1)
  switch (group) {
    case 0: ib0 = 0; ib1 = 3; break;
    case 1: ib0 = 3; ib1 = 6; break;
    case 2: ib0 = 6; ib1 = 10; break;
    default: ib0 = 10; ib1 = 14; break;
  }
}
CB_HD int band_ngpt(int ib) { return (ib == 13 ? 112 : band_gstart(ib + 1)) - band_gstart(ib); }

struct Unit {
  int band, g0, u;  // band = 16..29
};
#ifndef CB_SW_UMAX
#define CB_SW_UMAX 1      // g-points per thread (= per unit) of the transfer kernel (1, 2 or 4).  r01 B200, 8192 x 60 clear sky, SW step:
                          // 2.57 ms at 2 (24 warps/SM) -> 2.50 ms at 1 (32 warps/SM; the transfer kernel alone 1.90 -> 1.66 ms, part of it
                          // returned by twice the partial-flux rows); cloudy 4.61 -> 4.28 ms, McICA 16384 x 72 12.45 -> 11.62 ms.
                          // Splitting a 2-g-point unit over half-warps (shuffle) or over warp pairs (shared memory + a barrier per
                          // level) to keep the rows of the 2-g-point form was measured too: 1.87 / 1.83 ms for the kernel -- not kept.
#endif
#ifndef CB_SW_TAU_UMAX
#define CB_SW_TAU_UMAX 4  // g-points per thread of the taumol kernel (2 or 4)
#endif
constexpr int kMaxUnits = 112;  // one g-point per unit at most
inline int build_units(Unit* out, int umax) {  // host only
  int n = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int b = 16; b <= 29; ++b) {
      const bool heavy = kNSPA[b - 16] == 9;
      if ((pass == 0) != heavy) continue;
      const int ng = kNG[b - 16];
      for (int g0 = 0; g0 < ng; g0 += umax) {
        out[n].band = b;
        out[n].g0 = g0;
        out[n].u = (ng - g0) >= umax ? umax : (ng - g0);
        ++n;
      }
    }
  return n;
}

// fixed-order reduction over the groups of units: deterministic (the unit list order; a group sums its units in that order), not
// the Fortran's band order (spcvrt.f90:614-618) -- the sum is the same to rounding
CB_HD void sw_reduce_level(const Work& W, int ngroups, int nlay, int c0, int c, int lev, int ncol, const Out& out) {
  const int ncc = W.ncc;
  const size_t pstride = (size_t)(nlay + 1) * ncc;
  const int q0 = W.anycld[c] != 0 ? 0 : 2;  // cloud-free column: only the clear-sky sums were stored
  double tot[4] = {0., 0., 0., 0.};
  for (int k = 0; k < ngroups; ++k) {
    const double* p = W.part + (size_t)k * 4 * pstride + (size_t)lev * ncc + c;
    for (int q = q0; q < 4; ++q) tot[q] = tot[q] + p[q * pstride];
  }
  if (q0 == 2) { tot[0] = tot[2]; tot[1] = tot[3]; }
  const size_t o = (size_t)lev * ncol + (c0 + c);
  out.uflx[o] = tot[0];
  out.dflx[o] = tot[1];
  out.uflxc[o] = tot[2];
  out.dflxc[o] = tot[3];
}
// heating rates (rad.nomcica.f90:797-807)
CB_HD void sw_heating(const Tables& T, const In& in, const Out& out, int gcol, int l) {
  const int ncol = in.ncol;
  const size_t o0 = (size_t)l * ncol + gcol, o1 = o0 + ncol;
  const double zdpgcp = T.heatfac / (in.plev[o0] - in.plev[o1]);
  const double n1 = out.dflx[o1] - out.uflx[o1], n0 = out.dflx[o0] - out.uflx[o0];
  const double c1 = out.dflxc[o1] - out.uflxc[o1], cc0 = out.dflxc[o0] - out.uflxc[o0];
  out.hr[o0] = (n1 - n0) * zdpgcp;
  out.hrc[o0] = (c1 - cc0) * zdpgcp;
}

}  // namespace sw
}  // namespace cb
