// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into, imported by
// or called from the product (climt_b200/).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it.
//
// CPU restatement (C++17, fp64, serial over columns exactly like the reference) of the
// AER RRTMG_LW path that climt wraps.  Every routine cites the reference file:line it
// follows (paths relative to /root/reference/climt/_lib/rrtmg_lw/).  The Fortran cannot be
// compiled in this image (no Fortran compiler), so parity is pinned against the reference's
// own golden outputs (tests/golden/, from tests/cached_component_output/TestRRTMGLongwave*).
//
//   rrtmg_lw_ini   rrtmg_lw_init.f90:28-175   (+ lwdatinit :178-281, lwcmbdat :284-363, cmbgb1..16 :366-2015)
//   inatm          rrtmg_lw_rad.nomcica.f90:572-900
//   cldprop        rrtmg_lw_cldprop.f90:31-276
//   setcoef        rrtmg_lw_setcoef.f90:31-415
//   taumol         rrtmg_lw_taumol.f90:31-3147
//   rtrn           rrtmg_lw_rtrn.f90:32-587
//   rrtmg_lw       rrtmg_lw_rad.nomcica.f90:80-569
//
// Numerical quirks kept on purpose: float32 table abscissa in the exp
// table (init.f90:114), rec_6 = 0.166667, float32 literals 1.e20 / 0.92.. / 3.55e-4 in
// taumol, real->integer truncation on assignment.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <algorithm>
#include <map>

#include "ftn.hpp"
#include "mcica_gen.hpp"

namespace orc {

static const int nbndlw = 16, mg = 16, ngptlw = 140, ntbl = 10000;
// lwdatinit (rrtmg_lw_init.f90:195-209)
static const double delwave_[16] = {340., 150., 130., 70., 120., 160., 100., 100.,
                                    210., 90., 320., 280., 170., 130., 220., 650.};
static const int nspa_[16] = {1, 1, 9, 9, 9, 1, 9, 1, 9, 1, 1, 9, 9, 1, 9, 9};
static const int nspb_[16] = {1, 1, 5, 5, 5, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0, 0};
// lwcmbdat (rrtmg_lw_init.f90:300-345)
static const int ngc_[16] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
static const int ngs_[16] = {10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140};
static const int ngm_[256] = {
    1, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 10,          // band 1
    1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 11, 11, 12, 12,     // band 2
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,    // band 3
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 14, 14,    // band 4
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,    // band 5
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,           // band 6
    1, 1, 2, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 11, 12, 12,      // band 7
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,           // band 8
    1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 11, 11, 12, 12,     // band 9
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6,           // band 10
    1, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 7, 8, 8, 8,           // band 11
    1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8,           // band 12
    1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4,           // band 13
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2,           // band 14
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2,           // band 15
    1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2};          // band 16
static const int ngn_[140] = {
    1, 1, 2, 2, 2, 2, 2, 2, 1, 1,                              // band 1
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2,                        // band 2
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,            // band 3
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3,                  // band 4
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,            // band 5
    2, 2, 2, 2, 2, 2, 2, 2,                                    // band 6
    2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2,                        // band 7
    2, 2, 2, 2, 2, 2, 2, 2,                                    // band 8
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2,                        // band 9
    2, 2, 2, 2, 4, 4,                                          // band 10
    1, 1, 2, 2, 2, 2, 3, 3,                                    // band 11
    1, 1, 1, 1, 2, 2, 4, 4,                                    // band 12
    3, 3, 4, 6,                                                // band 13
    8, 8,                                                      // band 14
    8, 8,                                                      // band 15
    4, 12};                                                    // band 16
static const double wt_[16] = {0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544,
                               0.1181945205, 0.1019300893, 0.0832767040, 0.0626720116,
                               0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086,
                               0.0022199750, 0.0014140010, 0.0005330000, 0.0000750000};

struct LwBand {
  int ng = 0;
  std::map<std::string, A2> t;  // reduced arrays: (lead, ng) with the Fortran leading dims flattened
  const A2& operator[](const char* k) const {
    auto it = t.find(k);
    if (it == t.end()) throw std::runtime_error(std::string("oracle: missing reduced table ") + k);
    return it->second;
  }
};

struct LwState {
  // rrlw_con
  double pi, grav, planck, boltz, clight, avogad, alosmt, gascon, sbcnst, secdy;
  double heatfac, fluxfac, oneminus;
  // rrlw_tbl
  std::vector<double> tau_tbl, exp_tbl, tfn_tbl;
  double bpade;
  // rrlw_wvn / rrlw_ref
  double rwgt[256];
  A2 totplnk, chi_mls, totplnkderiv;
  A1 totplk16, totplk16deriv, pref, preflog, tref;
  // rrlw_cld
  double abscld1, absliq0;
  A1 absice0;
  A2 absice1, absice2, absice3, absliq1;
  LwBand band[17];
  bool ready = false;
};

static LwState S;

static A2 load2(const Blob& b, const std::string& k) {
  const BlobEntry& e = b.get(k);
  int n1 = (int)e.shape[0], n2 = e.shape.size() > 1 ? (int)e.shape[1] : 1;
  A2 a(n1, n2);
  std::memcpy(a.d.data(), e.p, sizeof(double) * (size_t)e.count);
  return a;
}
static A1 load1(const Blob& b, const std::string& k) {
  const BlobEntry& e = b.get(k);
  A1 a((int)e.count);
  std::memcpy(a.d.data(), e.p, sizeof(double) * (size_t)e.count);
  return a;
}

// cmbgbN (rrtmg_lw_init.f90:366-2015): every `*o` array of band N is reduced over its last
// (g) axis; k-type arrays with the rwgt weights, Planck fractions with a plain sum.
static A2 reduce_g(const BlobEntry& e, int ibnd, bool weighted) {
  int64_t lead = e.count / 16;
  int ngc = ngc_[ibnd - 1];
  int ngs_prev = ibnd >= 2 ? ngs_[ibnd - 2] : 0;
  bool g_first = false;
  // fracrefao(no,9)/fracrefbo(no,5): g is the FIRST axis there (rrlw_kg03.f90:29); all others g last.
  A2 out;
  if (!weighted && e.shape.size() == 2 && e.shape[0] == 16) g_first = true;
  if (!weighted && e.shape.size() == 1) g_first = false;
  if (g_first) {
    int np = (int)e.shape[1];
    out = A2(ngc, np);
    for (int jp = 1; jp <= np; ++jp) {
      int iprsm = 0;
      for (int igc = 1; igc <= ngc; ++igc) {
        double sumf = 0.;
        for (int ipr = 1; ipr <= ngn_[ngs_prev + igc - 1]; ++ipr) {
          iprsm++;
          sumf = sumf + e.p[(iprsm - 1) + 16 * (size_t)(jp - 1)];
        }
        out(igc, jp) = sumf;
      }
    }
    return out;
  }
  out = A2((int)lead, ngc);
  for (int64_t l = 0; l < lead; ++l) {
    int iprsm = 0;
    for (int igc = 1; igc <= ngc; ++igc) {
      double sumk = 0.;
      for (int ipr = 1; ipr <= ngn_[ngs_prev + igc - 1]; ++ipr) {
        iprsm++;
        double v = e.p[l + lead * (size_t)(iprsm - 1)];
        sumk = weighted ? sumk + v * S.rwgt[(iprsm - 1) + 16 * (ibnd - 1)] : sumk + v;
      }
      out((int)l + 1, igc) = sumk;
    }
  }
  return out;
}

static void lw_ini(const Blob& b, double cpdair) {
  // lwdatinit :279
  S.heatfac = S.grav * S.secdy / (cpdair * 1.e2);
  // exp/tau/tfn lookup tables, rrtmg_lw_init.f90:97-123
  const double pade = 0.278, expeps = 1.e-20;
  S.tau_tbl.assign(ntbl + 1, 0.);
  S.exp_tbl.assign(ntbl + 1, 0.);
  S.tfn_tbl.assign(ntbl + 1, 0.);
  S.tau_tbl[0] = 0.0;
  S.tau_tbl[ntbl] = 1.e10;
  S.exp_tbl[0] = 1.0;
  S.exp_tbl[ntbl] = expeps;
  S.tfn_tbl[0] = 0.0;
  S.tfn_tbl[ntbl] = 1.0;
  S.bpade = 1.0 / pade;
  for (int itr = 1; itr <= ntbl - 1; ++itr) {
    double tfn = (double)((float)itr / (float)ntbl);  // default-real quotient (:114)
    S.tau_tbl[itr] = S.bpade * tfn / (1. - tfn);
    S.exp_tbl[itr] = std::exp(-S.tau_tbl[itr]);
    if (S.exp_tbl[itr] <= expeps) S.exp_tbl[itr] = expeps;
    if (S.tau_tbl[itr] < 0.06)
      S.tfn_tbl[itr] = S.tau_tbl[itr] / 6.;
    else
      S.tfn_tbl[itr] = 1. - 2. * ((1. / S.tau_tbl[itr]) - (S.exp_tbl[itr] / (1. - S.exp_tbl[itr])));
  }
  // reduction weights rwgt, :130-153
  {
    int igcsm = 0;
    double wtsm[17];
    for (int ibnd = 1; ibnd <= nbndlw; ++ibnd) {
      int iprsm = 0;
      if (ngc_[ibnd - 1] < mg) {
        for (int igc = 1; igc <= ngc_[ibnd - 1]; ++igc) {
          igcsm++;
          double wtsum = 0.;
          for (int ipr = 1; ipr <= ngn_[igcsm - 1]; ++ipr) {
            iprsm++;
            wtsum = wtsum + wt_[iprsm - 1];
          }
          wtsm[igc] = wtsum;
        }
        for (int ig = 1; ig <= 16; ++ig) {
          int ind = (ibnd - 1) * mg + ig;
          S.rwgt[ind - 1] = wt_[ig - 1] / wtsm[ngm_[ind - 1]];
        }
      } else {
        for (int ig = 1; ig <= 16; ++ig) {
          igcsm++;
          int ind = (ibnd - 1) * mg + ig;
          S.rwgt[ind - 1] = 1.0;
        }
      }
    }
  }
  // reference data
  S.totplnk = load2(b, "rrlw_wvn.totplnk");
  S.totplnkderiv = load2(b, "rrlw_wvn.totplnkderiv");
  S.totplk16 = load1(b, "rrlw_wvn.totplk16");
  S.totplk16deriv = load1(b, "rrlw_wvn.totplk16deriv");
  S.chi_mls = load2(b, "rrlw_ref.chi_mls");
  S.pref = load1(b, "rrlw_ref.pref");
  S.preflog = load1(b, "rrlw_ref.preflog");
  S.tref = load1(b, "rrlw_ref.tref");
  S.abscld1 = b.get("rrlw_cld.abscld1").p[0];
  S.absliq0 = b.get("rrlw_cld.absliq0").p[0];
  S.absice0 = load1(b, "rrlw_cld.absice0");
  S.absice1 = load2(b, "rrlw_cld.absice1");
  S.absice2 = load2(b, "rrlw_cld.absice2");
  S.absice3 = load2(b, "rrlw_cld.absice3");
  S.absliq1 = load2(b, "rrlw_cld.absliq1");
  // cmbgb1..16
  static const char* names[] = {"kao", "kbo", "selfrefo", "forrefo", "fracrefao", "fracrefbo",
                                "kao_mn2", "kbo_mn2", "kao_mn2o", "kbo_mn2o", "kao_mo3", "kbo_mo3",
                                "kao_mco2", "kbo_mco2", "kao_mco", "kao_mo2", "kbo_mo2",
                                "ccl4o", "cfc11adjo", "cfc12o", "cfc22adjo"};
  for (int ibnd = 1; ibnd <= 16; ++ibnd) {
    char mod[32];
    std::snprintf(mod, sizeof mod, "rrlw_kg%02d.", ibnd);
    S.band[ibnd].ng = ngc_[ibnd - 1];
    S.band[ibnd].t.clear();
    for (const char* nm : names) {
      std::string key = std::string(mod) + nm;
      if (!b.has(key)) continue;
      std::string s(nm), red;
      if (s.rfind("kao", 0) == 0)
        red = "ka" + s.substr(3);
      else if (s.rfind("kbo", 0) == 0)
        red = "kb" + s.substr(3);
      else
        red = s.substr(0, s.size() - 1);
      bool frac = s.rfind("fracref", 0) == 0;
      S.band[ibnd].t[red] = reduce_g(b.get(key), ibnd, !frac);
    }
  }
  S.ready = true;
}

// ---------------------------------------------------------------------------------------------
// per-column work arrays (1-based like the Fortran automatic arrays)
struct Col {
  int nlayers;
  A1 pavel, tavel, pz, tz, coldry, wbrodl;
  A2 wkl, wx;
  double tbound, pwvcm;
  double semiss[17];
  A2 taua;  // (nlay, 16)
  // clouds
  int inflag, iceflag, liqflag, ncbands;
  A1 cldfrac, ciwp, clwp, rei, rel;
  A2 tauc;      // (16, nlay)
  A2 taucloud;  // (nlay, 16)
  // setcoef out
  int laytrop;
  std::vector<int> jp, jt, jt1, indself, indfor, indminor;
  A2 planklay, planklev;  // planklev(0:nlay,16)
  double plankbnd[17], dplankbnd_dt[17];
  A1 colh2o, colco2, colo3, coln2o, colco, colch4, colo2, colbrd;
  A1 fac00, fac01, fac10, fac11;
  A1 rat_h2oco2, rat_h2oco2_1, rat_h2oo3, rat_h2oo3_1, rat_h2on2o, rat_h2on2o_1, rat_h2och4, rat_h2och4_1,
      rat_n2oco2, rat_n2oco2_1, rat_o3co2, rat_o3co2_1;
  A1 selffac, selffrac, forfac, forfrac, minorfrac, scaleminor, scaleminorn2;
  A2 fracs, taug, taut;  // (nlay, 140)
  explicit Col(int nlay) : nlayers(nlay) {
    int n = nlay + 1;
    pavel = A1(n); tavel = A1(n); pz = A1(n + 1, 0); tz = A1(n + 1, 0); coldry = A1(n); wbrodl = A1(n);
    wkl = A2(38, n); wx = A2(4, n); taua = A2(n, 16);
    cldfrac = A1(n); ciwp = A1(n); clwp = A1(n); rei = A1(n); rel = A1(n);
    tauc = A2(16, n); taucloud = A2(n, 16);
    jp.assign(n + 1, 0); jt = jp; jt1 = jp; indself = jp; indfor = jp; indminor = jp;
    planklay = A2(n, 16); planklev = A2(n + 1, 16, 0, 1);
    for (A1* a : {&colh2o, &colco2, &colo3, &coln2o, &colco, &colch4, &colo2, &colbrd, &fac00, &fac01, &fac10,
                  &fac11, &rat_h2oco2, &rat_h2oco2_1, &rat_h2oo3, &rat_h2oo3_1, &rat_h2on2o, &rat_h2on2o_1,
                  &rat_h2och4, &rat_h2och4_1, &rat_n2oco2, &rat_n2oco2_1, &rat_o3co2, &rat_o3co2_1, &selffac,
                  &selffrac, &forfac, &forfrac, &minorfrac, &scaleminor, &scaleminorn2})
      *a = A1(n);
    fracs = A2(n, 140); taug = A2(n, 140); taut = A2(n, 140);
  }
};

// Fortran real->integer assignment truncates toward zero.
static inline int f2i(double x) { return (int)x; }

// ---------------------------------------------------------------------------------------------
// setcoef — rrtmg_lw_setcoef.f90:31-415
static void setcoef(Col& c, int istart, int idrv) {
  const int nlayers = c.nlayers;
  const double stpfac = 296. / 1013.;
  int indbound = f2i(c.tbound - 159.);
  if (indbound < 1) indbound = 1; else if (indbound > 180) indbound = 180;
  double tbndfrac = c.tbound - 159. - (double)(float)indbound;
  int indlev0 = f2i(c.tz(0) - 159.);
  if (indlev0 < 1) indlev0 = 1; else if (indlev0 > 180) indlev0 = 180;
  double t0frac = c.tz(0) - 159. - (double)(float)indlev0;
  c.laytrop = 0;
  for (int lay = 1; lay <= nlayers; ++lay) {
    int indlay = f2i(c.tavel(lay) - 159.);
    if (indlay < 1) indlay = 1; else if (indlay > 180) indlay = 180;
    double tlayfrac = c.tavel(lay) - 159. - (double)(float)indlay;
    int indlev = f2i(c.tz(lay) - 159.);
    if (indlev < 1) indlev = 1; else if (indlev > 180) indlev = 180;
    double tlevfrac = c.tz(lay) - 159. - (double)(float)indlev;
    double dbdtlev, dbdtlay;
    for (int iband = 1; iband <= 15; ++iband) {
      if (lay == 1) {
        dbdtlev = S.totplnk(indbound + 1, iband) - S.totplnk(indbound, iband);
        c.plankbnd[iband] = c.semiss[iband] * (S.totplnk(indbound, iband) + tbndfrac * dbdtlev);
        dbdtlev = S.totplnk(indlev0 + 1, iband) - S.totplnk(indlev0, iband);
        c.planklev(0, iband) = S.totplnk(indlev0, iband) + t0frac * dbdtlev;
        if (idrv == 1) {
          dbdtlev = S.totplnkderiv(indbound + 1, iband) - S.totplnkderiv(indbound, iband);
          c.dplankbnd_dt[iband] = c.semiss[iband] * (S.totplnkderiv(indbound, iband) + tbndfrac * dbdtlev);
        }
      }
      dbdtlev = S.totplnk(indlev + 1, iband) - S.totplnk(indlev, iband);
      dbdtlay = S.totplnk(indlay + 1, iband) - S.totplnk(indlay, iband);
      c.planklay(lay, iband) = S.totplnk(indlay, iband) + tlayfrac * dbdtlay;
      c.planklev(lay, iband) = S.totplnk(indlev, iband) + tlevfrac * dbdtlev;
    }
    int iband = 16;
    if (istart == 16) {
      if (lay == 1) {
        dbdtlev = S.totplk16(indbound + 1) - S.totplk16(indbound);
        c.plankbnd[iband] = c.semiss[iband] * (S.totplk16(indbound) + tbndfrac * dbdtlev);
        if (idrv == 1) {
          dbdtlev = S.totplk16deriv(indbound + 1) - S.totplk16deriv(indbound);
          c.dplankbnd_dt[iband] = c.semiss[iband] * (S.totplk16deriv(indbound) + tbndfrac * dbdtlev);
        }
        dbdtlev = S.totplnk(indlev0 + 1, iband) - S.totplnk(indlev0, iband);
        c.planklev(0, iband) = S.totplk16(indlev0) + t0frac * dbdtlev;
      }
      dbdtlev = S.totplk16(indlev + 1) - S.totplk16(indlev);
      dbdtlay = S.totplk16(indlay + 1) - S.totplk16(indlay);
      c.planklay(lay, iband) = S.totplk16(indlay) + tlayfrac * dbdtlay;
      c.planklev(lay, iband) = S.totplk16(indlev) + tlevfrac * dbdtlev;
    } else {
      if (lay == 1) {
        dbdtlev = S.totplnk(indbound + 1, iband) - S.totplnk(indbound, iband);
        c.plankbnd[iband] = c.semiss[iband] * (S.totplnk(indbound, iband) + tbndfrac * dbdtlev);
        if (idrv == 1) {
          dbdtlev = S.totplnkderiv(indbound + 1, iband) - S.totplnkderiv(indbound, iband);
          c.dplankbnd_dt[iband] = c.semiss[iband] * (S.totplnkderiv(indbound, iband) + tbndfrac * dbdtlev);
        }
        dbdtlev = S.totplnk(indlev0 + 1, iband) - S.totplnk(indlev0, iband);
        c.planklev(0, iband) = S.totplnk(indlev0, iband) + t0frac * dbdtlev;
      }
      dbdtlev = S.totplnk(indlev + 1, iband) - S.totplnk(indlev, iband);
      dbdtlay = S.totplnk(indlay + 1, iband) - S.totplnk(indlay, iband);
      c.planklay(lay, iband) = S.totplnk(indlay, iband) + tlayfrac * dbdtlay;
      c.planklev(lay, iband) = S.totplnk(indlev, iband) + tlevfrac * dbdtlev;
    }
    // :257-285
    double plog = std::log(c.pavel(lay));
    c.jp[lay] = (int)(36. - 5 * (plog + 0.04));
    if (c.jp[lay] < 1) c.jp[lay] = 1; else if (c.jp[lay] > 58) c.jp[lay] = 58;
    int jp1 = c.jp[lay] + 1;
    double fp = 5. * (S.preflog(c.jp[lay]) - plog);
    c.jt[lay] = (int)(3. + (c.tavel(lay) - S.tref(c.jp[lay])) / 15.);
    if (c.jt[lay] < 1) c.jt[lay] = 1; else if (c.jt[lay] > 4) c.jt[lay] = 4;
    double ft = ((c.tavel(lay) - S.tref(c.jp[lay])) / 15.) - (double)(float)(c.jt[lay] - 3);
    c.jt1[lay] = (int)(3. + (c.tavel(lay) - S.tref(jp1)) / 15.);
    if (c.jt1[lay] < 1) c.jt1[lay] = 1; else if (c.jt1[lay] > 4) c.jt1[lay] = 4;
    double ft1 = ((c.tavel(lay) - S.tref(jp1)) / 15.) - (double)(float)(c.jt1[lay] - 3);
    double water = c.wkl(1, lay) / c.coldry(lay);
    double scalefac = c.pavel(lay) * stpfac / c.tavel(lay);
    double factor;
    const int jpl = c.jp[lay];
    if (!(plog <= 4.56)) {
      // lower atmosphere :294-350
      c.laytrop = c.laytrop + 1;
      c.forfac(lay) = scalefac / (1. + water);
      factor = (332.0 - c.tavel(lay)) / 36.0;
      c.indfor[lay] = std::min(2, std::max(1, (int)factor));
      c.forfrac(lay) = factor - (double)(float)c.indfor[lay];
      c.selffac(lay) = water * c.forfac(lay);
      factor = (c.tavel(lay) - 188.0) / 7.2;
      c.indself[lay] = std::min(9, std::max(1, (int)factor - 7));
      c.selffrac(lay) = factor - (double)(float)(c.indself[lay] + 7);
      c.scaleminor(lay) = c.pavel(lay) / c.tavel(lay);
      c.scaleminorn2(lay) = (c.pavel(lay) / c.tavel(lay)) * (c.wbrodl(lay) / (c.coldry(lay) + c.wkl(1, lay)));
      factor = (c.tavel(lay) - 180.8) / 7.2;
      c.indminor[lay] = std::min(18, std::max(1, (int)factor));
      c.minorfrac(lay) = factor - (double)(float)c.indminor[lay];
      c.rat_h2oco2(lay) = S.chi_mls(1, jpl) / S.chi_mls(2, jpl);
      c.rat_h2oco2_1(lay) = S.chi_mls(1, jpl + 1) / S.chi_mls(2, jpl + 1);
      c.rat_h2oo3(lay) = S.chi_mls(1, jpl) / S.chi_mls(3, jpl);
      c.rat_h2oo3_1(lay) = S.chi_mls(1, jpl + 1) / S.chi_mls(3, jpl + 1);
      c.rat_h2on2o(lay) = S.chi_mls(1, jpl) / S.chi_mls(4, jpl);
      c.rat_h2on2o_1(lay) = S.chi_mls(1, jpl + 1) / S.chi_mls(4, jpl + 1);
      c.rat_h2och4(lay) = S.chi_mls(1, jpl) / S.chi_mls(6, jpl);
      c.rat_h2och4_1(lay) = S.chi_mls(1, jpl + 1) / S.chi_mls(6, jpl + 1);
      c.rat_n2oco2(lay) = S.chi_mls(4, jpl) / S.chi_mls(2, jpl);
      c.rat_n2oco2_1(lay) = S.chi_mls(4, jpl + 1) / S.chi_mls(2, jpl + 1);
    } else {
      // upper atmosphere :353-398
      c.forfac(lay) = scalefac / (1. + water);
      factor = (c.tavel(lay) - 188.0) / 36.0;
      c.indfor[lay] = 3;
      c.forfrac(lay) = factor - 1.0;
      c.selffac(lay) = water * c.forfac(lay);
      c.scaleminor(lay) = c.pavel(lay) / c.tavel(lay);
      c.scaleminorn2(lay) = (c.pavel(lay) / c.tavel(lay)) * (c.wbrodl(lay) / (c.coldry(lay) + c.wkl(1, lay)));
      factor = (c.tavel(lay) - 180.8) / 7.2;
      c.indminor[lay] = std::min(18, std::max(1, (int)factor));
      c.minorfrac(lay) = factor - (double)(float)c.indminor[lay];
      c.rat_h2oco2(lay) = S.chi_mls(1, jpl) / S.chi_mls(2, jpl);
      c.rat_h2oco2_1(lay) = S.chi_mls(1, jpl + 1) / S.chi_mls(2, jpl + 1);
      c.rat_o3co2(lay) = S.chi_mls(3, jpl) / S.chi_mls(2, jpl);
      c.rat_o3co2_1(lay) = S.chi_mls(3, jpl + 1) / S.chi_mls(2, jpl + 1);
    }
    // common to both branches (:334-349 / :380-395)
    c.colh2o(lay) = 1.e-20 * c.wkl(1, lay);
    c.colco2(lay) = 1.e-20 * c.wkl(2, lay);
    c.colo3(lay) = 1.e-20 * c.wkl(3, lay);
    c.coln2o(lay) = 1.e-20 * c.wkl(4, lay);
    c.colco(lay) = 1.e-20 * c.wkl(5, lay);
    c.colch4(lay) = 1.e-20 * c.wkl(6, lay);
    c.colo2(lay) = 1.e-20 * c.wkl(7, lay);
    if (c.colco2(lay) == 0.) c.colco2(lay) = 1.e-32 * c.coldry(lay);
    if (c.colo3(lay) == 0.) c.colo3(lay) = 1.e-32 * c.coldry(lay);
    if (c.coln2o(lay) == 0.) c.coln2o(lay) = 1.e-32 * c.coldry(lay);
    if (c.colco(lay) == 0.) c.colco(lay) = 1.e-32 * c.coldry(lay);
    if (c.colch4(lay) == 0.) c.colch4(lay) = 1.e-32 * c.coldry(lay);
    c.colbrd(lay) = 1.e-20 * c.wbrodl(lay);
    // :401-412
    double compfp = 1. - fp;
    c.fac10(lay) = compfp * ft;
    c.fac00(lay) = compfp * (1. - ft);
    c.fac11(lay) = fp * ft1;
    c.fac01(lay) = fp * (1. - ft1);
    c.selffac(lay) = c.colh2o(lay) * c.selffac(lay);
    c.forfac(lay) = c.colh2o(lay) * c.forfac(lay);
  }
}

// ---------------------------------------------------------------------------------------------
// inatm — rrtmg_lw_rad.nomcica.f90:572-900 (arrays are (ncol, nlay) column-fastest)
struct LwIn {
  int ncol, nlay;
  const double *play, *plev, *tlay, *tlev, *tsfc, *h2o, *o3, *co2, *ch4, *n2o, *o2, *cfc11, *cfc12, *cfc22, *ccl4,
      *emis, *cldfr, *taucld, *cicewp, *cliqwp, *reice, *reliq, *tauaer;
};
#define IN2(a, ip, l) in.a[(size_t)((ip)-1) + (size_t)in.ncol * ((l)-1)]

static void inatm(const LwIn& in, int iplon, int icld, int iaer, int inflglw, int iceflglw, int liqflglw, Col& c) {
  const double amd = 28.9660, amw = 18.0160;
  const int nlayers = in.nlay;
  std::fill(c.wkl.d.begin(), c.wkl.d.end(), 0.);
  std::fill(c.wx.d.begin(), c.wx.d.end(), 0.);
  std::fill(c.cldfrac.d.begin(), c.cldfrac.d.end(), 0.);
  std::fill(c.tauc.d.begin(), c.tauc.d.end(), 0.);
  std::fill(c.ciwp.d.begin(), c.ciwp.d.end(), 0.);
  std::fill(c.clwp.d.begin(), c.clwp.d.end(), 0.);
  std::fill(c.rei.d.begin(), c.rei.d.end(), 0.);
  std::fill(c.rel.d.begin(), c.rel.d.end(), 0.);
  std::fill(c.taua.d.begin(), c.taua.d.end(), 0.);
  double amttl = 0.0, wvttl = 0.0;
  c.tbound = in.tsfc[iplon - 1];
  c.pz(0) = IN2(plev, iplon, 1);
  c.tz(0) = IN2(tlev, iplon, 1);
  for (int l = 1; l <= nlayers; ++l) {
    c.pavel(l) = IN2(play, iplon, l);
    c.tavel(l) = IN2(tlay, iplon, l);
    c.pz(l) = IN2(plev, iplon, l + 1);
    c.tz(l) = IN2(tlev, iplon, l + 1);
    c.wkl(1, l) = IN2(h2o, iplon, l);
    c.wkl(2, l) = IN2(co2, iplon, l);
    c.wkl(3, l) = IN2(o3, iplon, l);
    c.wkl(4, l) = IN2(n2o, iplon, l);
    c.wkl(6, l) = IN2(ch4, iplon, l);
    c.wkl(7, l) = IN2(o2, iplon, l);
    double amm = (1. - c.wkl(1, l)) * amd + c.wkl(1, l) * amw;
    c.coldry(l) = (c.pz(l - 1) - c.pz(l)) * 1.e3 * S.avogad / (1.e2 * S.grav * amm * (1. + c.wkl(1, l)));
  }
  for (int l = 1; l <= nlayers; ++l) {
    c.wx(1, l) = IN2(ccl4, iplon, l);
    c.wx(2, l) = IN2(cfc11, iplon, l);
    c.wx(3, l) = IN2(cfc12, iplon, l);
    c.wx(4, l) = IN2(cfc22, iplon, l);
  }
  for (int l = 1; l <= nlayers; ++l) {
    double summol = 0.0;
    for (int imol = 2; imol <= 7; ++imol) summol = summol + c.wkl(imol, l);
    c.wbrodl(l) = c.coldry(l) * (1. - summol);
    for (int imol = 1; imol <= 7; ++imol) c.wkl(imol, l) = c.coldry(l) * c.wkl(imol, l);
    amttl = amttl + c.coldry(l) + c.wkl(1, l);
    wvttl = wvttl + c.wkl(1, l);
    for (int ix = 1; ix <= 4; ++ix) c.wx(ix, l) = c.coldry(l) * c.wx(ix, l) * 1.e-20;  // ixindx(ix)=ix
  }
  double wvsh = (amw * wvttl) / (amd * amttl);
  c.pwvcm = wvsh * (1.e3 * c.pz(0)) / (1.e2 * S.grav);
  for (int n = 1; n <= nbndlw; ++n) c.semiss[n] = in.emis[(size_t)(iplon - 1) + (size_t)in.ncol * (n - 1)];
  if (iaer >= 1)
    for (int l = 1; l <= nlayers; ++l)
      for (int ib = 1; ib <= nbndlw; ++ib)
        c.taua(l, ib) = in.tauaer[(size_t)(iplon - 1) + (size_t)in.ncol * ((l - 1) + (size_t)in.nlay * (ib - 1))];
  if (icld >= 1) {
    c.inflag = inflglw;
    c.iceflag = iceflglw;
    c.liqflag = liqflglw;
    for (int l = 1; l <= nlayers; ++l) {
      c.cldfrac(l) = IN2(cldfr, iplon, l);
      c.ciwp(l) = IN2(cicewp, iplon, l);
      c.clwp(l) = IN2(cliqwp, iplon, l);
      c.rei(l) = IN2(reice, iplon, l);
      c.rel(l) = IN2(reliq, iplon, l);
      for (int n = 1; n <= nbndlw; ++n)  // taucld(nbndlw, ncol, nlay)
        c.tauc(n, l) = in.taucld[(size_t)(n - 1) + 16 * ((size_t)(iplon - 1) + (size_t)in.ncol * (l - 1))];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cldprop — rrtmg_lw_cldprop.f90:31-276.  Returns nonzero (with message) where Fortran would `stop`.
static const int icb_[3][16] = {{1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
                                {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5},
                                {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}};
static int cldprop(Col& c, std::string& err) {
  const double cldmin = 1.e-20;
  const int nlayers = c.nlayers;
  double abscoice[17] = {0}, abscoliq[17] = {0};
  int iceind = 0, liqind = 0;
  c.ncbands = 1;
  std::vector<double> tauctot(nlayers + 2, 0.);
  for (int lay = 1; lay <= nlayers; ++lay)
    for (int ib = 1; ib <= nbndlw; ++ib) {
      c.taucloud(lay, ib) = 0.0;
      tauctot[lay] = tauctot[lay] + c.tauc(ib, lay);
    }
  for (int lay = 1; lay <= nlayers; ++lay) {
    double cwp = c.ciwp(lay) + c.clwp(lay);
    if (c.cldfrac(lay) >= cldmin && (cwp >= cldmin || tauctot[lay] >= cldmin)) {
      if (c.inflag == 0) {
        c.ncbands = 16;
        for (int ib = 1; ib <= c.ncbands; ++ib) c.taucloud(lay, ib) = c.tauc(ib, lay);
      } else if (c.inflag == 1) {
        c.ncbands = 16;
        for (int ib = 1; ib <= c.ncbands; ++ib) c.taucloud(lay, ib) = S.abscld1 * cwp;
      } else if (c.inflag == 2) {
        double radice = c.rei(lay);
        if (c.ciwp(lay) == 0.0) {
          abscoice[1] = 0.0;
          iceind = 0;
        } else if (c.iceflag == 0) {
          if (radice < 10.0) { err = "ICE RADIUS TOO SMALL"; return 1; }
          abscoice[1] = S.absice0(1) + S.absice0(2) / radice;
          iceind = 0;
        } else if (c.iceflag == 1) {
          if (radice < 13.0 || radice > 130.) { err = "ICE RADIUS OUT OF BOUNDS"; return 1; }
          c.ncbands = 5;
          for (int ib = 1; ib <= c.ncbands; ++ib) abscoice[ib] = S.absice1(1, ib) + S.absice1(2, ib) / radice;
          iceind = 1;
        } else if (c.iceflag == 2) {
          if (radice < 5.0 || radice > 131.0) { err = "ICE RADIUS OUT OF BOUNDS"; return 1; }
          c.ncbands = 16;
          double factor = (radice - 2.) / 3.;
          int index = (int)factor;
          if (index == 43) index = 42;
          double fint = factor - (double)(float)index;
          for (int ib = 1; ib <= c.ncbands; ++ib)
            abscoice[ib] = S.absice2(index, ib) + fint * (S.absice2(index + 1, ib) - (S.absice2(index, ib)));
          iceind = 2;
        } else if (c.iceflag == 3) {
          if (radice < 5.0 || radice > 140.0) { err = "ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; return 1; }
          c.ncbands = 16;
          double factor = (radice - 2.) / 3.;
          int index = (int)factor;
          if (index == 46) index = 45;
          double fint = factor - (double)(float)index;
          for (int ib = 1; ib <= c.ncbands; ++ib)
            abscoice[ib] = S.absice3(index, ib) + fint * (S.absice3(index + 1, ib) - (S.absice3(index, ib)));
          iceind = 2;
        }
        if (c.clwp(lay) == 0.0) {
          abscoliq[1] = 0.0;
          liqind = 0;
          if (iceind == 1) iceind = 2;
        } else if (c.liqflag == 0) {
          abscoliq[1] = S.absliq0;
          liqind = 0;
          if (iceind == 1) iceind = 2;
        } else if (c.liqflag == 1) {
          double radliq = c.rel(lay);
          if (radliq < 2.5 || radliq > 60.) { err = "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS"; return 1; }
          int index = (int)(radliq - 1.5);
          if (index == 0) index = 1;
          if (index == 58) index = 57;
          double fint = radliq - 1.5 - (double)(float)index;
          c.ncbands = 16;
          for (int ib = 1; ib <= c.ncbands; ++ib)
            abscoliq[ib] = S.absliq1(index, ib) + fint * (S.absliq1(index + 1, ib) - (S.absliq1(index, ib)));
          liqind = 2;
        }
        for (int ib = 1; ib <= c.ncbands; ++ib)
          c.taucloud(lay, ib) =
              c.ciwp(lay) * abscoice[icb_[iceind][ib - 1]] + c.clwp(lay) * abscoliq[icb_[liqind][ib - 1]];
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// taumol — rrtmg_lw_taumol.f90:31-3147
namespace {
struct Bin {  // binary-species parameter (e.g. :550-555)
  double speccomb, specparm, fs;
  int js;
};
inline Bin binspec(double cola, double rat, double colb, double n) {
  Bin b;
  b.speccomb = cola + rat * colb;
  b.specparm = cola / b.speccomb;
  if (b.specparm >= S.oneminus) b.specparm = S.oneminus;
  double specmult = n * b.specparm;
  b.js = 1 + (int)specmult;
  b.fs = std::fmod(specmult, 1.0);
  return b;
}
// major-species stencil: weights and row offsets of absa/absb relative to ind (e.g. :586-660)
struct Sten {
  double w[6];
  int off[6];
  int n;
};
inline Sten stencil3(double specparm, double fs, double fa, double fb, int nsp) {
  Sten s;
  if (specparm < 0.125) {
    double p = fs - 1, p4 = (p * p) * (p * p), fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
    s.n = 6;
    s.w[0] = fk0 * fa; s.off[0] = 0;
    s.w[1] = fk1 * fa; s.off[1] = 1;
    s.w[2] = fk2 * fa; s.off[2] = 2;
    s.w[3] = fk0 * fb; s.off[3] = nsp;
    s.w[4] = fk1 * fb; s.off[4] = nsp + 1;
    s.w[5] = fk2 * fb; s.off[5] = nsp + 2;
  } else if (specparm > 0.875) {
    double p = -fs, p4 = (p * p) * (p * p), fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
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
  }
  return s;
}
inline Sten stencil2(double fs, double fa, double fb, int nsp) {  // upper atmosphere (e.g. :693-700)
  Sten s;
  s.n = 4;
  s.w[0] = (1. - fs) * fa; s.off[0] = 0;
  s.w[1] = fs * fa;        s.off[1] = 1;
  s.w[2] = (1. - fs) * fb; s.off[2] = nsp;
  s.w[3] = fs * fb;        s.off[3] = nsp + 1;
  return s;
}
inline double dot(const Sten& s, const A2& a, int ind, int ig) {
  double acc = s.w[0] * a(ind + s.off[0], ig);
  for (int k = 1; k < s.n; ++k) acc = acc + s.w[k] * a(ind + s.off[k], ig);
  return acc;
}
inline double lin(const A2& a, int i, int ig, double f) { return a(i, ig) + f * (a(i + 1, ig) - a(i, ig)); }
// minor gas with (js, indm) bilinear interpolation, table (n1, 19, ng) flattened (e.g. :627-631)
inline double minor2(const A2& a, int n1, int j, int indm, int ig, double fj, double fm) {
  auto K = [&](int jj, int ii) { return a(jj + n1 * (ii - 1), ig); };
  double m1 = K(j, indm) + fj * (K(j + 1, indm) - K(j, indm));
  double m2 = K(j, indm + 1) + fj * (K(j + 1, indm + 1) - K(j, indm + 1));
  return m1 + fm * (m2 - m1);
}
}  // namespace

static void taumol(Col& c) {
  const int nlayers = c.nlayers, laytrop = c.laytrop;
  const double oneminus = S.oneminus;
  (void)oneminus;
  auto& chi = S.chi_mls;
  auto simple4 = [&](const A2& a, int ind0, int ind1, int lay, int ig) {
    return c.fac00(lay) * a(ind0, ig) + c.fac10(lay) * a(ind0 + 1, ig) + c.fac01(lay) * a(ind1, ig) +
           c.fac11(lay) * a(ind1 + 1, ig);
  };
  auto tself = [&](const A2& selfref, int lay, int ig) {
    int inds = c.indself[lay];
    return c.selffac(lay) * (selfref(inds, ig) + c.selffrac(lay) * (selfref(inds + 1, ig) - selfref(inds, ig)));
  };
  auto tfor = [&](const A2& forref, int lay, int ig) {
    int indf = c.indfor[lay];
    return c.forfac(lay) * (forref(indf, ig) + c.forfrac(lay) * (forref(indf + 1, ig) - forref(indf, ig)));
  };
  auto ind0a = [&](int lay, int b) { return ((c.jp[lay] - 1) * 5 + (c.jt[lay] - 1)) * nspa_[b - 1]; };
  auto ind1a = [&](int lay, int b) { return (c.jp[lay] * 5 + (c.jt1[lay] - 1)) * nspa_[b - 1]; };
  auto ind0b = [&](int lay, int b) { return ((c.jp[lay] - 13) * 5 + (c.jt[lay] - 1)) * nspb_[b - 1]; };
  auto ind1b = [&](int lay, int b) { return ((c.jp[lay] - 12) * 5 + (c.jt1[lay] - 1)) * nspb_[b - 1]; };

  // ---- band 1: 10-350 cm-1 (low key - h2o; low minor - n2) (high key - h2o; high minor - n2) :280-376
  {
    const LwBand& B = S.band[1];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &ka_mn2 = B["ka_mn2"], &kb_mn2 = B["kb_mn2"], &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 1) + 1, ind1 = ind1a(lay, 1) + 1, indm = c.indminor[lay];
      double pp = c.pavel(lay);
      double corradj = 1.;
      if (pp < 250.) corradj = 1. - 0.15 * (250. - pp) / 154.4;
      double scalen2 = c.colbrd(lay) * c.scaleminorn2(lay);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double taun2 = scalen2 * (ka_mn2(indm, ig) + c.minorfrac(lay) * (ka_mn2(indm + 1, ig) - ka_mn2(indm, ig)));
        c.taug(lay, ig) = corradj * (c.colh2o(lay) * simple4(absa, ind0, ind1, lay, ig) + tauself + taufor + taun2);
        c.fracs(lay, ig) = fracrefa(ig, 1);
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 1) + 1, ind1 = ind1b(lay, 1) + 1, indm = c.indminor[lay];
      double pp = c.pavel(lay);
      double corradj = 1. - 0.15 * (pp / 95.6);
      double scalen2 = c.colbrd(lay) * c.scaleminorn2(lay);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double taufor = tfor(forref, lay, ig);
        double taun2 = scalen2 * (kb_mn2(indm, ig) + c.minorfrac(lay) * (kb_mn2(indm + 1, ig) - kb_mn2(indm, ig)));
        c.taug(lay, ig) = corradj * (c.colh2o(lay) * simple4(absb, ind0, ind1, lay, ig) + taufor + taun2);
        c.fracs(lay, ig) = fracrefb(ig, 1);
      }
    }
  }
  // ---- band 2: 350-500 (low key - h2o; high key - h2o) :379-476
  {
    const LwBand& B = S.band[2];
    const int ngs1 = 10;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 2) + 1, ind1 = ind1a(lay, 2) + 1;
      double pp = c.pavel(lay);
      double corradj = 1. - .05 * (pp - 100.) / 900.;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        c.taug(lay, ngs1 + ig) = corradj * (c.colh2o(lay) * simple4(absa, ind0, ind1, lay, ig) + tauself + taufor);
        c.fracs(lay, ngs1 + ig) = fracrefa(ig, 1);
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 2) + 1, ind1 = ind1b(lay, 2) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double taufor = tfor(forref, lay, ig);
        c.taug(lay, ngs1 + ig) = c.colh2o(lay) * simple4(absb, ind0, ind1, lay, ig) + taufor;
        c.fracs(lay, ngs1 + ig) = fracrefb(ig, 1);
      }
    }
  }
  // ---- band 3: 500-630 (low key - h2o,co2; low minor - n2o) (high key - h2o,co2; high minor - n2o) :479-763
  {
    const LwBand& B = S.band[3];
    const int ngs2 = 22;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &ka_mn2o = B["ka_mn2o"], &kb_mn2o = B["kb_mn2o"], &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    double refrat_planck_a = chi(1, 9) / chi(2, 9), refrat_planck_b = chi(1, 13) / chi(2, 13);
    double refrat_m_a = chi(1, 3) / chi(2, 3), refrat_m_b = chi(1, 13) / chi(2, 13);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2oco2(lay), c.colco2(lay), 8.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2oco2_1(lay), c.colco2(lay), 8.);
      Bin sm = binspec(c.colh2o(lay), refrat_m_a, c.colco2(lay), 8.);
      int jmn2o = sm.js;
      double fmn2o = sm.fs;
      double chi_n2o = c.coln2o(lay) / c.coldry(lay);
      double ratn2o = 1.e20 * chi_n2o / chi(4, c.jp[lay] + 1);
      double adjcoln2o;
      if (ratn2o > 1.5) {
        double adjfac = 0.5 + std::pow(ratn2o - 0.5, 0.65);
        adjcoln2o = adjfac * chi(4, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcoln2o = c.coln2o(lay);
      Bin sp = binspec(c.colh2o(lay), refrat_planck_a, c.colco2(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 3) + s0.js, ind1 = ind1a(lay, 3) + s1.js, indm = c.indminor[lay];
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double absn2o = minor2(ka_mn2o, 9, jmn2o, indm, ig, fmn2o, c.minorfrac(lay));
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs2 + ig) = tau_major + tau_major1 + tauself + taufor + adjcoln2o * absn2o;
        c.fracs(lay, ngs2 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2oco2(lay), c.colco2(lay), 4.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2oco2_1(lay), c.colco2(lay), 4.);
      Sten t0 = stencil2(s0.fs, c.fac00(lay), c.fac10(lay), 5);
      Sten t1 = stencil2(s1.fs, c.fac01(lay), c.fac11(lay), 5);
      Bin sm = binspec(c.colh2o(lay), refrat_m_b, c.colco2(lay), 4.);
      int jmn2o = sm.js;
      double fmn2o = sm.fs;
      double chi_n2o = c.coln2o(lay) / c.coldry(lay);
      double ratn2o = (double)1.e20f * chi_n2o / chi(4, c.jp[lay] + 1);  // default-real literal 1.e20 (:715)
      double adjcoln2o;
      if (ratn2o > 1.5) {
        double adjfac = 0.5 + std::pow(ratn2o - 0.5, 0.65);
        adjcoln2o = adjfac * chi(4, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcoln2o = c.coln2o(lay);
      Bin sp = binspec(c.colh2o(lay), refrat_planck_b, c.colco2(lay), 4.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0b(lay, 3) + s0.js, ind1 = ind1b(lay, 3) + s1.js, indm = c.indminor[lay];
      for (int ig = 1; ig <= B.ng; ++ig) {
        double taufor = tfor(forref, lay, ig);
        double absn2o = minor2(kb_mn2o, 5, jmn2o, indm, ig, fmn2o, c.minorfrac(lay));
        c.taug(lay, ngs2 + ig) =
            s0.speccomb * dot(t0, absb, ind0, ig) + s1.speccomb * dot(t1, absb, ind1, ig) + taufor + adjcoln2o * absn2o;
        c.fracs(lay, ngs2 + ig) = fracrefb(ig, jpl) + fpl * (fracrefb(ig, jpl + 1) - fracrefb(ig, jpl));
      }
    }
  }
  // ---- band 4: 630-700 (low key - h2o,co2; high key - o3,co2) :766-1018
  {
    const LwBand& B = S.band[4];
    const int ngs3 = 38;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    double refrat_planck_a = chi(1, 11) / chi(2, 11), refrat_planck_b = chi(3, 13) / chi(2, 13);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2oco2(lay), c.colco2(lay), 8.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2oco2_1(lay), c.colco2(lay), 8.);
      Bin sp = binspec(c.colh2o(lay), refrat_planck_a, c.colco2(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 4) + s0.js, ind1 = ind1a(lay, 4) + s1.js;
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs3 + ig) = tau_major + tau_major1 + tauself + taufor;
        c.fracs(lay, ngs3 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      Bin s0 = binspec(c.colo3(lay), c.rat_o3co2(lay), c.colco2(lay), 4.);
      Bin s1 = binspec(c.colo3(lay), c.rat_o3co2_1(lay), c.colco2(lay), 4.);
      Sten t0 = stencil2(s0.fs, c.fac00(lay), c.fac10(lay), 5);
      Sten t1 = stencil2(s1.fs, c.fac01(lay), c.fac11(lay), 5);
      Bin sp = binspec(c.colo3(lay), refrat_planck_b, c.colco2(lay), 4.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0b(lay, 4) + s0.js, ind1 = ind1b(lay, 4) + s1.js;
      for (int ig = 1; ig <= B.ng; ++ig) {
        c.taug(lay, ngs3 + ig) = s0.speccomb * dot(t0, absb, ind0, ig) + s1.speccomb * dot(t1, absb, ind1, ig);
        c.fracs(lay, ngs3 + ig) = fracrefb(ig, jpl) + fpl * (fracrefb(ig, jpl + 1) - fracrefb(ig, jpl));
      }
      // empirical rescaling, default-real literals (:1009-1015)
      c.taug(lay, ngs3 + 8) = c.taug(lay, ngs3 + 8) * (double)0.92f;
      c.taug(lay, ngs3 + 9) = c.taug(lay, ngs3 + 9) * (double)0.88f;
      c.taug(lay, ngs3 + 10) = c.taug(lay, ngs3 + 10) * (double)1.07f;
      c.taug(lay, ngs3 + 11) = c.taug(lay, ngs3 + 11) * (double)1.1f;
      c.taug(lay, ngs3 + 12) = c.taug(lay, ngs3 + 12) * (double)0.99f;
      c.taug(lay, ngs3 + 13) = c.taug(lay, ngs3 + 13) * (double)0.88f;
      c.taug(lay, ngs3 + 14) = c.taug(lay, ngs3 + 14) * (double)0.943f;
    }
  }
  // ---- band 5: 700-820 (low key - h2o,co2; low minor - o3, ccl4) (high key - o3,co2) :1021-1300
  {
    const LwBand& B = S.band[5];
    const int ngs4 = 52;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &ka_mo3 = B["ka_mo3"], &ccl4 = B["ccl4"], &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    double refrat_planck_a = chi(1, 5) / chi(2, 5), refrat_planck_b = chi(3, 43) / chi(2, 43);
    double refrat_m_a = chi(1, 7) / chi(2, 7);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2oco2(lay), c.colco2(lay), 8.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2oco2_1(lay), c.colco2(lay), 8.);
      Bin sm = binspec(c.colh2o(lay), refrat_m_a, c.colco2(lay), 8.);
      int jmo3 = sm.js;
      double fmo3 = sm.fs;
      Bin sp = binspec(c.colh2o(lay), refrat_planck_a, c.colco2(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 5) + s0.js, ind1 = ind1a(lay, 5) + s1.js, indm = c.indminor[lay];
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double abso3 = minor2(ka_mo3, 9, jmo3, indm, ig, fmo3, c.minorfrac(lay));
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs4 + ig) =
            tau_major + tau_major1 + tauself + taufor + abso3 * c.colo3(lay) + c.wx(1, lay) * ccl4(1, ig);
        c.fracs(lay, ngs4 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      Bin s0 = binspec(c.colo3(lay), c.rat_o3co2(lay), c.colco2(lay), 4.);
      Bin s1 = binspec(c.colo3(lay), c.rat_o3co2_1(lay), c.colco2(lay), 4.);
      Sten t0 = stencil2(s0.fs, c.fac00(lay), c.fac10(lay), 5);
      Sten t1 = stencil2(s1.fs, c.fac01(lay), c.fac11(lay), 5);
      Bin sp = binspec(c.colo3(lay), refrat_planck_b, c.colco2(lay), 4.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0b(lay, 5) + s0.js, ind1 = ind1b(lay, 5) + s1.js;
      for (int ig = 1; ig <= B.ng; ++ig) {
        c.taug(lay, ngs4 + ig) = s0.speccomb * dot(t0, absb, ind0, ig) + s1.speccomb * dot(t1, absb, ind1, ig) +
                                 c.wx(1, lay) * ccl4(1, ig);
        c.fracs(lay, ngs4 + ig) = fracrefb(ig, jpl) + fpl * (fracrefb(ig, jpl + 1) - fracrefb(ig, jpl));
      }
    }
  }
  // ---- band 6: 820-980 (low key - h2o; low minor - co2) (high: cfc11, cfc12 only) :1303-1391
  {
    const LwBand& B = S.band[6];
    const int ngs5 = 68;
    const A2 &absa = B["ka"], &selfref = B["selfref"], &forref = B["forref"], &ka_mco2 = B["ka_mco2"];
    const A2 &cfc11adj = B["cfc11adj"], &cfc12 = B["cfc12"], &fracrefa = B["fracrefa"];
    for (int lay = 1; lay <= laytrop; ++lay) {
      double chi_co2 = c.colco2(lay) / (c.coldry(lay));
      double ratco2 = 1.e20 * chi_co2 / chi(2, c.jp[lay] + 1);
      double adjcolco2;
      if (ratco2 > 3.0) {
        double adjfac = 2.0 + std::pow(ratco2 - 2.0, 0.77);
        adjcolco2 = adjfac * chi(2, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcolco2 = c.colco2(lay);
      int ind0 = ind0a(lay, 6) + 1, ind1 = ind1a(lay, 6) + 1, indm = c.indminor[lay];
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double absco2 = (ka_mco2(indm, ig) + c.minorfrac(lay) * (ka_mco2(indm + 1, ig) - ka_mco2(indm, ig)));
        c.taug(lay, ngs5 + ig) = c.colh2o(lay) * simple4(absa, ind0, ind1, lay, ig) + tauself + taufor +
                                 adjcolco2 * absco2 + c.wx(2, lay) * cfc11adj(1, ig) + c.wx(3, lay) * cfc12(1, ig);
        c.fracs(lay, ngs5 + ig) = fracrefa(ig, 1);
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay)
      for (int ig = 1; ig <= B.ng; ++ig) {
        c.taug(lay, ngs5 + ig) = 0.0 + c.wx(2, lay) * cfc11adj(1, ig) + c.wx(3, lay) * cfc12(1, ig);
        c.fracs(lay, ngs5 + ig) = fracrefa(ig, 1);
      }
  }
  // ---- band 7: 980-1080 (low key - h2o,o3; low minor - co2) (high key - o3; high minor - co2) :1394-1653
  {
    const LwBand& B = S.band[7];
    const int ngs6 = 76;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &ka_mco2 = B["ka_mco2"], &kb_mco2 = B["kb_mco2"], &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    double refrat_planck_a = chi(1, 3) / chi(3, 3), refrat_m_a = chi(1, 3) / chi(3, 3);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2oo3(lay), c.colo3(lay), 8.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2oo3_1(lay), c.colo3(lay), 8.);
      Bin sm = binspec(c.colh2o(lay), refrat_m_a, c.colo3(lay), 8.);
      int jmco2 = sm.js;
      double fmco2 = sm.fs;
      double chi_co2 = c.colco2(lay) / (c.coldry(lay));
      double ratco2 = (double)1.e20f * chi_co2 / chi(2, c.jp[lay] + 1);  // default-real 1.e20 (:1462)
      double adjcolco2;
      if (ratco2 > 3.0) {
        double adjfac = 3.0 + std::pow(ratco2 - 3.0, 0.79);
        adjcolco2 = adjfac * chi(2, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcolco2 = c.colco2(lay);
      Bin sp = binspec(c.colh2o(lay), refrat_planck_a, c.colo3(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 7) + s0.js, ind1 = ind1a(lay, 7) + s1.js, indm = c.indminor[lay];
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double absco2 = minor2(ka_mco2, 9, jmco2, indm, ig, fmco2, c.minorfrac(lay));
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs6 + ig) = tau_major + tau_major1 + tauself + taufor + adjcolco2 * absco2;
        c.fracs(lay, ngs6 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      double chi_co2 = c.colco2(lay) / (c.coldry(lay));
      double ratco2 = (double)1.e20f * chi_co2 / chi(2, c.jp[lay] + 1);  // default-real 1.e20 (:1618)
      double adjcolco2;
      if (ratco2 > 3.0) {
        double adjfac = 2.0 + std::pow(ratco2 - 2.0, 0.79);
        adjcolco2 = adjfac * chi(2, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcolco2 = c.colco2(lay);
      int ind0 = ind0b(lay, 7) + 1, ind1 = ind1b(lay, 7) + 1, indm = c.indminor[lay];
      for (int ig = 1; ig <= B.ng; ++ig) {
        double absco2 = kb_mco2(indm, ig) + c.minorfrac(lay) * (kb_mco2(indm + 1, ig) - kb_mco2(indm, ig));
        c.taug(lay, ngs6 + ig) = c.colo3(lay) * simple4(absb, ind0, ind1, lay, ig) + adjcolco2 * absco2;
        c.fracs(lay, ngs6 + ig) = fracrefb(ig, 1);
      }
      // :1642-1650 (kind=rb literals here)
      c.taug(lay, ngs6 + 6) = c.taug(lay, ngs6 + 6) * 0.92;
      c.taug(lay, ngs6 + 7) = c.taug(lay, ngs6 + 7) * 0.88;
      c.taug(lay, ngs6 + 8) = c.taug(lay, ngs6 + 8) * 1.07;
      c.taug(lay, ngs6 + 9) = c.taug(lay, ngs6 + 9) * 1.1;
      c.taug(lay, ngs6 + 10) = c.taug(lay, ngs6 + 10) * 0.99;
      c.taug(lay, ngs6 + 11) = c.taug(lay, ngs6 + 11) * 0.855;
    }
  }
  // ---- band 8: 1080-1180 (low key - h2o; low minor - co2,o3,n2o) (high key - o3; high minor - co2, n2o) :1656-1791
  {
    const LwBand& B = S.band[8];
    const int ngs7 = 88;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &ka_mco2 = B["ka_mco2"], &ka_mn2o = B["ka_mn2o"], &ka_mo3 = B["ka_mo3"], &kb_mco2 = B["kb_mco2"],
             &kb_mn2o = B["kb_mn2o"], &cfc12 = B["cfc12"], &cfc22adj = B["cfc22adj"], &fracrefa = B["fracrefa"],
             &fracrefb = B["fracrefb"];
    for (int lay = 1; lay <= laytrop; ++lay) {
      double chi_co2 = c.colco2(lay) / (c.coldry(lay));
      double ratco2 = 1.e20 * chi_co2 / chi(2, c.jp[lay] + 1);
      double adjcolco2;
      if (ratco2 > 3.0) {
        double adjfac = 2.0 + std::pow(ratco2 - 2.0, 0.65);
        adjcolco2 = adjfac * chi(2, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcolco2 = c.colco2(lay);
      int ind0 = ind0a(lay, 8) + 1, ind1 = ind1a(lay, 8) + 1, indm = c.indminor[lay];
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double absco2 = (ka_mco2(indm, ig) + c.minorfrac(lay) * (ka_mco2(indm + 1, ig) - ka_mco2(indm, ig)));
        double abso3 = (ka_mo3(indm, ig) + c.minorfrac(lay) * (ka_mo3(indm + 1, ig) - ka_mo3(indm, ig)));
        double absn2o = (ka_mn2o(indm, ig) + c.minorfrac(lay) * (ka_mn2o(indm + 1, ig) - ka_mn2o(indm, ig)));
        c.taug(lay, ngs7 + ig) = c.colh2o(lay) * simple4(absa, ind0, ind1, lay, ig) + tauself + taufor +
                                 adjcolco2 * absco2 + c.colo3(lay) * abso3 + c.coln2o(lay) * absn2o +
                                 c.wx(3, lay) * cfc12(1, ig) + c.wx(4, lay) * cfc22adj(1, ig);
        c.fracs(lay, ngs7 + ig) = fracrefa(ig, 1);
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      double chi_co2 = c.colco2(lay) / c.coldry(lay);
      double ratco2 = 1.e20 * chi_co2 / chi(2, c.jp[lay] + 1);
      double adjcolco2;
      if (ratco2 > 3.0) {
        double adjfac = 2.0 + std::pow(ratco2 - 2.0, 0.65);
        adjcolco2 = adjfac * chi(2, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcolco2 = c.colco2(lay);
      int ind0 = ind0b(lay, 8) + 1, ind1 = ind1b(lay, 8) + 1, indm = c.indminor[lay];
      for (int ig = 1; ig <= B.ng; ++ig) {
        double absco2 = (kb_mco2(indm, ig) + c.minorfrac(lay) * (kb_mco2(indm + 1, ig) - kb_mco2(indm, ig)));
        double absn2o = (kb_mn2o(indm, ig) + c.minorfrac(lay) * (kb_mn2o(indm + 1, ig) - kb_mn2o(indm, ig)));
        c.taug(lay, ngs7 + ig) = c.colo3(lay) * simple4(absb, ind0, ind1, lay, ig) + adjcolco2 * absco2 +
                                 c.coln2o(lay) * absn2o + c.wx(3, lay) * cfc12(1, ig) + c.wx(4, lay) * cfc22adj(1, ig);
        c.fracs(lay, ngs7 + ig) = fracrefb(ig, 1);
      }
    }
  }
  // ---- band 9: 1180-1390 (low key - h2o,ch4; low minor - n2o) (high key - ch4; high minor - n2o) :1794-2040
  {
    const LwBand& B = S.band[9];
    const int ngs8 = 96;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &ka_mn2o = B["ka_mn2o"], &kb_mn2o = B["kb_mn2o"], &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    double refrat_planck_a = chi(1, 9) / chi(6, 9), refrat_m_a = chi(1, 3) / chi(6, 3);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2och4(lay), c.colch4(lay), 8.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2och4_1(lay), c.colch4(lay), 8.);
      Bin sm = binspec(c.colh2o(lay), refrat_m_a, c.colch4(lay), 8.);
      int jmn2o = sm.js;
      double fmn2o = sm.fs;
      double chi_n2o = c.coln2o(lay) / (c.coldry(lay));
      double ratn2o = 1.e20 * chi_n2o / chi(4, c.jp[lay] + 1);
      double adjcoln2o;
      if (ratn2o > 1.5) {
        double adjfac = 0.5 + std::pow(ratn2o - 0.5, 0.65);
        adjcoln2o = adjfac * chi(4, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcoln2o = c.coln2o(lay);
      Bin sp = binspec(c.colh2o(lay), refrat_planck_a, c.colch4(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 9) + s0.js, ind1 = ind1a(lay, 9) + s1.js, indm = c.indminor[lay];
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double absn2o = minor2(ka_mn2o, 9, jmn2o, indm, ig, fmn2o, c.minorfrac(lay));
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs8 + ig) = tau_major + tau_major1 + tauself + taufor + adjcoln2o * absn2o;
        c.fracs(lay, ngs8 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      double chi_n2o = c.coln2o(lay) / (c.coldry(lay));
      double ratn2o = 1.e20 * chi_n2o / chi(4, c.jp[lay] + 1);
      double adjcoln2o;
      if (ratn2o > 1.5) {
        double adjfac = 0.5 + std::pow(ratn2o - 0.5, 0.65);
        adjcoln2o = adjfac * chi(4, c.jp[lay] + 1) * c.coldry(lay) * 1.e-20;
      } else
        adjcoln2o = c.coln2o(lay);
      int ind0 = ind0b(lay, 9) + 1, ind1 = ind1b(lay, 9) + 1, indm = c.indminor[lay];
      for (int ig = 1; ig <= B.ng; ++ig) {
        double absn2o = kb_mn2o(indm, ig) + c.minorfrac(lay) * (kb_mn2o(indm + 1, ig) - kb_mn2o(indm, ig));
        c.taug(lay, ngs8 + ig) = c.colch4(lay) * simple4(absb, ind0, ind1, lay, ig) + adjcoln2o * absn2o;
        c.fracs(lay, ngs8 + ig) = fracrefb(ig, 1);
      }
    }
  }
  // ---- band 10: 1390-1480 (h2o / h2o) :2043-2110
  {
    const LwBand& B = S.band[10];
    const int ngs9 = 108;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 10) + 1, ind1 = ind1a(lay, 10) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        c.taug(lay, ngs9 + ig) = c.colh2o(lay) * simple4(absa, ind0, ind1, lay, ig) + tauself + taufor;
        c.fracs(lay, ngs9 + ig) = fracrefa(ig, 1);
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 10) + 1, ind1 = ind1b(lay, 10) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double taufor = tfor(forref, lay, ig);
        c.taug(lay, ngs9 + ig) = c.colh2o(lay) * simple4(absb, ind0, ind1, lay, ig) + taufor;
        c.fracs(lay, ngs9 + ig) = fracrefb(ig, 1);
      }
    }
  }
  // ---- band 11: 1480-1800 (h2o; minor o2 / h2o; minor o2) :2113-2195
  {
    const LwBand& B = S.band[11];
    const int ngs10 = 114;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &ka_mo2 = B["ka_mo2"], &kb_mo2 = B["kb_mo2"], &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 11) + 1, ind1 = ind1a(lay, 11) + 1, indm = c.indminor[lay];
      double scaleo2 = c.colo2(lay) * c.scaleminor(lay);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double tauo2 = scaleo2 * (ka_mo2(indm, ig) + c.minorfrac(lay) * (ka_mo2(indm + 1, ig) - ka_mo2(indm, ig)));
        c.taug(lay, ngs10 + ig) = c.colh2o(lay) * simple4(absa, ind0, ind1, lay, ig) + tauself + taufor + tauo2;
        c.fracs(lay, ngs10 + ig) = fracrefa(ig, 1);
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 11) + 1, ind1 = ind1b(lay, 11) + 1, indm = c.indminor[lay];
      double scaleo2 = c.colo2(lay) * c.scaleminor(lay);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double taufor = tfor(forref, lay, ig);
        double tauo2 = scaleo2 * (kb_mo2(indm, ig) + c.minorfrac(lay) * (kb_mo2(indm + 1, ig) - kb_mo2(indm, ig)));
        c.taug(lay, ngs10 + ig) = c.colh2o(lay) * simple4(absb, ind0, ind1, lay, ig) + taufor + tauo2;
        c.fracs(lay, ngs10 + ig) = fracrefb(ig, 1);
      }
    }
  }
  // ---- band 12: 1800-2080 (low key - h2o,co2; high - nothing) :2198-2395
  {
    const LwBand& B = S.band[12];
    const int ngs11 = 122;
    const A2 &absa = B["ka"], &selfref = B["selfref"], &forref = B["forref"], &fracrefa = B["fracrefa"];
    double refrat_planck_a = chi(1, 10) / chi(2, 10);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2oco2(lay), c.colco2(lay), 8.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2oco2_1(lay), c.colco2(lay), 8.);
      Bin sp = binspec(c.colh2o(lay), refrat_planck_a, c.colco2(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 12) + s0.js, ind1 = ind1a(lay, 12) + s1.js;
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs11 + ig) = tau_major + tau_major1 + tauself + taufor;
        c.fracs(lay, ngs11 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay)
      for (int ig = 1; ig <= B.ng; ++ig) {
        c.taug(lay, ngs11 + ig) = 0.0;
        c.fracs(lay, ngs11 + ig) = 0.0;
      }
  }
  // ---- band 13: 2080-2250 (low key - h2o,n2o; low minor - co2, co; high minor - o3) :2398-2650
  {
    const LwBand& B = S.band[13];
    const int ngs12 = 130;
    const A2 &absa = B["ka"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &ka_mco2 = B["ka_mco2"], &ka_mco = B["ka_mco"], &kb_mo3 = B["kb_mo3"], &fracrefa = B["fracrefa"],
             &fracrefb = B["fracrefb"];
    double refrat_planck_a = chi(1, 5) / chi(4, 5), refrat_m_a = chi(1, 1) / chi(4, 1),
           refrat_m_a3 = chi(1, 3) / chi(4, 3);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2on2o(lay), c.coln2o(lay), 8.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2on2o_1(lay), c.coln2o(lay), 8.);
      Bin sm = binspec(c.colh2o(lay), refrat_m_a, c.coln2o(lay), 8.);
      int jmco2 = sm.js;
      double fmco2 = sm.fs;
      double chi_co2 = c.colco2(lay) / (c.coldry(lay));
      double ratco2 = 1.e20 * chi_co2 / 3.55e-4;
      double adjcolco2;
      if (ratco2 > 3.0) {
        double adjfac = 2.0 + std::pow(ratco2 - 2.0, 0.68);
        adjcolco2 = adjfac * (double)3.55e-4f * c.coldry(lay) * 1.e-20;  // default-real 3.55e-4 (:2479)
      } else
        adjcolco2 = c.colco2(lay);
      Bin sm3 = binspec(c.colh2o(lay), refrat_m_a3, c.coln2o(lay), 8.);
      int jmco = sm3.js;
      double fmco = sm3.fs;
      Bin sp = binspec(c.colh2o(lay), refrat_planck_a, c.coln2o(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 13) + s0.js, ind1 = ind1a(lay, 13) + s1.js, indm = c.indminor[lay];
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double absco2 = minor2(ka_mco2, 9, jmco2, indm, ig, fmco2, c.minorfrac(lay));
        double absco = minor2(ka_mco, 9, jmco, indm, ig, fmco, c.minorfrac(lay));
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs12 + ig) =
            tau_major + tau_major1 + tauself + taufor + adjcolco2 * absco2 + c.colco(lay) * absco;
        c.fracs(lay, ngs12 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int indm = c.indminor[lay];
      for (int ig = 1; ig <= B.ng; ++ig) {
        double abso3 = kb_mo3(indm, ig) + c.minorfrac(lay) * (kb_mo3(indm + 1, ig) - kb_mo3(indm, ig));
        c.taug(lay, ngs12 + ig) = c.colo3(lay) * abso3;
        c.fracs(lay, ngs12 + ig) = fracrefb(ig, 1);
      }
    }
  }
  // ---- band 14: 2250-2380 (co2 / co2) :2653-2718
  {
    const LwBand& B = S.band[14];
    const int ngs13 = 134;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 14) + 1, ind1 = ind1a(lay, 14) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        c.taug(lay, ngs13 + ig) = c.colco2(lay) * simple4(absa, ind0, ind1, lay, ig) + tauself + taufor;
        c.fracs(lay, ngs13 + ig) = fracrefa(ig, 1);
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 14) + 1, ind1 = ind1b(lay, 14) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        c.taug(lay, ngs13 + ig) = c.colco2(lay) * simple4(absb, ind0, ind1, lay, ig);
        c.fracs(lay, ngs13 + ig) = fracrefb(ig, 1);
      }
    }
  }
  // ---- band 15: 2380-2600 (low key - n2o,co2; low minor - n2; high - nothing) :2721-2936
  {
    const LwBand& B = S.band[15];
    const int ngs14 = 136;
    const A2 &absa = B["ka"], &selfref = B["selfref"], &forref = B["forref"], &ka_mn2 = B["ka_mn2"],
             &fracrefa = B["fracrefa"];
    double refrat_planck_a = chi(4, 1) / chi(2, 1), refrat_m_a = chi(4, 1) / chi(2, 1);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.coln2o(lay), c.rat_n2oco2(lay), c.colco2(lay), 8.);
      Bin s1 = binspec(c.coln2o(lay), c.rat_n2oco2_1(lay), c.colco2(lay), 8.);
      Bin sm = binspec(c.coln2o(lay), refrat_m_a, c.colco2(lay), 8.);
      int jmn2 = sm.js;
      double fmn2 = sm.fs;
      Bin sp = binspec(c.coln2o(lay), refrat_planck_a, c.colco2(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 15) + s0.js, ind1 = ind1a(lay, 15) + s1.js, indm = c.indminor[lay];
      double scalen2 = c.colbrd(lay) * c.scaleminor(lay);
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double taun2 = scalen2 * minor2(ka_mn2, 9, jmn2, indm, ig, fmn2, c.minorfrac(lay));
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs14 + ig) = tau_major + tau_major1 + tauself + taufor + taun2;
        c.fracs(lay, ngs14 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay)
      for (int ig = 1; ig <= B.ng; ++ig) {
        c.taug(lay, ngs14 + ig) = 0.0;
        c.fracs(lay, ngs14 + ig) = 0.0;
      }
  }
  // ---- band 16: 2600-3250 (low key - h2o,ch4; high key - ch4) :2939-3145
  {
    const LwBand& B = S.band[16];
    const int ngs15 = 138;
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const A2 &fracrefa = B["fracrefa"], &fracrefb = B["fracrefb"];
    double refrat_planck_a = chi(1, 6) / chi(6, 6);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin s0 = binspec(c.colh2o(lay), c.rat_h2och4(lay), c.colch4(lay), 8.);
      Bin s1 = binspec(c.colh2o(lay), c.rat_h2och4_1(lay), c.colch4(lay), 8.);
      Bin sp = binspec(c.colh2o(lay), refrat_planck_a, c.colch4(lay), 8.);
      int jpl = sp.js;
      double fpl = sp.fs;
      int ind0 = ind0a(lay, 16) + s0.js, ind1 = ind1a(lay, 16) + s1.js;
      Sten t0 = stencil3(s0.specparm, s0.fs, c.fac00(lay), c.fac10(lay), 9);
      Sten t1 = stencil3(s1.specparm, s1.fs, c.fac01(lay), c.fac11(lay), 9);
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauself = tself(selfref, lay, ig), taufor = tfor(forref, lay, ig);
        double tau_major = s0.speccomb * dot(t0, absa, ind0, ig);
        double tau_major1 = s1.speccomb * dot(t1, absa, ind1, ig);
        c.taug(lay, ngs15 + ig) = tau_major + tau_major1 + tauself + taufor;
        c.fracs(lay, ngs15 + ig) = fracrefa(ig, jpl) + fpl * (fracrefa(ig, jpl + 1) - fracrefa(ig, jpl));
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      // nspb(16) = 0 in lwdatinit, so both indices collapse to 1 (restated literally, :3133-3134)
      int ind0 = ind0b(lay, 16) + 1, ind1 = ind1b(lay, 16) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        c.taug(lay, ngs15 + ig) = c.colch4(lay) * simple4(absb, ind0, ind1, lay, ig);
        c.fracs(lay, ngs15 + ig) = fracrefb(ig, 1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// rtrn — rrtmg_lw_rtrn.f90:32-587 (random overlap / clear)
static const int ipat_[3][16] = {{1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
                                 {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5},
                                 {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}};
static const double a0_[16] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66};
static const double a1_[16] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
static const double a2_[16] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};

struct Flux {
  A1 totuflux, totdflux, fnet, htr, totuclfl, totdclfl, fnetc, htrc, dtotuflux_dt, dtotuclfl_dt;
  explicit Flux(int nlay) {
    for (A1* a : {&totuflux, &totdflux, &fnet, &htr, &totuclfl, &totdclfl, &fnetc, &htrc, &dtotuflux_dt, &dtotuclfl_dt})
      *a = A1(nlay + 2, 0);
  }
};

static void rtrn(Col& c, int istart, int iend, int idrv, Flux& F) {
  const int nlayers = c.nlayers, ncbands = c.ncbands;
  const double wtdiff = 0.5, rec_6 = 0.166667;
  const double tblint = 10000.0, bpade = S.bpade;
  const std::vector<double>&tau_tbl = S.tau_tbl, &exp_tbl = S.exp_tbl, &tfn_tbl = S.tfn_tbl;
  int n = nlayers + 2;
  A1 urad(n, 0), drad(n, 0), clrurad(n, 0), clrdrad(n, 0), d_urad_dt(n, 0), d_clrurad_dt(n, 0);
  A1 atrans(n), atot(n), bbugas(n), bbutot(n);
  A2 odcld(n, 16), abscld(n, 16), efclfrac(n, 16);
  std::vector<int> icldlyr(n, 0);
  double secdiff[17];
  for (int ibnd = 1; ibnd <= nbndlw; ++ibnd) {
    if (ibnd == 1 || ibnd == 4 || ibnd >= 10)
      secdiff[ibnd] = 1.66;
    else {
      secdiff[ibnd] = a0_[ibnd - 1] + a1_[ibnd - 1] * std::exp(a2_[ibnd - 1] * c.pwvcm);
      if (secdiff[ibnd] > 1.80) secdiff[ibnd] = 1.80;
      if (secdiff[ibnd] < 1.50) secdiff[ibnd] = 1.50;
    }
  }
  for (int lay = 0; lay <= nlayers; ++lay) {
    urad(lay) = 0.0; drad(lay) = 0.0; F.totuflux(lay) = 0.0; F.totdflux(lay) = 0.0;
    clrurad(lay) = 0.0; clrdrad(lay) = 0.0; F.totuclfl(lay) = 0.0; F.totdclfl(lay) = 0.0;
    d_urad_dt(lay) = 0.0; d_clrurad_dt(lay) = 0.0; F.dtotuflux_dt(lay) = 0.0; F.dtotuclfl_dt(lay) = 0.0;
    if (lay == 0) continue;
    for (int ib = 1; ib <= ncbands; ++ib) {
      if (c.cldfrac(lay) >= 1.e-6) {
        odcld(lay, ib) = secdiff[ib] * c.taucloud(lay, ib);
        double transcld = std::exp(-odcld(lay, ib));
        abscld(lay, ib) = 1. - transcld;
        efclfrac(lay, ib) = abscld(lay, ib) * c.cldfrac(lay);
        icldlyr[lay] = 1;
      } else {
        odcld(lay, ib) = 0.0;
        abscld(lay, ib) = 0.0;
        efclfrac(lay, ib) = 0.0;
        icldlyr[lay] = 0;
      }
    }
  }
  int igc = 1;
  for (int iband = istart; iband <= iend; ++iband) {
    int ib = 1;
    if (ncbands == 1) ib = ipat_[0][iband - 1];
    else if (ncbands == 5) ib = ipat_[1][iband - 1];
    else if (ncbands == 16) ib = ipat_[2][iband - 1];
    do {  // g-point loop (label 1000)
      double radld = 0., radclrd = 0.;
      int iclddn = 0;
      for (int lev = nlayers; lev >= 1; --lev) {
        double plfrac = c.fracs(lev, igc);
        double blay = c.planklay(lev, iband);
        double dplankup = c.planklev(lev, iband) - blay;
        double dplankdn = c.planklev(lev - 1, iband) - blay;
        double odepth = secdiff[iband] * c.taut(lev, igc);
        if (odepth < 0.0) odepth = 0.0;
        double bbd;
        if (icldlyr[lev] == 1) {
          iclddn = 1;
          double odtot = odepth + odcld(lev, ib);
          double gassrc, bbdtot;
          if (odtot < 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            double odepth_rec = rec_6 * odepth;
            gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans(lev);
            atot(lev) = odtot - 0.5 * odtot * odtot;
            double odtot_rec = rec_6 * odtot;
            bbdtot = plfrac * (blay + dplankdn * odtot_rec);
            bbd = plfrac * (blay + dplankdn * odepth_rec);
            radld = radld - radld * (atrans(lev) + efclfrac(lev, ib) * (1. - atrans(lev))) + gassrc +
                    c.cldfrac(lev) * (bbdtot * atot(lev) - gassrc);
            drad(lev - 1) = drad(lev - 1) + radld;
            bbugas(lev) = plfrac * (blay + dplankup * odepth_rec);
            bbutot(lev) = plfrac * (blay + dplankup * odtot_rec);
          } else if (odepth <= 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            double odepth_rec = rec_6 * odepth;
            gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans(lev);
            odtot = odepth + odcld(lev, ib);
            double tblind = odtot / (bpade + odtot);
            int ittot = f2i(tblint * tblind + 0.5);
            double tfactot = tfn_tbl[ittot];
            bbdtot = plfrac * (blay + tfactot * dplankdn);
            bbd = plfrac * (blay + dplankdn * odepth_rec);
            atot(lev) = 1. - exp_tbl[ittot];
            radld = radld - radld * (atrans(lev) + efclfrac(lev, ib) * (1. - atrans(lev))) + gassrc +
                    c.cldfrac(lev) * (bbdtot * atot(lev) - gassrc);
            drad(lev - 1) = drad(lev - 1) + radld;
            bbugas(lev) = plfrac * (blay + dplankup * odepth_rec);
            bbutot(lev) = plfrac * (blay + tfactot * dplankup);
          } else {
            double tblind = odepth / (bpade + odepth);
            int itgas = f2i(tblint * tblind + 0.5);
            odepth = tau_tbl[itgas];
            atrans(lev) = 1. - exp_tbl[itgas];
            double tfacgas = tfn_tbl[itgas];
            gassrc = atrans(lev) * plfrac * (blay + tfacgas * dplankdn);
            odtot = odepth + odcld(lev, ib);
            tblind = odtot / (bpade + odtot);
            int ittot = f2i(tblint * tblind + 0.5);
            double tfactot = tfn_tbl[ittot];
            bbdtot = plfrac * (blay + tfactot * dplankdn);
            bbd = plfrac * (blay + tfacgas * dplankdn);
            atot(lev) = 1. - exp_tbl[ittot];
            radld = radld - radld * (atrans(lev) + efclfrac(lev, ib) * (1. - atrans(lev))) + gassrc +
                    c.cldfrac(lev) * (bbdtot * atot(lev) - gassrc);
            drad(lev - 1) = drad(lev - 1) + radld;
            bbugas(lev) = plfrac * (blay + tfacgas * dplankup);
            bbutot(lev) = plfrac * (blay + tfactot * dplankup);
          }
        } else {
          if (odepth <= 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            odepth = rec_6 * odepth;
            bbd = plfrac * (blay + dplankdn * odepth);
            bbugas(lev) = plfrac * (blay + dplankup * odepth);
          } else {
            double tblind = odepth / (bpade + odepth);
            int itr = f2i(tblint * tblind + 0.5);
            double transc = exp_tbl[itr];
            atrans(lev) = 1. - transc;
            double tausfac = tfn_tbl[itr];
            bbd = plfrac * (blay + tausfac * dplankdn);
            bbugas(lev) = plfrac * (blay + tausfac * dplankup);
          }
          radld = radld + (bbd - radld) * atrans(lev);
          drad(lev - 1) = drad(lev - 1) + radld;
        }
        if (iclddn == 1) {
          radclrd = radclrd + (bbd - radclrd) * atrans(lev);
          clrdrad(lev - 1) = clrdrad(lev - 1) + radclrd;
        } else {
          radclrd = radld;
          clrdrad(lev - 1) = drad(lev - 1);
        }
      }
      double rad0 = c.fracs(1, igc) * c.plankbnd[iband];
      double d_rad0_dt = 0., d_radlu_dt = 0., d_radclru_dt = 0.;
      if (idrv == 1) d_rad0_dt = c.fracs(1, igc) * c.dplankbnd_dt[iband];
      double reflect = 1. - c.semiss[iband];
      double radlu = rad0 + reflect * radld;
      double radclru = rad0 + reflect * radclrd;
      urad(0) = urad(0) + radlu;
      clrurad(0) = clrurad(0) + radclru;
      if (idrv == 1) {
        d_radlu_dt = d_rad0_dt;
        d_urad_dt(0) = d_urad_dt(0) + d_radlu_dt;
        d_radclru_dt = d_rad0_dt;
        d_clrurad_dt(0) = d_clrurad_dt(0) + d_radclru_dt;
      }
      for (int lev = 1; lev <= nlayers; ++lev) {
        if (icldlyr[lev] == 1) {
          double gassrc = bbugas(lev) * atrans(lev);
          radlu = radlu - radlu * (atrans(lev) + efclfrac(lev, ib) * (1. - atrans(lev))) + gassrc +
                  c.cldfrac(lev) * (bbutot(lev) * atot(lev) - gassrc);
          urad(lev) = urad(lev) + radlu;
          if (idrv == 1) {
            d_radlu_dt = d_radlu_dt * c.cldfrac(lev) * (1.0 - atot(lev)) +
                         d_radlu_dt * (1.0 - c.cldfrac(lev)) * (1.0 - atrans(lev));
            d_urad_dt(lev) = d_urad_dt(lev) + d_radlu_dt;
          }
        } else {
          radlu = radlu + (bbugas(lev) - radlu) * atrans(lev);
          urad(lev) = urad(lev) + radlu;
          if (idrv == 1) {
            d_radlu_dt = d_radlu_dt * (1.0 - atrans(lev));
            d_urad_dt(lev) = d_urad_dt(lev) + d_radlu_dt;
          }
        }
        if (iclddn == 1) {
          radclru = radclru + (bbugas(lev) - radclru) * atrans(lev);
          clrurad(lev) = clrurad(lev) + radclru;
        } else {
          radclru = radlu;
          clrurad(lev) = urad(lev);
        }
        if (idrv == 1) {
          if (iclddn == 1) {
            d_radclru_dt = d_radclru_dt * (1.0 - atrans(lev));
            d_clrurad_dt(lev) = d_clrurad_dt(lev) + d_radclru_dt;
          } else {
            d_radclru_dt = d_radlu_dt;
            d_clrurad_dt(lev) = d_urad_dt(lev);
          }
        }
      }
      igc = igc + 1;
    } while (igc <= ngs_[iband - 1]);
    for (int lev = nlayers; lev >= 0; --lev) {
      double uflux = urad(lev) * wtdiff, dflux = drad(lev) * wtdiff;
      urad(lev) = 0.0;
      drad(lev) = 0.0;
      F.totuflux(lev) = F.totuflux(lev) + uflux * delwave_[iband - 1];
      F.totdflux(lev) = F.totdflux(lev) + dflux * delwave_[iband - 1];
      double uclfl = clrurad(lev) * wtdiff, dclfl = clrdrad(lev) * wtdiff;
      clrurad(lev) = 0.0;
      clrdrad(lev) = 0.0;
      F.totuclfl(lev) = F.totuclfl(lev) + uclfl * delwave_[iband - 1];
      F.totdclfl(lev) = F.totdclfl(lev) + dclfl * delwave_[iband - 1];
    }
    if (idrv == 1)
      for (int lev = nlayers; lev >= 0; --lev) {
        double duflux_dt = d_urad_dt(lev) * wtdiff;
        d_urad_dt(lev) = 0.0;
        F.dtotuflux_dt(lev) = F.dtotuflux_dt(lev) + duflux_dt * delwave_[iband - 1] * S.fluxfac;
        double duclfl_dt = d_clrurad_dt(lev) * wtdiff;
        d_clrurad_dt(lev) = 0.0;
        F.dtotuclfl_dt(lev) = F.dtotuclfl_dt(lev) + duclfl_dt * delwave_[iband - 1] * S.fluxfac;
      }
  }
  F.totuflux(0) = F.totuflux(0) * S.fluxfac;
  F.totdflux(0) = F.totdflux(0) * S.fluxfac;
  F.fnet(0) = F.totuflux(0) - F.totdflux(0);
  F.totuclfl(0) = F.totuclfl(0) * S.fluxfac;
  F.totdclfl(0) = F.totdclfl(0) * S.fluxfac;
  F.fnetc(0) = F.totuclfl(0) - F.totdclfl(0);
  for (int lev = 1; lev <= nlayers; ++lev) {
    F.totuflux(lev) = F.totuflux(lev) * S.fluxfac;
    F.totdflux(lev) = F.totdflux(lev) * S.fluxfac;
    F.fnet(lev) = F.totuflux(lev) - F.totdflux(lev);
    F.totuclfl(lev) = F.totuclfl(lev) * S.fluxfac;
    F.totdclfl(lev) = F.totdclfl(lev) * S.fluxfac;
    F.fnetc(lev) = F.totuclfl(lev) - F.totdclfl(lev);
    int l = lev - 1;
    F.htr(l) = S.heatfac * (F.fnet(l) - F.fnet(lev)) / (c.pz(l) - c.pz(lev));
    F.htrc(l) = S.heatfac * (F.fnetc(l) - F.fnetc(lev)) / (c.pz(l) - c.pz(lev));
  }
  F.htr(nlayers) = 0.0;
  F.htrc(nlayers) = 0.0;
}

// rtrnmr -- rrtmg_lw_rtrnmr.f90:32-779: rtrn with maximum-random cloud overlap (non-McICA icld = 2, 3; also what the
// reference calls for icld = 0, where no layer is cloudy).  cldfrac(0), read at :400-401 for lev = 1 (one element before
// the array; always multiplied by faccld2(1) = facclr2(1) = 0 there), is taken as 0.
static void rtrnmr(Col& c, int istart, int iend, int idrv, Flux& F) {
  const int nlayers = c.nlayers, ncbands = c.ncbands;
  const double wtdiff = 0.5, rec_6 = 0.166667;
  const double tblint = 10000.0, bpade = S.bpade;
  const std::vector<double>&tau_tbl = S.tau_tbl, &exp_tbl = S.exp_tbl, &tfn_tbl = S.tfn_tbl;
  int n = nlayers + 2;
  A1 urad(n, 0), drad(n, 0), clrurad(n, 0), clrdrad(n, 0), d_urad_dt(n, 0), d_clrurad_dt(n, 0);
  A1 atrans(n), atot(n), bbugas(n), bbutot(n);
  A2 odcld(n, 16);
  A1 faccld1(n + 1), faccld2(n + 1), facclr1(n + 1), facclr2(n + 1), faccmb1(n + 1), faccmb2(n + 1);
  A1 faccld1d(n + 1, 0), faccld2d(n + 1, 0), facclr1d(n + 1, 0), facclr2d(n + 1, 0), faccmb1d(n + 1, 0), faccmb2d(n + 1, 0);
  std::vector<int> istcld(n + 2, 0), istcldd(n + 2, 0);
  double fmax, fmin, rat1 = 0., rat2 = 0.;
  double clrradd = 0., cldradd = 0., clrradu = 0., cldradu = 0., oldclr = 0., oldcld = 0., rad = 0., cldsrc, radmod, ttot;
  std::vector<int> icldlyr(n, 0);
  double secdiff[17];
  for (int ibnd = 1; ibnd <= nbndlw; ++ibnd) {
    if (ibnd == 1 || ibnd == 4 || ibnd >= 10)
      secdiff[ibnd] = 1.66;
    else {
      secdiff[ibnd] = a0_[ibnd - 1] + a1_[ibnd - 1] * std::exp(a2_[ibnd - 1] * c.pwvcm);
      if (secdiff[ibnd] > 1.80) secdiff[ibnd] = 1.80;
      if (secdiff[ibnd] < 1.50) secdiff[ibnd] = 1.50;
    }
  }
  for (int lay = 0; lay <= nlayers; ++lay) {
    urad(lay) = 0.0; drad(lay) = 0.0; F.totuflux(lay) = 0.0; F.totdflux(lay) = 0.0;
    clrurad(lay) = 0.0; clrdrad(lay) = 0.0; F.totuclfl(lay) = 0.0; F.totdclfl(lay) = 0.0;
    d_urad_dt(lay) = 0.0; d_clrurad_dt(lay) = 0.0; F.dtotuflux_dt(lay) = 0.0; F.dtotuclfl_dt(lay) = 0.0;
    if (lay == 0) continue;
    for (int ib = 1; ib <= ncbands; ++ib) {
      if (c.cldfrac(lay) >= 1.e-6) {
        odcld(lay, ib) = secdiff[ib] * c.taucloud(lay, ib);
        icldlyr[lay] = 1;
      } else {
        odcld(lay, ib) = 0.0;
        icldlyr[lay] = 0;
      }
    }
  }
  // Maximum/Random cloud overlap parameters (:322-479)
  auto cf = [&](int lev) { return lev >= 1 && lev <= nlayers ? c.cldfrac(lev) : 0.0; };
  istcld[1] = 1;
  istcldd[nlayers] = 1;
  for (int lev = 1; lev <= nlayers; ++lev) {
    if (icldlyr[lev] == 1) {
      istcld[lev + 1] = 0;
      if (lev == nlayers) {
        faccld1(lev + 1) = 0.; faccld2(lev + 1) = 0.; facclr1(lev + 1) = 0.; facclr2(lev + 1) = 0.;
        faccmb1(lev + 1) = 0.; faccmb2(lev + 1) = 0.;
      } else if (cf(lev + 1) >= cf(lev)) {
        faccld1(lev + 1) = 0.;
        faccld2(lev + 1) = 0.;
        if (istcld[lev] == 1) {
          facclr1(lev + 1) = 0.;
          facclr2(lev + 1) = 0.;
          if (cf(lev) < 1.) facclr2(lev + 1) = (cf(lev + 1) - cf(lev)) / (1. - cf(lev));
          facclr2(lev) = 0.;
          faccld2(lev) = 0.;
        } else {
          fmax = std::max(cf(lev), cf(lev - 1));
          if (cf(lev + 1) > fmax) {
            facclr1(lev + 1) = rat2;
            facclr2(lev + 1) = (cf(lev + 1) - fmax) / (1. - fmax);
          } else if (cf(lev + 1) < fmax) {
            facclr1(lev + 1) = (cf(lev + 1) - cf(lev)) / (cf(lev - 1) - cf(lev));
            facclr2(lev + 1) = 0.;
          } else {
            facclr1(lev + 1) = rat2;
            facclr2(lev + 1) = 0.;
          }
        }
        if (facclr1(lev + 1) > 0. || facclr2(lev + 1) > 0.) { rat1 = 1.; rat2 = 0.; }
        else { rat1 = 0.; rat2 = 0.; }
      } else {
        facclr1(lev + 1) = 0.;
        facclr2(lev + 1) = 0.;
        if (istcld[lev] == 1) {
          faccld1(lev + 1) = 0.;
          faccld2(lev + 1) = (cf(lev) - cf(lev + 1)) / cf(lev);
          facclr2(lev) = 0.;
          faccld2(lev) = 0.;
        } else {
          fmin = std::min(cf(lev), cf(lev - 1));
          if (cf(lev + 1) <= fmin) {
            faccld1(lev + 1) = rat1;
            faccld2(lev + 1) = (fmin - cf(lev + 1)) / fmin;
          } else {
            faccld1(lev + 1) = (cf(lev) - cf(lev + 1)) / (cf(lev) - fmin);
            faccld2(lev + 1) = 0.;
          }
        }
        if (faccld1(lev + 1) > 0. || faccld2(lev + 1) > 0.) { rat1 = 0.; rat2 = 1.; }
        else { rat1 = 0.; rat2 = 0.; }
      }
      faccmb1(lev + 1) = facclr1(lev + 1) * faccld2(lev) * cf(lev - 1);
      faccmb2(lev + 1) = faccld1(lev + 1) * facclr2(lev) * (1. - cf(lev - 1));
    } else {
      istcld[lev + 1] = 1;
    }
  }
  for (int lev = nlayers; lev >= 1; --lev) {
    if (icldlyr[lev] == 1) {
      istcldd[lev - 1] = 0;
      if (lev == 1) {
        faccld1d(lev - 1) = 0.; faccld2d(lev - 1) = 0.; facclr1d(lev - 1) = 0.; facclr2d(lev - 1) = 0.;
        faccmb1d(lev - 1) = 0.; faccmb2d(lev - 1) = 0.;
      } else if (cf(lev - 1) >= cf(lev)) {
        faccld1d(lev - 1) = 0.;
        faccld2d(lev - 1) = 0.;
        if (istcldd[lev] == 1) {
          facclr1d(lev - 1) = 0.;
          facclr2d(lev - 1) = 0.;
          if (cf(lev) < 1.) facclr2d(lev - 1) = (cf(lev - 1) - cf(lev)) / (1. - cf(lev));
          facclr2d(lev) = 0.;
          faccld2d(lev) = 0.;
        } else {
          fmax = std::max(cf(lev), cf(lev + 1));
          if (cf(lev - 1) > fmax) {
            facclr1d(lev - 1) = rat2;
            facclr2d(lev - 1) = (cf(lev - 1) - fmax) / (1. - fmax);
          } else if (cf(lev - 1) < fmax) {
            facclr1d(lev - 1) = (cf(lev - 1) - cf(lev)) / (cf(lev + 1) - cf(lev));
            facclr2d(lev - 1) = 0.;
          } else {
            facclr1d(lev - 1) = rat2;
            facclr2d(lev - 1) = 0.;
          }
        }
        if (facclr1d(lev - 1) > 0. || facclr2d(lev - 1) > 0.) { rat1 = 1.; rat2 = 0.; }
        else { rat1 = 0.; rat2 = 0.; }
      } else {
        facclr1d(lev - 1) = 0.;
        facclr2d(lev - 1) = 0.;
        if (istcldd[lev] == 1) {
          faccld1d(lev - 1) = 0.;
          faccld2d(lev - 1) = (cf(lev) - cf(lev - 1)) / cf(lev);
          facclr2d(lev) = 0.;
          faccld2d(lev) = 0.;
        } else {
          fmin = std::min(cf(lev), cf(lev + 1));
          if (cf(lev - 1) <= fmin) {
            faccld1d(lev - 1) = rat1;
            faccld2d(lev - 1) = (fmin - cf(lev - 1)) / fmin;
          } else {
            faccld1d(lev - 1) = (cf(lev) - cf(lev - 1)) / (cf(lev) - fmin);
            faccld2d(lev - 1) = 0.;
          }
        }
        if (faccld1d(lev - 1) > 0. || faccld2d(lev - 1) > 0.) { rat1 = 0.; rat2 = 1.; }
        else { rat1 = 0.; rat2 = 0.; }
      }
      faccmb1d(lev - 1) = facclr1d(lev - 1) * faccld2d(lev) * cf(lev + 1);
      faccmb2d(lev - 1) = faccld1d(lev - 1) * facclr2d(lev) * (1. - cf(lev + 1));
    } else {
      istcldd[lev - 1] = 1;
    }
  }
  int igc = 1;
  for (int iband = istart; iband <= iend; ++iband) {
    int ib = 1;
    if (ncbands == 1) ib = ipat_[0][iband - 1];
    else if (ncbands == 5) ib = ipat_[1][iband - 1];
    else if (ncbands == 16) ib = ipat_[2][iband - 1];
    do {  // g-point loop (label 1000)
      double radld = 0., radclrd = 0.;
      int iclddn = 0;
      for (int lev = nlayers; lev >= 1; --lev) {
        double plfrac = c.fracs(lev, igc);
        double blay = c.planklay(lev, iband);
        double dplankup = c.planklev(lev, iband) - blay;
        double dplankdn = c.planklev(lev - 1, iband) - blay;
        double odepth = secdiff[iband] * c.taut(lev, igc);
        if (odepth < 0.0) odepth = 0.0;
        double bbd;
        if (icldlyr[lev] == 1) {
          iclddn = 1;
          double odtot = odepth + odcld(lev, ib);
          double gassrc, bbdtot;
          if (odtot < 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            double odepth_rec = rec_6 * odepth;
            gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans(lev);
            atot(lev) = odtot - 0.5 * odtot * odtot;
            double odtot_rec = rec_6 * odtot;
            bbdtot = plfrac * (blay + dplankdn * odtot_rec);
            bbd = plfrac * (blay + dplankdn * odepth_rec);
            bbugas(lev) = plfrac * (blay + dplankup * odepth_rec);
            bbutot(lev) = plfrac * (blay + dplankup * odtot_rec);
          } else if (odepth <= 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            double odepth_rec = rec_6 * odepth;
            gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans(lev);
            odtot = odepth + odcld(lev, ib);
            double tblind = odtot / (bpade + odtot);
            int ittot = f2i(tblint * tblind + 0.5);
            double tfactot = tfn_tbl[ittot];
            bbdtot = plfrac * (blay + tfactot * dplankdn);
            bbd = plfrac * (blay + dplankdn * odepth_rec);
            atot(lev) = 1. - exp_tbl[ittot];
            bbugas(lev) = plfrac * (blay + dplankup * odepth_rec);
            bbutot(lev) = plfrac * (blay + tfactot * dplankup);
          } else {
            double tblind = odepth / (bpade + odepth);
            int itgas = f2i(tblint * tblind + 0.5);
            odepth = tau_tbl[itgas];
            atrans(lev) = 1. - exp_tbl[itgas];
            double tfacgas = tfn_tbl[itgas];
            gassrc = atrans(lev) * plfrac * (blay + tfacgas * dplankdn);
            odtot = odepth + odcld(lev, ib);
            tblind = odtot / (bpade + odtot);
            int ittot = f2i(tblint * tblind + 0.5);
            double tfactot = tfn_tbl[ittot];
            bbdtot = plfrac * (blay + tfactot * dplankdn);
            bbd = plfrac * (blay + tfacgas * dplankdn);
            atot(lev) = 1. - exp_tbl[ittot];
            bbugas(lev) = plfrac * (blay + tfacgas * dplankup);
            bbutot(lev) = plfrac * (blay + tfactot * dplankup);
          }
          // :591-616
          if (istcldd[lev] == 1) {
            cldradd = c.cldfrac(lev) * radld;
            clrradd = radld - cldradd;
            oldcld = cldradd;
            oldclr = clrradd;
            rad = 0.;
          }
          ttot = 1. - atot(lev);
          cldsrc = bbdtot * atot(lev);
          cldradd = cldradd * ttot + c.cldfrac(lev) * cldsrc;
          clrradd = clrradd * (1. - atrans(lev)) + (1. - c.cldfrac(lev)) * gassrc;
          radld = cldradd + clrradd;
          drad(lev - 1) = drad(lev - 1) + radld;
          radmod = rad * (facclr1d(lev - 1) * (1. - atrans(lev)) + faccld1d(lev - 1) * ttot) - faccmb1d(lev - 1) * gassrc +
                   faccmb2d(lev - 1) * cldsrc;
          oldcld = cldradd - radmod;
          oldclr = clrradd + radmod;
          rad = -radmod + facclr2d(lev - 1) * oldclr - faccld2d(lev - 1) * oldcld;
          cldradd = cldradd + rad;
          clrradd = clrradd - rad;
        } else {
          if (odepth <= 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            odepth = rec_6 * odepth;
            bbd = plfrac * (blay + dplankdn * odepth);
            bbugas(lev) = plfrac * (blay + dplankup * odepth);
          } else {
            double tblind = odepth / (bpade + odepth);
            int itr = f2i(tblint * tblind + 0.5);
            double transc = exp_tbl[itr];
            atrans(lev) = 1. - transc;
            double tausfac = tfn_tbl[itr];
            bbd = plfrac * (blay + tausfac * dplankdn);
            bbugas(lev) = plfrac * (blay + tausfac * dplankup);
          }
          radld = radld + (bbd - radld) * atrans(lev);
          drad(lev - 1) = drad(lev - 1) + radld;
        }
        if (iclddn == 1) {
          radclrd = radclrd + (bbd - radclrd) * atrans(lev);
          clrdrad(lev - 1) = clrdrad(lev - 1) + radclrd;
        } else {
          radclrd = radld;
          clrdrad(lev - 1) = drad(lev - 1);
        }
      }
      double rad0 = c.fracs(1, igc) * c.plankbnd[iband];
      double d_rad0_dt = 0., d_radlu_dt = 0., d_radclru_dt = 0.;
      if (idrv == 1) d_rad0_dt = c.fracs(1, igc) * c.dplankbnd_dt[iband];
      double reflect = 1. - c.semiss[iband];
      double radlu = rad0 + reflect * radld;
      double radclru = rad0 + reflect * radclrd;
      urad(0) = urad(0) + radlu;
      clrurad(0) = clrurad(0) + radclru;
      if (idrv == 1) {
        d_radlu_dt = d_rad0_dt;
        d_urad_dt(0) = d_urad_dt(0) + d_radlu_dt;
        d_radclru_dt = d_rad0_dt;
        d_clrurad_dt(0) = d_clrurad_dt(0) + d_radclru_dt;
      }
      for (int lev = 1; lev <= nlayers; ++lev) {
        if (icldlyr[lev] == 1) {
          double gassrc = bbugas(lev) * atrans(lev);
          if (istcld[lev] == 1) {
            cldradu = c.cldfrac(lev) * radlu;
            clrradu = radlu - cldradu;
            oldcld = cldradu;
            oldclr = clrradu;
            rad = 0.;
          }
          ttot = 1. - atot(lev);
          cldsrc = bbutot(lev) * atot(lev);
          cldradu = cldradu * ttot + c.cldfrac(lev) * cldsrc;
          clrradu = clrradu * (1.0 - atrans(lev)) + (1. - c.cldfrac(lev)) * gassrc;
          radlu = cldradu + clrradu;
          urad(lev) = urad(lev) + radlu;
          radmod = rad * (facclr1(lev + 1) * (1.0 - atrans(lev)) + faccld1(lev + 1) * ttot) - faccmb1(lev + 1) * gassrc +
                   faccmb2(lev + 1) * cldsrc;
          oldcld = cldradu - radmod;
          oldclr = clrradu + radmod;
          rad = -radmod + facclr2(lev + 1) * oldclr - faccld2(lev + 1) * oldcld;
          cldradu = cldradu + rad;
          clrradu = clrradu - rad;
          if (idrv == 1) {
            d_radlu_dt = d_radlu_dt * c.cldfrac(lev) * (1.0 - atot(lev)) +
                         d_radlu_dt * (1.0 - c.cldfrac(lev)) * (1.0 - atrans(lev));
            d_urad_dt(lev) = d_urad_dt(lev) + d_radlu_dt;
          }
        } else {
          radlu = radlu + (bbugas(lev) - radlu) * atrans(lev);
          urad(lev) = urad(lev) + radlu;
          if (idrv == 1) {
            d_radlu_dt = d_radlu_dt * (1.0 - atrans(lev));
            d_urad_dt(lev) = d_urad_dt(lev) + d_radlu_dt;
          }
        }
        if (iclddn == 1) {
          radclru = radclru + (bbugas(lev) - radclru) * atrans(lev);
          clrurad(lev) = clrurad(lev) + radclru;
        } else {
          radclru = radlu;
          clrurad(lev) = urad(lev);
        }
        if (idrv == 1) {
          if (iclddn == 1) {
            d_radclru_dt = d_radclru_dt * (1.0 - atrans(lev));
            d_clrurad_dt(lev) = d_clrurad_dt(lev) + d_radclru_dt;
          } else {
            d_radclru_dt = d_radlu_dt;
            d_clrurad_dt(lev) = d_urad_dt(lev);
          }
        }
      }
      igc = igc + 1;
    } while (igc <= ngs_[iband - 1]);
    for (int lev = nlayers; lev >= 0; --lev) {
      double uflux = urad(lev) * wtdiff, dflux = drad(lev) * wtdiff;
      urad(lev) = 0.0;
      drad(lev) = 0.0;
      F.totuflux(lev) = F.totuflux(lev) + uflux * delwave_[iband - 1];
      F.totdflux(lev) = F.totdflux(lev) + dflux * delwave_[iband - 1];
      double uclfl = clrurad(lev) * wtdiff, dclfl = clrdrad(lev) * wtdiff;
      clrurad(lev) = 0.0;
      clrdrad(lev) = 0.0;
      F.totuclfl(lev) = F.totuclfl(lev) + uclfl * delwave_[iband - 1];
      F.totdclfl(lev) = F.totdclfl(lev) + dclfl * delwave_[iband - 1];
    }
    if (idrv == 1)
      for (int lev = nlayers; lev >= 0; --lev) {
        double duflux_dt = d_urad_dt(lev) * wtdiff;
        d_urad_dt(lev) = 0.0;
        F.dtotuflux_dt(lev) = F.dtotuflux_dt(lev) + duflux_dt * delwave_[iband - 1] * S.fluxfac;
        double duclfl_dt = d_clrurad_dt(lev) * wtdiff;
        d_clrurad_dt(lev) = 0.0;
        F.dtotuclfl_dt(lev) = F.dtotuclfl_dt(lev) + duclfl_dt * delwave_[iband - 1] * S.fluxfac;
      }
  }
  F.totuflux(0) = F.totuflux(0) * S.fluxfac;
  F.totdflux(0) = F.totdflux(0) * S.fluxfac;
  F.fnet(0) = F.totuflux(0) - F.totdflux(0);
  F.totuclfl(0) = F.totuclfl(0) * S.fluxfac;
  F.totdclfl(0) = F.totdclfl(0) * S.fluxfac;
  F.fnetc(0) = F.totuclfl(0) - F.totdclfl(0);
  for (int lev = 1; lev <= nlayers; ++lev) {
    F.totuflux(lev) = F.totuflux(lev) * S.fluxfac;
    F.totdflux(lev) = F.totdflux(lev) * S.fluxfac;
    F.fnet(lev) = F.totuflux(lev) - F.totdflux(lev);
    F.totuclfl(lev) = F.totuclfl(lev) * S.fluxfac;
    F.totdclfl(lev) = F.totdclfl(lev) * S.fluxfac;
    F.fnetc(lev) = F.totuclfl(lev) - F.totdclfl(lev);
    int l = lev - 1;
    F.htr(l) = S.heatfac * (F.fnet(l) - F.fnet(lev)) / (c.pz(l) - c.pz(lev));
    F.htrc(l) = S.heatfac * (F.fnetc(l) - F.fnetc(lev)) / (c.pz(l) - c.pz(lev));
  }
  F.htr(nlayers) = 0.0;
  F.htrc(nlayers) = 0.0;
}

}  // namespace orc

// =============================================================================================
// C entry points (ctypes)
using namespace orc;
static std::string g_err;

extern "C" const char* orc_last_error() { return g_err.c_str(); }

// rrtmg_set_constants (rrlw_con.f90:46-71)
extern "C" void orc_lw_set_constants(double pi, double grav, double planck, double boltz, double clight, double avogad,
                                     double alosmt, double gascon, double sbcnst, double secdy) {
  S.pi = pi; S.grav = grav; S.planck = planck; S.boltz = boltz; S.clight = clight; S.avogad = avogad;
  S.alosmt = alosmt; S.gascon = gascon; S.sbcnst = sbcnst; S.secdy = secdy;
}

// rrtmg_lw_ini_wrapper (rrtmg_lw_c_binder.f90:39-48); raw_blob = packed 16-g tables
extern "C" int orc_lw_ini(const char* raw_blob, double cpdair) {
  try {
    Blob b(raw_blob);
    lw_ini(b, cpdair);
  } catch (std::exception& e) {
    g_err = e.what();
    return 1;
  }
  return 0;
}

// debugging/inspection hook for tests: copy a reduced table out (Fortran order (lead, ng))
extern "C" int orc_lw_get_reduced(int band, const char* name, double* out, int64_t cap) {
  auto it = S.band[band].t.find(name);
  if (it == S.band[band].t.end()) return -1;
  int64_t n = (int64_t)it->second.d.size();
  if (out && cap >= n) std::memcpy(out, it->second.d.data(), sizeof(double) * (size_t)n);
  return (int)n;
}
extern "C" void orc_lw_get_exp_tables(double* tau, double* ex, double* tfn) {
  std::memcpy(tau, S.tau_tbl.data(), sizeof(double) * (ntbl + 1));
  std::memcpy(ex, S.exp_tbl.data(), sizeof(double) * (ntbl + 1));
  std::memcpy(tfn, S.tfn_tbl.data(), sizeof(double) * (ntbl + 1));
}

// rrtmg_lw_nomcica_wrapper (rrtmg_lw_c_binder.f90:176-256) -> rrtmg_lw (rrtmg_lw_rad.nomcica.f90:80-569).
// All arrays use the reference ABI layout (Fortran (ncol,nlay[,..]) == C (.., nlay, ncol)).
// Optional debug outputs taug/fracs: (ncol, nlay, 140) Fortran order, may be null.
extern "C" int orc_lw_nomcica(int ncol, int nlay, int* icld, int idrv, const double* play, const double* plev,
                              const double* tlay, const double* tlev, const double* tsfc, const double* h2ovmr,
                              const double* o3vmr, const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                              const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                              const double* cfc22vmr, const double* ccl4vmr, const double* emis, int inflglw,
                              int iceflglw, int liqflglw, const double* cldfr, const double* taucld,
                              const double* cicewp, const double* cliqwp, const double* reice, const double* reliq,
                              const double* tauaer, double* uflx, double* dflx, double* hr, double* uflxc,
                              double* dflxc, double* hrc, double* duflx_dt, double* duflxc_dt, double* dbg_taug,
                              double* dbg_fracs) {
  if (!S.ready) { g_err = "orc_lw_ini not called"; return 1; }
  S.oneminus = 1. - 1.e-6;
  S.pi = 2. * std::asin(1.);
  S.fluxfac = S.pi * 2.e4;
  const int istart = 1, iend = 16;
  if (*icld < 0 || *icld > 3) *icld = 2;
  const int iaer = 10;
  LwIn in{ncol, nlay, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr,
          cfc22vmr, ccl4vmr, emis, cldfr, taucld, cicewp, cliqwp, reice, reliq, tauaer};
  Col c(nlay);
  Flux F(nlay);
  for (int iplon = 1; iplon <= ncol; ++iplon) {
    inatm(in, iplon, *icld, iaer, inflglw, iceflglw, liqflglw, c);
    if (cldprop(c, g_err)) return 2;
    setcoef(c, istart, idrv);
    taumol(c);
    for (int k = 1; k <= nlay; ++k)
      for (int ig = 1; ig <= ngptlw; ++ig) {
        static const int* ngb = nullptr;
        (void)ngb;
        int ib = 1;
        while (ig > ngs_[ib - 1]) ++ib;  // ngb(ig)
        c.taut(k, ig) = c.taug(k, ig) + c.taua(k, ib);
      }
    if (*icld == 1) {
      rtrn(c, istart, iend, idrv, F);
    } else if (*icld == 0) {
      // icld = 0 goes through rtrnmr in the reference (rad.nomcica.f90:527-541); with inatm skipping the
      // cloud copy (cldfrac = 0) rtrnmr's clear path is arithmetically rtrn's clear path.
      rtrn(c, istart, iend, idrv, F);
    } else {
      rtrnmr(c, istart, iend, idrv, F);
    }
    for (int k = 0; k <= nlay; ++k) {
      size_t o = (size_t)(iplon - 1) + (size_t)ncol * k;
      uflx[o] = F.totuflux(k);
      dflx[o] = F.totdflux(k);
      uflxc[o] = F.totuclfl(k);
      dflxc[o] = F.totdclfl(k);
    }
    for (int k = 0; k <= nlay - 1; ++k) {
      size_t o = (size_t)(iplon - 1) + (size_t)ncol * k;
      hr[o] = F.htr(k);
      hrc[o] = F.htrc(k);
    }
    if (idrv == 1)
      for (int k = 0; k <= nlay; ++k) {
        size_t o = (size_t)(iplon - 1) + (size_t)ncol * k;
        duflx_dt[o] = F.dtotuflux_dt(k);
        duflxc_dt[o] = F.dtotuclfl_dt(k);
      }
    if (dbg_taug)
      for (int ig = 1; ig <= ngptlw; ++ig)
        for (int k = 1; k <= nlay; ++k) {
          size_t o = (size_t)(iplon - 1) + (size_t)ncol * ((k - 1) + (size_t)nlay * (ig - 1));
          dbg_taug[o] = c.taug(k, ig);
          if (dbg_fracs) dbg_fracs[o] = c.fracs(k, ig);
        }
  }
  return 0;
}

// =============================================================================================
// McICA longwave: mcica_subcol_lw + cldprmc + rtrnmc (rrtmg_lw_rad.f90:80-576)
namespace orc {

static int ngb_lw_[140];
static void init_ngb_lw() {
  int ib = 1;
  for (int ig = 1; ig <= 140; ++ig) {
    while (ig > ngs_[ib - 1]) ++ib;
    ngb_lw_[ig - 1] = ib;
  }
}

// cldprmc — rrtmg_lw_cldprmc.f90:32-254 (per column; cldfmc/ciwpmc/clwpmc/taucmc are (140, nlay), 1-based views)
static int cldprmc(int nlayers, int inflag, int iceflag, int liqflag, const A2& cldfmc, const A2& ciwpmc, const A2& clwpmc,
                   const A1& reicmc, const A1& relqmc, A2& taucmc, std::string& err) {
  const double cldmin = 1.e-20;
  static const int icb[16] = {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5};
  double abscoice[141] = {0}, abscoliq[141] = {0};
  for (int lay = 1; lay <= nlayers; ++lay)
    for (int ig = 1; ig <= ngptlw; ++ig) {
      double cwp = ciwpmc(ig, lay) + clwpmc(ig, lay);
      if (cldfmc(ig, lay) >= cldmin && (cwp >= cldmin || taucmc(ig, lay) >= cldmin)) {
        if (inflag == 0) return 0;
        else if (inflag == 1) { err = "INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA"; return 1; }
        else if (inflag == 2) {
          double radice = reicmc(lay);
          if (ciwpmc(ig, lay) == 0.0) abscoice[ig] = 0.0;
          else if (iceflag == 0) {
            if (radice < 10.0) { err = "ICE RADIUS TOO SMALL"; return 1; }
            abscoice[ig] = S.absice0(1) + S.absice0(2) / radice;
          } else if (iceflag == 1) {
            if (radice < 13.0 || radice > 130.) { err = "ICE RADIUS OUT OF BOUNDS"; return 1; }
            int ib = icb[ngb_lw_[ig - 1] - 1];
            abscoice[ig] = S.absice1(1, ib) + S.absice1(2, ib) / radice;
          } else if (iceflag == 2) {
            if (radice < 5.0 || radice > 131.0) { err = "ICE RADIUS OUT OF BOUNDS"; return 1; }
            double factor = (radice - 2.) / 3.;
            int index = (int)factor;
            if (index == 43) index = 42;
            double fint = factor - (double)(float)index;
            int ib = ngb_lw_[ig - 1];
            abscoice[ig] = S.absice2(index, ib) + fint * (S.absice2(index + 1, ib) - (S.absice2(index, ib)));
          } else if (iceflag == 3) {
            if (radice < 5.0 || radice > 140.0) { err = "ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; return 1; }
            double factor = (radice - 2.) / 3.;
            int index = (int)factor;
            if (index == 46) index = 45;
            double fint = factor - (double)(float)index;
            int ib = ngb_lw_[ig - 1];
            abscoice[ig] = S.absice3(index, ib) + fint * (S.absice3(index + 1, ib) - (S.absice3(index, ib)));
          }
          if (clwpmc(ig, lay) == 0.0) abscoliq[ig] = 0.0;
          else if (liqflag == 0) abscoliq[ig] = S.absliq0;
          else if (liqflag == 1) {
            double radliq = relqmc(lay);
            if (radliq < 2.5 || radliq > 60.) { err = "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS"; return 1; }
            int index = (int)(radliq - 1.5);
            if (index == 0) index = 1;
            if (index == 58) index = 57;
            double fint = radliq - 1.5 - (double)(float)index;
            int ib = ngb_lw_[ig - 1];
            abscoliq[ig] = S.absliq1(index, ib) + fint * (S.absliq1(index + 1, ib) - (S.absliq1(index, ib)));
          }
          taucmc(ig, lay) = ciwpmc(ig, lay) * abscoice[ig] + clwpmc(ig, lay) * abscoliq[ig];
        }
      }
    }
  return 0;
}

// rtrnmc — rrtmg_lw_rtrnmc.f90:32-576 (idrv = 0 path)
static void rtrnmc(Col& c, int istart, int iend, const A2& cldfmc, const A2& taucmc, Flux& F) {
  const int nlayers = c.nlayers;
  const double wtdiff = 0.5, rec_6 = 0.166667, tblint = 10000.0, bpade = S.bpade;
  const std::vector<double>&tau_tbl = S.tau_tbl, &exp_tbl = S.exp_tbl, &tfn_tbl = S.tfn_tbl;
  int n = nlayers + 2;
  A1 urad(n, 0), drad(n, 0), clrurad(n, 0), clrdrad(n, 0), atrans(n), atot(n), bbugas(n), bbutot(n);
  A2 odcld(n, 140), abscld(n, 140), efclfrac(n, 140);
  std::vector<int> icldlyr(n, 0);
  double secdiff[17];
  for (int ibnd = 1; ibnd <= nbndlw; ++ibnd) {
    if (ibnd == 1 || ibnd == 4 || ibnd >= 10) secdiff[ibnd] = 1.66;
    else {
      secdiff[ibnd] = a0_[ibnd - 1] + a1_[ibnd - 1] * std::exp(a2_[ibnd - 1] * c.pwvcm);
      if (secdiff[ibnd] > 1.80) secdiff[ibnd] = 1.80;
      if (secdiff[ibnd] < 1.50) secdiff[ibnd] = 1.50;
    }
  }
  for (int lay = 0; lay <= nlayers; ++lay) {
    urad(lay) = 0.; drad(lay) = 0.; F.totuflux(lay) = 0.; F.totdflux(lay) = 0.;
    clrurad(lay) = 0.; clrdrad(lay) = 0.; F.totuclfl(lay) = 0.; F.totdclfl(lay) = 0.;
    if (lay == 0) continue;
    icldlyr[lay] = 0;
    for (int ig = 1; ig <= ngptlw; ++ig) {
      if (cldfmc(ig, lay) == 1.) {
        int ib = ngb_lw_[ig - 1];
        odcld(lay, ig) = secdiff[ib] * taucmc(ig, lay);
        double transcld = std::exp(-odcld(lay, ig));
        abscld(lay, ig) = 1. - transcld;
        efclfrac(lay, ig) = abscld(lay, ig) * cldfmc(ig, lay);
        icldlyr[lay] = 1;
      } else {
        odcld(lay, ig) = 0.; abscld(lay, ig) = 0.; efclfrac(lay, ig) = 0.;
      }
    }
  }
  int igc = 1;
  for (int iband = istart; iband <= iend; ++iband) {
    do {
      double radld = 0., radclrd = 0.;
      int iclddn = 0;
      for (int lev = nlayers; lev >= 1; --lev) {
        double plfrac = c.fracs(lev, igc), blay = c.planklay(lev, iband);
        double dplankup = c.planklev(lev, iband) - blay, dplankdn = c.planklev(lev - 1, iband) - blay;
        double odepth = secdiff[iband] * c.taut(lev, igc);
        if (odepth < 0.0) odepth = 0.0;
        double bbd;
        if (icldlyr[lev] == 1) {
          iclddn = 1;
          double odtot = odepth + odcld(lev, igc), gassrc, bbdtot;
          if (odtot < 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            double odepth_rec = rec_6 * odepth;
            gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans(lev);
            atot(lev) = odtot - 0.5 * odtot * odtot;
            double odtot_rec = rec_6 * odtot;
            bbdtot = plfrac * (blay + dplankdn * odtot_rec);
            bbd = plfrac * (blay + dplankdn * odepth_rec);
            bbugas(lev) = plfrac * (blay + dplankup * odepth_rec);
            bbutot(lev) = plfrac * (blay + dplankup * odtot_rec);
          } else if (odepth <= 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            double odepth_rec = rec_6 * odepth;
            gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans(lev);
            odtot = odepth + odcld(lev, igc);
            double tblind = odtot / (bpade + odtot);
            int ittot = f2i(tblint * tblind + 0.5);
            double tfactot = tfn_tbl[ittot];
            bbdtot = plfrac * (blay + tfactot * dplankdn);
            bbd = plfrac * (blay + dplankdn * odepth_rec);
            atot(lev) = 1. - exp_tbl[ittot];
            bbugas(lev) = plfrac * (blay + dplankup * odepth_rec);
            bbutot(lev) = plfrac * (blay + tfactot * dplankup);
          } else {
            double tblind = odepth / (bpade + odepth);
            int itgas = f2i(tblint * tblind + 0.5);
            odepth = tau_tbl[itgas];
            atrans(lev) = 1. - exp_tbl[itgas];
            double tfacgas = tfn_tbl[itgas];
            gassrc = atrans(lev) * plfrac * (blay + tfacgas * dplankdn);
            odtot = odepth + odcld(lev, igc);
            tblind = odtot / (bpade + odtot);
            int ittot = f2i(tblint * tblind + 0.5);
            double tfactot = tfn_tbl[ittot];
            bbdtot = plfrac * (blay + tfactot * dplankdn);
            bbd = plfrac * (blay + tfacgas * dplankdn);
            atot(lev) = 1. - exp_tbl[ittot];
            bbugas(lev) = plfrac * (blay + tfacgas * dplankup);
            bbutot(lev) = plfrac * (blay + tfactot * dplankup);
          }
          radld = radld - radld * (atrans(lev) + efclfrac(lev, igc) * (1. - atrans(lev))) + gassrc +
                  cldfmc(igc, lev) * (bbdtot * atot(lev) - gassrc);
          drad(lev - 1) = drad(lev - 1) + radld;
        } else {
          if (odepth <= 0.06) {
            atrans(lev) = odepth - 0.5 * odepth * odepth;
            odepth = rec_6 * odepth;
            bbd = plfrac * (blay + dplankdn * odepth);
            bbugas(lev) = plfrac * (blay + dplankup * odepth);
          } else {
            double tblind = odepth / (bpade + odepth);
            int itr = f2i(tblint * tblind + 0.5);
            double transc = exp_tbl[itr];
            atrans(lev) = 1. - transc;
            double tausfac = tfn_tbl[itr];
            bbd = plfrac * (blay + tausfac * dplankdn);
            bbugas(lev) = plfrac * (blay + tausfac * dplankup);
          }
          radld = radld + (bbd - radld) * atrans(lev);
          drad(lev - 1) = drad(lev - 1) + radld;
        }
        if (iclddn == 1) {
          radclrd = radclrd + (bbd - radclrd) * atrans(lev);
          clrdrad(lev - 1) = clrdrad(lev - 1) + radclrd;
        } else {
          radclrd = radld;
          clrdrad(lev - 1) = drad(lev - 1);
        }
      }
      double rad0 = c.fracs(1, igc) * c.plankbnd[iband];
      double reflect = 1. - c.semiss[iband];
      double radlu = rad0 + reflect * radld, radclru = rad0 + reflect * radclrd;
      urad(0) = urad(0) + radlu;
      clrurad(0) = clrurad(0) + radclru;
      for (int lev = 1; lev <= nlayers; ++lev) {
        if (icldlyr[lev] == 1) {
          double gassrc = bbugas(lev) * atrans(lev);
          radlu = radlu - radlu * (atrans(lev) + efclfrac(lev, igc) * (1. - atrans(lev))) + gassrc +
                  cldfmc(igc, lev) * (bbutot(lev) * atot(lev) - gassrc);
          urad(lev) = urad(lev) + radlu;
        } else {
          radlu = radlu + (bbugas(lev) - radlu) * atrans(lev);
          urad(lev) = urad(lev) + radlu;
        }
        if (iclddn == 1) {
          radclru = radclru + (bbugas(lev) - radclru) * atrans(lev);
          clrurad(lev) = clrurad(lev) + radclru;
        } else {
          radclru = radlu;
          clrurad(lev) = urad(lev);
        }
      }
      igc = igc + 1;
    } while (igc <= ngs_[iband - 1]);
    for (int lev = nlayers; lev >= 0; --lev) {
      double uflux = urad(lev) * wtdiff, dflux = drad(lev) * wtdiff;
      urad(lev) = 0.; drad(lev) = 0.;
      F.totuflux(lev) = F.totuflux(lev) + uflux * delwave_[iband - 1];
      F.totdflux(lev) = F.totdflux(lev) + dflux * delwave_[iband - 1];
      double uclfl = clrurad(lev) * wtdiff, dclfl = clrdrad(lev) * wtdiff;
      clrurad(lev) = 0.; clrdrad(lev) = 0.;
      F.totuclfl(lev) = F.totuclfl(lev) + uclfl * delwave_[iband - 1];
      F.totdclfl(lev) = F.totdclfl(lev) + dclfl * delwave_[iband - 1];
    }
  }
  F.totuflux(0) = F.totuflux(0) * S.fluxfac; F.totdflux(0) = F.totdflux(0) * S.fluxfac;
  F.fnet(0) = F.totuflux(0) - F.totdflux(0);
  F.totuclfl(0) = F.totuclfl(0) * S.fluxfac; F.totdclfl(0) = F.totdclfl(0) * S.fluxfac;
  F.fnetc(0) = F.totuclfl(0) - F.totdclfl(0);
  for (int lev = 1; lev <= nlayers; ++lev) {
    F.totuflux(lev) = F.totuflux(lev) * S.fluxfac; F.totdflux(lev) = F.totdflux(lev) * S.fluxfac;
    F.fnet(lev) = F.totuflux(lev) - F.totdflux(lev);
    F.totuclfl(lev) = F.totuclfl(lev) * S.fluxfac; F.totdclfl(lev) = F.totdclfl(lev) * S.fluxfac;
    F.fnetc(lev) = F.totuclfl(lev) - F.totdclfl(lev);
    int l = lev - 1;
    F.htr(l) = S.heatfac * (F.fnet(l) - F.fnet(lev)) / (c.pz(l) - c.pz(lev));
    F.htrc(l) = S.heatfac * (F.fnetc(l) - F.fnetc(lev)) / (c.pz(l) - c.pz(lev));
  }
  F.htr(nlayers) = 0.; F.htrc(nlayers) = 0.;
}
}  // namespace orc

// mcica_subcol_lw_wrapper + rrtmg_lw_mcica_wrapper (rrtmg_lw_c_binder.f90:50-174), as chained by
// _rrtmg_lw.pyx:214-320.  cldfmcl_out (140, ncol, nlay) may be null (debug: the generated cloud mask).
extern "C" int orc_lw_mcica(int ncol, int nlay, int* icld, int permuteseed, int irng, const double* play,
                            const double* plev, const double* tlay, const double* tlev, const double* tsfc,
                            const double* h2ovmr, const double* o3vmr, const double* co2vmr, const double* ch4vmr,
                            const double* n2ovmr, const double* o2vmr, const double* cfc11vmr, const double* cfc12vmr,
                            const double* cfc22vmr, const double* ccl4vmr, const double* emis, int inflglw, int iceflglw,
                            int liqflglw, const double* cldfr, const double* taucld, const double* cicewp,
                            const double* cliqwp, const double* reice, const double* reliq, const double* tauaer,
                            double* uflx, double* dflx, double* hr, double* uflxc, double* dflxc, double* hrc,
                            double* cldfmcl_out) {
  if (!S.ready) { g_err = "orc_lw_ini not called"; return 1; }
  init_ngb_lw();
  S.oneminus = 1. - 1.e-6;
  S.pi = 2. * std::asin(1.);
  S.fluxfac = S.pi * 2.e4;
  if (*icld < 0 || *icld > 3) { g_err = "MCICA_SUBCOL: INVALID ICLD"; return 2; }
  const size_t nsub = 140, tot = nsub * (size_t)ncol * nlay;
  std::vector<double> cldfmcl(tot, 0.), ciwpmcl(tot, 0.), clwpmcl(tot, 0.), taucmcl(tot, 0.);
  if (*icld != 0) {
    std::vector<double> pmid((size_t)ncol * nlay);
    for (size_t i = 0; i < pmid.size(); ++i) pmid[i] = play[i] * 1.e2;
    if (generate_stochastic_clouds(ncol, nlay, 140, *icld, irng, pmid.data(), cldfr, cliqwp, cicewp, taucld, 16, ngb_lw_,
                                   0, cldfmcl.data(), clwpmcl.data(), ciwpmcl.data(), taucmcl.data(), permuteseed, g_err,
                                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr))
      return 2;
  }
  if (cldfmcl_out) std::memcpy(cldfmcl_out, cldfmcl.data(), tot * sizeof(double));
  const int istart = 1, iend = 16, iaer = 10, idrv = 0;
  LwIn in{ncol, nlay, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, cfc11vmr, cfc12vmr,
          cfc22vmr, ccl4vmr, emis, cldfr, taucld, cicewp, cliqwp, reice, reliq, tauaer};
  Col c(nlay);
  Flux F(nlay);
  A2 cldfmc(140, nlay + 1), taucmc(140, nlay + 1), ciwpmc(140, nlay + 1), clwpmc(140, nlay + 1);
  A1 reicmc(nlay + 1), relqmc(nlay + 1);
  for (int iplon = 1; iplon <= ncol; ++iplon) {
    inatm(in, iplon, 0, iaer, inflglw, iceflglw, liqflglw, c);  // gas/aerosol part (cloud copy is per g-point below)
    std::fill(cldfmc.d.begin(), cldfmc.d.end(), 0.); std::fill(taucmc.d.begin(), taucmc.d.end(), 0.);
    std::fill(ciwpmc.d.begin(), ciwpmc.d.end(), 0.); std::fill(clwpmc.d.begin(), clwpmc.d.end(), 0.);
    if (*icld >= 1)
      for (int l = 1; l <= nlay; ++l) {
        for (int ig = 1; ig <= 140; ++ig) {
          size_t o = (size_t)(ig - 1) + nsub * ((size_t)(iplon - 1) + (size_t)ncol * (l - 1));
          cldfmc(ig, l) = cldfmcl[o]; taucmc(ig, l) = taucmcl[o]; ciwpmc(ig, l) = ciwpmcl[o]; clwpmc(ig, l) = clwpmcl[o];
        }
        reicmc(l) = reice[(size_t)(iplon - 1) + (size_t)ncol * (l - 1)];
        relqmc(l) = reliq[(size_t)(iplon - 1) + (size_t)ncol * (l - 1)];
      }
    if (cldprmc(nlay, inflglw, iceflglw, liqflglw, cldfmc, ciwpmc, clwpmc, reicmc, relqmc, taucmc, g_err)) return 2;
    setcoef(c, istart, idrv);
    taumol(c);
    for (int k = 1; k <= nlay; ++k)
      for (int ig = 1; ig <= ngptlw; ++ig) c.taut(k, ig) = c.taug(k, ig) + c.taua(k, ngb_lw_[ig - 1]);
    rtrnmc(c, istart, iend, cldfmc, taucmc, F);
    for (int k = 0; k <= nlay; ++k) {
      size_t o = (size_t)(iplon - 1) + (size_t)ncol * k;
      uflx[o] = F.totuflux(k); dflx[o] = F.totdflux(k); uflxc[o] = F.totuclfl(k); dflxc[o] = F.totdclfl(k);
    }
    for (int k = 0; k <= nlay - 1; ++k) {
      size_t o = (size_t)(iplon - 1) + (size_t)ncol * k;
      hr[o] = F.htr(k); hrc[o] = F.htrc(k);
    }
  }
  return 0;
}
