// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never linked into, imported by or called from the
// product (climt_b200/).
//
// CPU restatement (C++17, fp64, column-serial like the reference) of the AER RRTMG_SW path that climt wraps,
// non-McICA driver.  File:line citations are relative to /root/reference/climt/_lib/rrtmg_sw/.
//
//   rrtmg_sw_ini     rrtmg_sw_init.f90:47-173 (+ swdatinit :176-260, swcmbdat :263-386, cmbgb16s..29 :492-1689)
//   inatm_sw         rrtmg_sw_rad.nomcica.f90:846-1539     earth_sun :818-843
//   cldprop_sw       rrtmg_sw_cldprop.f90:53-365
//   setcoef_sw       rrtmg_sw_setcoef.f90:49-305
//   taumol_sw        rrtmg_sw_taumol.f90:50-1790
//   reftra_sw        rrtmg_sw_reftra.f90:48-324
//   vrtqdr_sw        rrtmg_sw_vrtqdr.f90:47-171
//   spcvrt_sw        rrtmg_sw_spcvrt.f90:53-667
//   rrtmg_sw         rrtmg_sw_rad.nomcica.f90:97-816
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ftn.hpp"
#include "mcica_gen.hpp"

namespace orcsw {
using orc::A1;
using orc::A2;
using orc::Blob;
using orc::BlobEntry;

static const int nbndsw = 14, mg = 16, ngptsw = 112, ntbl = 10000, jpb1 = 16, jpb2 = 29, naerec = 6;
static const double wavenum2_[14] = {3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000.,
                                     38000., 50000., 2600.};
static const int nspa_[14] = {9, 9, 9, 9, 1, 9, 9, 1, 9, 1, 0, 1, 9, 1};
static const int nspb_[14] = {1, 5, 1, 1, 1, 5, 1, 0, 1, 0, 0, 1, 5, 1};
static const int ngc_[14] = {6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12};
static const int ngs_[14] = {6, 18, 26, 34, 44, 54, 56, 66, 74, 80, 86, 94, 100, 112};
static const int ngm_[224] = {
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6,            // band 16
    1, 2, 3, 4, 5, 6, 6, 7, 8, 8, 9, 10, 10, 11, 12, 12,       // band 17
    1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8,            // band 18
    1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8,            // band 19
    1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 10, 10, 10, 10,      // band 20
    1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 10, 10, 10, 10,      // band 21
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2,            // band 22
    1, 1, 2, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 10, 10,        // band 23
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,            // band 24
    1, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6,            // band 25
    1, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6,            // band 26
    1, 2, 3, 4, 5, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8,            // band 27
    1, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6,            // band 28
    1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 10, 11, 12};        // band 29
static const int ngn_[112] = {2, 2, 2, 2, 4, 4,                              // band 16
                              1, 1, 1, 1, 1, 2, 1, 2, 1, 2, 1, 2,            // band 17
                              1, 1, 1, 1, 2, 2, 4, 4,                        // band 18
                              1, 1, 1, 1, 2, 2, 4, 4,                        // band 19
                              1, 1, 1, 1, 1, 1, 1, 1, 2, 6,                  // band 20
                              1, 1, 1, 1, 1, 1, 1, 1, 2, 6,                  // band 21
                              8, 8,                                          // band 22
                              2, 2, 1, 1, 1, 1, 1, 1, 2, 4,                  // band 23
                              2, 2, 2, 2, 2, 2, 2, 2,                        // band 24
                              1, 1, 2, 2, 4, 6,                              // band 25
                              1, 1, 2, 2, 4, 6,                              // band 26
                              1, 1, 1, 1, 1, 1, 4, 6,                        // band 27
                              1, 1, 2, 2, 4, 6,                              // band 28
                              1, 1, 1, 1, 2, 2, 2, 2, 1, 1, 1, 1};           // band 29
static const double wt_[16] = {0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893,
                               0.0832767040, 0.0626720116, 0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086,
                               0.0022199750, 0.0014140010, 0.0005330000, 0.0000750000};

struct SwBand {
  int ng = 0;
  double rayl = 0.;
  std::map<std::string, A2> t;
  const A2& operator[](const char* k) const {
    auto it = t.find(k);
    if (it == t.end()) throw std::runtime_error(std::string("sw oracle: missing reduced table ") + k);
    return it->second;
  }
};

struct SwState {
  double pi, grav, avogad, secdy, heatfac, oneminus, bpade;
  std::vector<double> exp_tbl;
  double rwgt[224];
  A1 preflog, tref;
  // rrsw_cld (band index 16..29 -> stored 1..14)
  A2 extliq1, ssaliq1, asyliq1, extice2, ssaice2, asyice2, extice3, ssaice3, asyice3, fdlice3;
  A1 abari, bbari, cbari, dbari, ebari, fbari;
  A2 rsrtaua, rsrpiza, rsrasya;
  SwBand band[15];  // 1..14
  bool ready = false;
};
static SwState S;

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

// cmbgb16s..cmbgb29 (rrtmg_sw_init.f90:492-1689)
static A2 reduce_g(const BlobEntry& e, int ibnd, bool weighted, bool g_first) {
  int ngc = ngc_[ibnd - 1];
  int ngs_prev = ibnd >= 2 ? ngs_[ibnd - 2] : 0;
  if (g_first) {
    int np = e.shape.size() > 1 ? (int)e.shape[1] : 1;
    A2 out(ngc, np);
    for (int jp = 1; jp <= np; ++jp) {
      int iprsm = 0;
      for (int igc = 1; igc <= ngc; ++igc) {
        double sum = 0.;
        for (int ipr = 1; ipr <= ngn_[ngs_prev + igc - 1]; ++ipr) {
          iprsm++;
          double v = e.p[(iprsm - 1) + 16 * (size_t)(jp - 1)];
          sum = weighted ? sum + v * S.rwgt[(iprsm - 1) + 16 * (ibnd - 1)] : sum + v;
        }
        out(igc, jp) = sum;
      }
    }
    return out;
  }
  int64_t lead = e.count / 16;
  A2 out((int)lead, ngc);
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

static void sw_ini(const Blob& b, double cpdair) {
  S.heatfac = S.grav * S.secdy / (cpdair * 1.e2);
  const double pade = 0.278, expeps = 1.e-20;
  S.exp_tbl.assign(ntbl + 1, 0.);
  S.exp_tbl[0] = 1.0;
  S.exp_tbl[ntbl] = expeps;
  S.bpade = 1.0 / pade;
  for (int itr = 1; itr <= ntbl - 1; ++itr) {
    double tfn = (double)itr / (double)ntbl;  // kind=rb here (rrtmg_sw_init.f90:119)
    double tau_tbl = S.bpade * tfn / (1. - tfn);
    S.exp_tbl[itr] = std::exp(-tau_tbl);
    if (S.exp_tbl[itr] <= expeps) S.exp_tbl[itr] = expeps;
  }
  {
    int igcsm = 0;
    double wtsm[17];
    for (int ibnd = 1; ibnd <= nbndsw; ++ibnd) {
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
  S.preflog = load1(b, "rrsw_ref.preflog");
  S.tref = load1(b, "rrsw_ref.tref");
  for (auto pr : {std::make_pair(&S.extliq1, "extliq1"), std::make_pair(&S.ssaliq1, "ssaliq1"),
                  std::make_pair(&S.asyliq1, "asyliq1"), std::make_pair(&S.extice2, "extice2"),
                  std::make_pair(&S.ssaice2, "ssaice2"), std::make_pair(&S.asyice2, "asyice2"),
                  std::make_pair(&S.extice3, "extice3"), std::make_pair(&S.ssaice3, "ssaice3"),
                  std::make_pair(&S.asyice3, "asyice3"), std::make_pair(&S.fdlice3, "fdlice3")})
    *pr.first = load2(b, std::string("rrsw_cld.") + pr.second);
  S.abari = load1(b, "rrsw_cld.abari"); S.bbari = load1(b, "rrsw_cld.bbari"); S.cbari = load1(b, "rrsw_cld.cbari");
  S.dbari = load1(b, "rrsw_cld.dbari"); S.ebari = load1(b, "rrsw_cld.ebari"); S.fbari = load1(b, "rrsw_cld.fbari");
  S.rsrtaua = load2(b, "rrsw_aer.rsrtaua"); S.rsrpiza = load2(b, "rrsw_aer.rsrpiza"); S.rsrasya = load2(b, "rrsw_aer.rsrasya");
  static const char* names[] = {"kao", "kbo", "selfrefo", "forrefo", "sfluxrefo", "irradnceo", "facbrghto", "snsptdrko",
                                "raylo", "raylao", "raylbo", "abso3ao", "abso3bo", "absch4o", "absco2o", "absh2oo"};
  for (int ibnd = 1; ibnd <= 14; ++ibnd) {
    char mod[32];
    std::snprintf(mod, sizeof mod, "rrsw_kg%02d.", ibnd + 15);
    SwBand& B = S.band[ibnd];
    B.ng = ngc_[ibnd - 1];
    B.t.clear();
    B.rayl = b.has(std::string(mod) + "rayl") ? b.get(std::string(mod) + "rayl").p[0] : 0.;
    for (const char* nm : names) {
      std::string key = std::string(mod) + nm;
      if (!b.has(key)) continue;
      std::string s(nm), red;
      if (s == "kao") red = "ka";
      else if (s == "kbo") red = "kb";
      else red = s.substr(0, s.size() - 1);
      bool plain = (s == "sfluxrefo" || s == "irradnceo" || s == "facbrghto" || s == "snsptdrko");
      const BlobEntry& e = b.get(key);
      bool g_first = (plain || s == "raylao") && e.shape.size() == 2;
      if (e.shape.size() == 1) g_first = true;  // 1-D vectors over g: out(igc, 1)
      B.t[red] = reduce_g(e, ibnd, !plain, g_first);
    }
  }
  S.ready = true;
}

// ---------------------------------------------------------------------------------------------
struct SwIn {
  int ncol, nlay;
  const double *play, *plev, *tlay, *tlev, *tsfc, *h2o, *o3, *co2, *ch4, *n2o, *o2, *asdir, *asdif, *aldir, *aldif,
      *coszen, *cldfr, *taucld, *ssacld, *asmcld, *fsfcld, *cicewp, *cliqwp, *reice, *reliq, *tauaer, *ssaaer, *asmaer,
      *ecaer;
};
#define IN2(a, ip, l) in.a[(size_t)((ip)-1) + (size_t)in.ncol * ((l)-1)]

struct Col {
  int nlayers;
  A1 pavel, tavel, pz, tz, pdp, coldry;
  A2 wkl;
  double tbound;
  double adjflux[30];
  int inflag, iceflag, liqflag;
  A1 cldfrac, ciwp, clwp, rei, rel;
  A2 tauc, ssac, asmc, fsfc;              // (14, nlay)
  A2 taua, ssaa, asma;                    // (nlay, 14)
  A2 taucldorig, taucloud, ssacloud, asmcloud;  // (nlay, 16:29) stored with band offset
  double svar_f, svar_s, svar_i, svar_f_bnd[30], svar_s_bnd[30], svar_i_bnd[30];
  int laytrop, layswtch, laylow;
  std::vector<int> jp, jt, jt1, indself, indfor;
  A1 co2mult, colch4, colco2, colh2o, colmol, coln2o, colo2, colo3, fac00, fac01, fac10, fac11, selffac, selffrac,
      forfac, forfrac;
  explicit Col(int nlay) : nlayers(nlay) {
    int n = nlay + 1;
    pavel = A1(n); tavel = A1(n); pz = A1(n + 1, 0); tz = A1(n + 1, 0); pdp = A1(n); coldry = A1(n);
    wkl = A2(38, n);
    cldfrac = A1(n); ciwp = A1(n); clwp = A1(n); rei = A1(n); rel = A1(n);
    tauc = A2(14, n); ssac = A2(14, n); asmc = A2(14, n); fsfc = A2(14, n);
    taua = A2(n, 14); ssaa = A2(n, 14); asma = A2(n, 14);
    taucldorig = A2(n, 29); taucloud = A2(n, 29); ssacloud = A2(n, 29); asmcloud = A2(n, 29);
    jp.assign(n + 2, 0); jt = jp; jt1 = jp; indself = jp; indfor = jp;
    for (A1* a : {&co2mult, &colch4, &colco2, &colh2o, &colmol, &coln2o, &colo2, &colo3, &fac00, &fac01, &fac10,
                  &fac11, &selffac, &selffrac, &forfac, &forfrac})
      *a = A1(n);
  }
};

// earth_sun — rrtmg_sw_rad.nomcica.f90:818-843
static double earth_sun(int idn) {
  double gamma = 2. * S.pi * (idn - 1) / 365.;
  return 1.000110 + .034221 * std::cos(gamma) + .001289 * std::sin(gamma) + .000719 * std::cos(2. * gamma) +
         .000077 * std::sin(2. * gamma);
}

// NRLSSI2 facular / sunspot mean-cycle indices (rrtmg_sw_rad.nomcica.f90:1122-1167)
static const double mgavgcyc[132] = {
    0.150737, 0.150733, 0.150718, 0.150725, 0.150762, 0.150828, 0.150918, 0.151017, 0.151113, 0.151201, 0.151292,
    0.151403, 0.151557, 0.151766, 0.152023, 0.152322, 0.152646, 0.152969, 0.153277, 0.153579, 0.153899, 0.154252,
    0.154651, 0.155104, 0.155608, 0.156144, 0.156681, 0.157178, 0.157605, 0.157971, 0.158320, 0.158702, 0.159133,
    0.159583, 0.160018, 0.160408, 0.160725, 0.160960, 0.161131, 0.161280, 0.161454, 0.161701, 0.162034, 0.162411,
    0.162801, 0.163186, 0.163545, 0.163844, 0.164029, 0.164054, 0.163910, 0.163621, 0.163239, 0.162842, 0.162525,
    0.162344, 0.162275, 0.162288, 0.162369, 0.162500, 0.162671, 0.162878, 0.163091, 0.163251, 0.163320, 0.163287,
    0.163153, 0.162927, 0.162630, 0.162328, 0.162083, 0.161906, 0.161766, 0.161622, 0.161458, 0.161266, 0.161014,
    0.160666, 0.160213, 0.159690, 0.159190, 0.158831, 0.158664, 0.158634, 0.158605, 0.158460, 0.158152, 0.157691,
    0.157152, 0.156631, 0.156180, 0.155827, 0.155575, 0.155406, 0.155280, 0.155145, 0.154972, 0.154762, 0.154554,
    0.154388, 0.154267, 0.154152, 0.154002, 0.153800, 0.153567, 0.153348, 0.153175, 0.153044, 0.152923, 0.152793,
    0.152652, 0.152510, 0.152384, 0.152282, 0.152194, 0.152099, 0.151980, 0.151844, 0.151706, 0.151585, 0.151496,
    0.151437, 0.151390, 0.151347, 0.151295, 0.151220, 0.151115, 0.150993, 0.150883, 0.150802, 0.150752, 0.150737};
static const double sbavgcyc[132] = {
    50.3550, 52.0179, 59.2231, 66.3702, 71.7545, 76.8671, 83.4723, 91.1574, 98.4915, 105.3173, 115.1791, 130.9432,
    155.0483, 186.5379, 221.5456, 256.9212, 291.5276, 325.2953, 356.4789, 387.2470, 422.8557, 466.1698, 521.5139,
    593.2833, 676.6234, 763.6930, 849.1200, 928.4259, 994.9705, 1044.2605, 1087.5703, 1145.0623, 1224.3491, 1320.6497,
    1413.0979, 1472.1591, 1485.7531, 1464.1610, 1439.1617, 1446.2449, 1496.4323, 1577.8394, 1669.5933, 1753.0408,
    1821.9296, 1873.2789, 1906.5240, 1920.4482, 1904.6881, 1861.8397, 1802.7661, 1734.0215, 1665.0562, 1608.8999,
    1584.8208, 1594.0162, 1616.1486, 1646.6031, 1687.1962, 1736.4778, 1787.2419, 1824.9084, 1835.5236, 1810.2161,
    1768.6124, 1745.1085, 1748.7762, 1756.1239, 1738.9929, 1700.0656, 1658.2209, 1629.2925, 1620.9709, 1622.5157,
    1623.4703, 1612.3083, 1577.3031, 1516.7953, 1430.0403, 1331.5112, 1255.5171, 1226.7653, 1241.4419, 1264.6549,
    1255.5559, 1203.0286, 1120.2747, 1025.5101, 935.4602, 855.0434, 781.0189, 718.0328, 678.5850, 670.4219, 684.1906,
    697.0376, 694.8083, 674.1456, 638.8199, 602.3454, 577.6292, 565.6213, 553.7846, 531.7452, 503.9732, 476.9708,
    452.4296, 426.2826, 394.6636, 360.1086, 324.9731, 297.2957, 286.1536, 287.4195, 288.9029, 282.7594, 267.7211,
    246.6594, 224.7318, 209.2318, 204.5217, 204.1653, 200.0440, 191.0689, 175.7699, 153.9869, 128.4389, 103.8445,
    85.6083, 73.6264, 64.4393, 50.3550};

// inatm_sw — rrtmg_sw_rad.nomcica.f90:846-1539
static void inatm_sw(const SwIn& in, int iplon, int icld, int iaer, double adjes, int dyofyr, double scon, int isolvar,
                     int inflgsw, int iceflgsw, int liqflgsw, const double* bndsolvar, double* indsolvar,
                     double solcycfrac, Col& c) {
  const double amd = 28.9660, amw = 18.0160;
  const double rrsw_scon = (double)1.36822e+03f;  // default-real parameter (parrrsw.f90:115)
  const double Iint = 1360.37, Fint = 0.996047, Sint = -0.511590, Foffset = 0.14959542, Soffset = 0.00066696,
               svar_f_avg = 0.1568113, svar_s_avg = 909.21910;
  const int nsolfrac = 132;
  const int nlayers = in.nlay;
  std::fill(c.wkl.d.begin(), c.wkl.d.end(), 0.);
  std::fill(c.cldfrac.d.begin(), c.cldfrac.d.end(), 0.);
  std::fill(c.tauc.d.begin(), c.tauc.d.end(), 0.);
  std::fill(c.ssac.d.begin(), c.ssac.d.end(), 1.);
  std::fill(c.asmc.d.begin(), c.asmc.d.end(), 0.);
  std::fill(c.fsfc.d.begin(), c.fsfc.d.end(), 0.);
  std::fill(c.ciwp.d.begin(), c.ciwp.d.end(), 0.);
  std::fill(c.clwp.d.begin(), c.clwp.d.end(), 0.);
  std::fill(c.rei.d.begin(), c.rei.d.end(), 0.);
  std::fill(c.rel.d.begin(), c.rel.d.end(), 0.);
  std::fill(c.taua.d.begin(), c.taua.d.end(), 0.);
  std::fill(c.ssaa.d.begin(), c.ssaa.d.end(), 1.);
  std::fill(c.asma.d.begin(), c.asma.d.end(), 0.);
  double solvar[30];
  for (int i = 0; i < 30; ++i) { solvar[i] = 1.0; c.adjflux[i] = 1.0; c.svar_f_bnd[i] = 1.0; c.svar_s_bnd[i] = 1.0; c.svar_i_bnd[i] = 1.0; }
  c.svar_f = 1.0; c.svar_s = 1.0; c.svar_i = 1.0;
  double wgt;
  if (isolvar == 1) {
    if (indsolvar[0] != 1.0 || indsolvar[1] != 1.0) {
      if (solcycfrac >= 0.0 && solcycfrac < 0.0229) {
        wgt = (solcycfrac + 1.0 - 0.3817) / (1.0229 - 0.3817);
        indsolvar[0] = indsolvar[0] + wgt * (1.0 - indsolvar[0]);
        indsolvar[1] = indsolvar[1] + wgt * (1.0 - indsolvar[1]);
      }
      if (solcycfrac >= 0.0229 && solcycfrac <= 0.3817) {
        wgt = (solcycfrac - 0.0229) / (0.3817 - 0.0229);
        indsolvar[0] = 1.0 + wgt * (indsolvar[0] - 1.0);
        indsolvar[1] = 1.0 + wgt * (indsolvar[1] - 1.0);
      }
      if (solcycfrac > 0.3817 && solcycfrac <= 1.0) {
        wgt = (solcycfrac - 0.3817) / (1.0229 - 0.3817);
        indsolvar[0] = indsolvar[0] + wgt * (1.0 - indsolvar[0]);
        indsolvar[1] = indsolvar[1] + wgt * (1.0 - indsolvar[1]);
      }
    }
  }
  double adjflx = adjes;
  if (dyofyr > 0) adjflx = earth_sun(dyofyr);
  auto cycle_indices = [&](double& a0, double& b0) {
    if (solcycfrac <= 0.0) { a0 = mgavgcyc[0]; b0 = sbavgcyc[0]; }
    else if (solcycfrac >= 1.0) { a0 = mgavgcyc[nsolfrac - 1]; b0 = sbavgcyc[nsolfrac - 1]; }
    else {
      int sfid = (int)std::floor(solcycfrac * (nsolfrac - 1)) + 1;
      double nsfm1_inv = 1.0 / (nsolfrac - 1);
      double fraclo = (sfid - 1) * nsfm1_inv, frachi = sfid * nsfm1_inv;
      double intfrac = (solcycfrac - fraclo) / (frachi - fraclo);
      a0 = mgavgcyc[sfid - 1] + intfrac * (mgavgcyc[sfid] - mgavgcyc[sfid - 1]);
      b0 = sbavgcyc[sfid - 1] + intfrac * (sbavgcyc[sfid] - sbavgcyc[sfid - 1]);
    }
  };
  if (scon == 0.0) {
    if (isolvar == -1) for (int ib = jpb1; ib <= jpb2; ++ib) solvar[ib] = bndsolvar[ib - jpb1];
    if (isolvar == 0) { c.svar_f = 1.0; c.svar_s = 1.0; c.svar_i = 1.0; }
    if (isolvar == 1) {
      double a0, b0;
      cycle_indices(a0, b0);
      c.svar_f = indsolvar[0] * (a0 - Foffset) / (svar_f_avg - Foffset);
      c.svar_s = indsolvar[1] * (b0 - Soffset) / (svar_s_avg - Soffset);
      c.svar_i = 1.0;
    }
    if (isolvar == 2) {
      c.svar_f = (indsolvar[0] - Foffset) / (svar_f_avg - Foffset);
      c.svar_s = (indsolvar[1] - Soffset) / (svar_s_avg - Soffset);
      c.svar_i = 1.0;
    }
    if (isolvar == 3)
      for (int ib = jpb1; ib <= jpb2; ++ib) {
        solvar[ib] = bndsolvar[ib - jpb1];
        c.svar_f_bnd[ib] = solvar[ib]; c.svar_s_bnd[ib] = solvar[ib]; c.svar_i_bnd[ib] = solvar[ib];
      }
  }
  if (scon > 0.0) {
    if (isolvar == -1) for (int ib = jpb1; ib <= jpb2; ++ib) solvar[ib] = bndsolvar[ib - jpb1] * scon / rrsw_scon;
    if (isolvar == 0) {
      double svar_cprim = Fint + Sint + Iint;
      double svar_r = scon / svar_cprim;
      c.svar_f = svar_r; c.svar_s = svar_r; c.svar_i = svar_r;
    }
    if (isolvar == 1) {
      double a0, b0;
      cycle_indices(a0, b0);
      c.svar_i = (scon - (indsolvar[0] * Fint + indsolvar[1] * Sint)) / Iint;
      c.svar_f = indsolvar[0] * (a0 - Foffset) / (svar_f_avg - Foffset);
      c.svar_s = indsolvar[1] * (b0 - Soffset) / (svar_s_avg - Soffset);
    }
    // isolvar == 2 is not available for scon > 0 (commented out in the reference, :1368-1375): multipliers stay 1
    if (isolvar == 3) {
      double svar_cprim = Fint + Sint + Iint;
      for (int ib = jpb1; ib <= jpb2; ++ib) {
        solvar[ib] = bndsolvar[ib - jpb1] * scon / svar_cprim;
        c.svar_f_bnd[ib] = solvar[ib]; c.svar_s_bnd[ib] = solvar[ib]; c.svar_i_bnd[ib] = solvar[ib];
      }
    }
  }
  if (isolvar < 0) for (int ib = jpb1; ib <= jpb2; ++ib) c.adjflux[ib] = adjflx * solvar[ib];
  if (isolvar >= 0) for (int ib = jpb1; ib <= jpb2; ++ib) c.adjflux[ib] = adjflx;
  c.tbound = in.tsfc[iplon - 1];
  c.pz(0) = IN2(plev, iplon, 1);
  c.tz(0) = IN2(tlev, iplon, 1);
  for (int l = 1; l <= nlayers; ++l) {
    c.pavel(l) = IN2(play, iplon, l);
    c.tavel(l) = IN2(tlay, iplon, l);
    c.pz(l) = IN2(plev, iplon, l + 1);
    c.tz(l) = IN2(tlev, iplon, l + 1);
    c.pdp(l) = c.pz(l - 1) - c.pz(l);
    c.wkl(1, l) = IN2(h2o, iplon, l);
    c.wkl(2, l) = IN2(co2, iplon, l);
    c.wkl(3, l) = IN2(o3, iplon, l);
    c.wkl(4, l) = IN2(n2o, iplon, l);
    c.wkl(6, l) = IN2(ch4, iplon, l);
    c.wkl(7, l) = IN2(o2, iplon, l);
    double amm = (1. - c.wkl(1, l)) * amd + c.wkl(1, l) * amw;
    c.coldry(l) = (c.pz(l - 1) - c.pz(l)) * 1.e3 * S.avogad / (1.e2 * S.grav * amm * (1. + c.wkl(1, l)));
  }
  for (int l = 1; l <= nlayers; ++l)
    for (int imol = 1; imol <= 7; ++imol) c.wkl(imol, l) = c.coldry(l) * c.wkl(imol, l);
  if (iaer >= 1)
    for (int l = 1; l <= nlayers; ++l)
      for (int ib = 1; ib <= nbndsw; ++ib) {
        size_t o = (size_t)(iplon - 1) + (size_t)in.ncol * ((l - 1) + (size_t)in.nlay * (ib - 1));
        c.taua(l, ib) = in.tauaer[o];
        c.ssaa(l, ib) = in.ssaaer[o];
        c.asma(l, ib) = in.asmaer[o];
      }
  if (icld >= 1) {
    c.inflag = inflgsw; c.iceflag = iceflgsw; c.liqflag = liqflgsw;
    for (int l = 1; l <= nlayers; ++l) {
      c.cldfrac(l) = IN2(cldfr, iplon, l);
      c.ciwp(l) = IN2(cicewp, iplon, l);
      c.clwp(l) = IN2(cliqwp, iplon, l);
      c.rei(l) = IN2(reice, iplon, l);
      c.rel(l) = IN2(reliq, iplon, l);
      for (int n = 1; n <= nbndsw; ++n) {
        size_t o = (size_t)(n - 1) + 14 * ((size_t)(iplon - 1) + (size_t)in.ncol * (l - 1));
        c.tauc(n, l) = in.taucld[o];
        c.ssac(n, l) = in.ssacld[o];
        c.asmc(n, l) = in.asmcld[o];
        c.fsfc(n, l) = in.fsfcld[o];
      }
    }
  }
}

// cldprop_sw — rrtmg_sw_cldprop.f90:53-365.  Band-resolved cloud tables are declared (n, 16:29): stored column 1..14.
static int cldprop_sw(Col& c, std::string& err) {
  const double eps = 1.e-06, cldmin = 1.e-20;
  const int nlayers = c.nlayers;
  double extcoice[30] = {0}, gice[30] = {0}, ssacoice[30] = {0}, forwice[30] = {0}, extcoliq[30] = {0}, gliq[30] = {0},
         ssacoliq[30] = {0}, forwliq[30] = {0}, fdelta[30] = {0};
  std::vector<double> tauctot(nlayers + 2, 0.);
  for (int lay = 1; lay <= nlayers; ++lay)
    for (int ib = jpb1; ib <= jpb2; ++ib) {
      c.taucldorig(lay, ib) = c.tauc(ib - 15, lay);
      c.taucloud(lay, ib) = 0.0;
      c.ssacloud(lay, ib) = 1.0;
      c.asmcloud(lay, ib) = 0.0;
      tauctot[lay] = tauctot[lay] + c.tauc(ib - 15, lay);
    }
  for (int lay = 1; lay <= nlayers; ++lay) {
    double cwp = c.ciwp(lay) + c.clwp(lay);
    if (c.cldfrac(lay) >= cldmin && (cwp >= cldmin || tauctot[lay] >= cldmin)) {
      if (c.inflag == 0) {
        for (int ib = jpb1; ib <= jpb2; ++ib) {
          double taucldorig_a = c.tauc(ib - 15, lay);
          double ffp = c.fsfc(ib - 15, lay);
          double ffp1 = 1.0 - ffp;
          double ffpssa = 1.0 - ffp * c.ssac(ib - 15, lay);
          double ssacloud_a = ffp1 * c.ssac(ib - 15, lay) / ffpssa;
          double taucloud_a = ffpssa * taucldorig_a;
          c.taucldorig(lay, ib) = taucldorig_a;
          c.ssacloud(lay, ib) = ssacloud_a;
          c.taucloud(lay, ib) = taucloud_a;
          c.asmcloud(lay, ib) = (c.asmc(ib - 15, lay) - ffp) / (ffp1);
        }
      } else if (c.inflag == 2) {
        double radice = c.rei(lay);
        if (c.ciwp(lay) == 0.0) {
          for (int ib = jpb1; ib <= jpb2; ++ib) { extcoice[ib] = 0.0; ssacoice[ib] = 0.0; gice[ib] = 0.0; forwice[ib] = 0.0; }
        } else if (c.iceflag == 1) {
          if (radice < 13.0 || radice > 130.) { err = "ICE RADIUS OUT OF BOUNDS"; return 1; }
          for (int ib = jpb1; ib <= jpb2; ++ib) {
            int icx = 5;
            double w2 = wavenum2_[ib - 16];
            if (w2 > 1.43e04) icx = 1;
            else if (w2 > 7.7e03) icx = 2;
            else if (w2 > 5.3e03) icx = 3;
            else if (w2 > 4.0e03) icx = 4;
            else if (w2 >= 2.5e03) icx = 5;
            extcoice[ib] = S.abari(icx) + S.bbari(icx) / radice;
            ssacoice[ib] = 1. - S.cbari(icx) - S.dbari(icx) * radice;
            gice[ib] = S.ebari(icx) + S.fbari(icx) * radice;
            if (gice[ib] >= 1.0) gice[ib] = 1.0 - eps;
            forwice[ib] = gice[ib] * gice[ib];
            if (extcoice[ib] < 0.0) { err = "ICE EXTINCTION LESS THAN 0.0"; return 1; }
            if (ssacoice[ib] > 1.0) { err = "ICE SSA GRTR THAN 1.0"; return 1; }
            if (ssacoice[ib] < 0.0) { err = "ICE SSA LESS THAN 0.0"; return 1; }
            if (gice[ib] > 1.0) { err = "ICE ASYM GRTR THAN 1.0"; return 1; }
            if (gice[ib] < 0.0) { err = "ICE ASYM LESS THAN 0.0"; return 1; }
          }
        } else if (c.iceflag == 2) {
          if (radice < 5.0 || radice > 131.0) { err = "ICE RADIUS OUT OF BOUNDS"; return 1; }
          double factor = (radice - 2.) / 3.;
          int index = (int)factor;
          if (index == 43) index = 42;
          double fint = factor - (double)index;
          for (int ib = jpb1; ib <= jpb2; ++ib) {
            int k = ib - 15;
            extcoice[ib] = S.extice2(index, k) + fint * (S.extice2(index + 1, k) - S.extice2(index, k));
            ssacoice[ib] = S.ssaice2(index, k) + fint * (S.ssaice2(index + 1, k) - S.ssaice2(index, k));
            gice[ib] = S.asyice2(index, k) + fint * (S.asyice2(index + 1, k) - S.asyice2(index, k));
            forwice[ib] = gice[ib] * gice[ib];
            if (extcoice[ib] < 0.0) { err = "ICE EXTINCTION LESS THAN 0.0"; return 1; }
            if (ssacoice[ib] > 1.0) { err = "ICE SSA GRTR THAN 1.0"; return 1; }
            if (ssacoice[ib] < 0.0) { err = "ICE SSA LESS THAN 0.0"; return 1; }
            if (gice[ib] > 1.0) { err = "ICE ASYM GRTR THAN 1.0"; return 1; }
            if (gice[ib] < 0.0) { err = "ICE ASYM LESS THAN 0.0"; return 1; }
          }
        } else if (c.iceflag == 3) {
          if (radice < 5.0 || radice > 140.0) { err = "ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; return 1; }
          double factor = (radice - 2.) / 3.;
          int index = (int)factor;
          if (index == 46) index = 45;
          double fint = factor - (double)index;
          for (int ib = jpb1; ib <= jpb2; ++ib) {
            int k = ib - 15;
            extcoice[ib] = S.extice3(index, k) + fint * (S.extice3(index + 1, k) - S.extice3(index, k));
            ssacoice[ib] = S.ssaice3(index, k) + fint * (S.ssaice3(index + 1, k) - S.ssaice3(index, k));
            gice[ib] = S.asyice3(index, k) + fint * (S.asyice3(index + 1, k) - S.asyice3(index, k));
            fdelta[ib] = S.fdlice3(index, k) + fint * (S.fdlice3(index + 1, k) - S.fdlice3(index, k));
            if (fdelta[ib] < 0.0) { err = "FDELTA LESS THAN 0.0"; return 1; }
            if (fdelta[ib] > 1.0) { err = "FDELTA GT THAN 1.0"; return 1; }
            forwice[ib] = fdelta[ib] + 0.5 / ssacoice[ib];
            if (forwice[ib] > gice[ib]) forwice[ib] = gice[ib];
            if (extcoice[ib] < 0.0) { err = "ICE EXTINCTION LESS THAN 0.0"; return 1; }
            if (ssacoice[ib] > 1.0) { err = "ICE SSA GRTR THAN 1.0"; return 1; }
            if (ssacoice[ib] < 0.0) { err = "ICE SSA LESS THAN 0.0"; return 1; }
            if (gice[ib] > 1.0) { err = "ICE ASYM GRTR THAN 1.0"; return 1; }
            if (gice[ib] < 0.0) { err = "ICE ASYM LESS THAN 0.0"; return 1; }
          }
        }
        if (c.clwp(lay) == 0.0) {
          for (int ib = jpb1; ib <= jpb2; ++ib) { extcoliq[ib] = 0.0; ssacoliq[ib] = 0.0; gliq[ib] = 0.0; forwliq[ib] = 0.0; }
        } else if (c.liqflag == 1) {
          double radliq = c.rel(lay);
          if (radliq < 2.5 || radliq > 60.) { err = "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS"; return 1; }
          int index = (int)(radliq - 1.5);
          if (index == 0) index = 1;
          if (index == 58) index = 57;
          double fint = radliq - 1.5 - (double)index;
          for (int ib = jpb1; ib <= jpb2; ++ib) {
            int k = ib - 15;
            extcoliq[ib] = S.extliq1(index, k) + fint * (S.extliq1(index + 1, k) - S.extliq1(index, k));
            ssacoliq[ib] = S.ssaliq1(index, k) + fint * (S.ssaliq1(index + 1, k) - S.ssaliq1(index, k));
            if (fint < 0. && ssacoliq[ib] > 1.) ssacoliq[ib] = S.ssaliq1(index, k);
            gliq[ib] = S.asyliq1(index, k) + fint * (S.asyliq1(index + 1, k) - S.asyliq1(index, k));
            forwliq[ib] = gliq[ib] * gliq[ib];
            if (extcoliq[ib] < 0.0) { err = "LIQUID EXTINCTION LESS THAN 0.0"; return 1; }
            if (ssacoliq[ib] > 1.0) { err = "LIQUID SSA GRTR THAN 1.0"; return 1; }
            if (ssacoliq[ib] < 0.0) { err = "LIQUID SSA LESS THAN 0.0"; return 1; }
            if (gliq[ib] > 1.0) { err = "LIQUID ASYM GRTR THAN 1.0"; return 1; }
            if (gliq[ib] < 0.0) { err = "LIQUID ASYM LESS THAN 0.0"; return 1; }
          }
        }
        for (int ib = jpb1; ib <= jpb2; ++ib) {
          double tauliqorig = c.clwp(lay) * extcoliq[ib];
          double tauiceorig = c.ciwp(lay) * extcoice[ib];
          c.taucldorig(lay, ib) = tauliqorig + tauiceorig;
          double ssaliq = ssacoliq[ib] * (1.0 - forwliq[ib]) / (1.0 - forwliq[ib] * ssacoliq[ib]);
          double tauliq = (1.0 - forwliq[ib] * ssacoliq[ib]) * tauliqorig;
          double ssaice = ssacoice[ib] * (1.0 - forwice[ib]) / (1.0 - forwice[ib] * ssacoice[ib]);
          double tauice = (1.0 - forwice[ib] * ssacoice[ib]) * tauiceorig;
          double scatliq = ssaliq * tauliq;
          double scatice = ssaice * tauice;
          c.taucloud(lay, ib) = tauliq + tauice;
          if (c.taucloud(lay, ib) == 0.0) c.taucloud(lay, ib) = cldmin;
          if (scatice == 0.0) scatice = cldmin;
          c.ssacloud(lay, ib) = (scatliq + scatice) / c.taucloud(lay, ib);
          if (c.iceflag == 3) {
            c.asmcloud(lay, ib) = (1.0 / (scatliq + scatice)) *
                                  (scatliq * (gliq[ib] - forwliq[ib]) / (1.0 - forwliq[ib]) +
                                   scatice * ((gice[ib] - forwice[ib]) / (1.0 - forwice[ib])));
          } else {
            c.asmcloud(lay, ib) = (scatliq * (gliq[ib] - forwliq[ib]) / (1.0 - forwliq[ib]) +
                                   scatice * (gice[ib] - forwice[ib]) / (1.0 - forwice[ib])) /
                                  (scatliq + scatice);
          }
        }
      }
    }
  }
  return 0;
}

// setcoef_sw — rrtmg_sw_setcoef.f90:49-305
static void setcoef_sw(Col& c) {
  const int nlayers = c.nlayers;
  const double stpfac = 296. / 1013.;
  c.laytrop = 0; c.layswtch = 0; c.laylow = 0;
  for (int lay = 1; lay <= nlayers; ++lay) {
    double plog = std::log(c.pavel(lay));
    c.jp[lay] = (int)(36. - 5 * (plog + 0.04));
    if (c.jp[lay] < 1) c.jp[lay] = 1; else if (c.jp[lay] > 58) c.jp[lay] = 58;
    int jp1 = c.jp[lay] + 1;
    double fp = 5. * (S.preflog(c.jp[lay]) - plog);
    c.jt[lay] = (int)(3. + (c.tavel(lay) - S.tref(c.jp[lay])) / 15.);
    if (c.jt[lay] < 1) c.jt[lay] = 1; else if (c.jt[lay] > 4) c.jt[lay] = 4;
    double ft = ((c.tavel(lay) - S.tref(c.jp[lay])) / 15.) - (double)(c.jt[lay] - 3);
    c.jt1[lay] = (int)(3. + (c.tavel(lay) - S.tref(jp1)) / 15.);
    if (c.jt1[lay] < 1) c.jt1[lay] = 1; else if (c.jt1[lay] > 4) c.jt1[lay] = 4;
    double ft1 = ((c.tavel(lay) - S.tref(jp1)) / 15.) - (double)(c.jt1[lay] - 3);
    double water = c.wkl(1, lay) / c.coldry(lay);
    double scalefac = c.pavel(lay) * stpfac / c.tavel(lay);
    double factor;
    if (!(plog <= 4.56)) {
      c.laytrop = c.laytrop + 1;
      if (plog >= 6.62) c.laylow = c.laylow + 1;
      c.forfac(lay) = scalefac / (1. + water);
      factor = (332.0 - c.tavel(lay)) / 36.0;
      c.indfor[lay] = std::min(2, std::max(1, (int)factor));
      c.forfrac(lay) = factor - (double)c.indfor[lay];
      c.selffac(lay) = water * c.forfac(lay);
      factor = (c.tavel(lay) - 188.0) / 7.2;
      c.indself[lay] = std::min(9, std::max(1, (int)factor - 7));
      c.selffrac(lay) = factor - (double)(c.indself[lay] + 7);
    } else {
      c.forfac(lay) = scalefac / (1. + water);
      factor = (c.tavel(lay) - 188.0) / 36.0;
      c.indfor[lay] = 3;
      c.forfrac(lay) = factor - 1.0;
    }
    c.colh2o(lay) = 1.e-20 * c.wkl(1, lay);
    c.colco2(lay) = 1.e-20 * c.wkl(2, lay);
    c.colo3(lay) = 1.e-20 * c.wkl(3, lay);
    c.coln2o(lay) = 1.e-20 * c.wkl(4, lay);
    c.colch4(lay) = 1.e-20 * c.wkl(6, lay);
    c.colo2(lay) = 1.e-20 * c.wkl(7, lay);
    c.colmol(lay) = 1.e-20 * c.coldry(lay) + c.colh2o(lay);
    if (c.colco2(lay) == 0.) c.colco2(lay) = 1.e-32 * c.coldry(lay);
    if (c.coln2o(lay) == 0.) c.coln2o(lay) = 1.e-32 * c.coldry(lay);
    if (c.colch4(lay) == 0.) c.colch4(lay) = 1.e-32 * c.coldry(lay);
    if (c.colo2(lay) == 0.) c.colo2(lay) = 1.e-32 * c.coldry(lay);
    double co2reg = 3.55e-24 * c.coldry(lay);
    c.co2mult(lay) = (c.colco2(lay) - co2reg) * 272.63 * std::exp(-1919.4 / c.tavel(lay)) / (8.7604e-4 * c.tavel(lay));
    if (plog <= 4.56) {
      c.selffac(lay) = 0.;
      c.selffrac(lay) = 0.;
      c.indself[lay] = 0;
    }
    double compfp = 1. - fp;
    c.fac10(lay) = compfp * ft;
    c.fac00(lay) = compfp * (1. - ft);
    c.fac11(lay) = fp * ft1;
    c.fac01(lay) = fp * (1. - ft1);
  }
}

// ---------------------------------------------------------------------------------------------
// taumol_sw — rrtmg_sw_taumol.f90:50-1790
struct Spectral {
  A2 taug, taur;         // (nlay, 112)
  double ssi[113], sfluxzen[113];
  explicit Spectral(int nlay) : taug(nlay + 1, 112), taur(nlay + 1, 112) {
    for (int i = 0; i < 113; ++i) { ssi[i] = 0.; sfluxzen[i] = 0.; }
  }
};

static void taumol_sw(Col& c, int isolvar, Spectral& sp) {
  const int nlayers = c.nlayers, laytrop = c.laytrop;
  const double oneminus = S.oneminus;
  auto ind0a = [&](int lay, int ib) { return ((c.jp[lay] - 1) * 5 + (c.jt[lay] - 1)) * nspa_[ib - 1]; };
  auto ind1a = [&](int lay, int ib) { return (c.jp[lay] * 5 + (c.jt1[lay] - 1)) * nspa_[ib - 1]; };
  auto ind0b = [&](int lay, int ib) { return ((c.jp[lay] - 13) * 5 + (c.jt[lay] - 1)) * nspb_[ib - 1]; };
  auto ind1b = [&](int lay, int ib) { return ((c.jp[lay] - 12) * 5 + (c.jt1[lay] - 1)) * nspb_[ib - 1]; };
  struct Bin { double speccomb, fs; int js; double f[8]; };
  auto binspec = [&](double cola, double strrat, double colb, double n, int lay) {
    Bin b;
    b.speccomb = cola + strrat * colb;
    double specparm = cola / b.speccomb;
    if (specparm >= oneminus) specparm = oneminus;
    double specmult = n * specparm;
    b.js = 1 + (int)specmult;
    b.fs = std::fmod(specmult, 1.);
    b.f[0] = (1. - b.fs) * c.fac00(lay);  // fac000
    b.f[1] = b.fs * c.fac00(lay);         // fac100
    b.f[2] = (1. - b.fs) * c.fac10(lay);  // fac010
    b.f[3] = b.fs * c.fac10(lay);         // fac110
    b.f[4] = (1. - b.fs) * c.fac01(lay);  // fac001
    b.f[5] = b.fs * c.fac01(lay);         // fac101
    b.f[6] = (1. - b.fs) * c.fac11(lay);  // fac011
    b.f[7] = b.fs * c.fac11(lay);         // fac111
    return b;
  };
  auto major8 = [&](const Bin& b, const A2& a, int ind0, int ind1, int nsp, int ig) {
    return b.f[0] * a(ind0, ig) + b.f[1] * a(ind0 + 1, ig) + b.f[2] * a(ind0 + nsp, ig) + b.f[3] * a(ind0 + nsp + 1, ig) +
           b.f[4] * a(ind1, ig) + b.f[5] * a(ind1 + 1, ig) + b.f[6] * a(ind1 + nsp, ig) + b.f[7] * a(ind1 + nsp + 1, ig);
  };
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
  // solar source at layer laysolfr: 1-D tables (js=0) or interpolated in js
  auto solar = [&](const SwBand& B, int ibm, int gabs, int ig, int js, double fs, double scale) {
    const A2 &sfl = B["sfluxref"], &irr = B["irradnce"], &fac = B["facbrght"], &sns = B["snsptdrk"];
    auto val = [&](const A2& t) { return js == 0 ? t(ig, 1) : t(ig, js) + fs * (t(ig, js + 1) - t(ig, js)); };
    sp.sfluxzen[gabs] = js == 0 ? scale * sfl(ig, 1) : val(sfl);
    if (isolvar >= 0 && isolvar <= 2) sp.ssi[gabs] = c.svar_f * val(fac) + c.svar_s * val(sns) + c.svar_i * val(irr);
    if (isolvar == 3)
      sp.ssi[gabs] = c.svar_f_bnd[ibm + 15] * val(fac) + c.svar_s_bnd[ibm + 15] * val(sns) + c.svar_i_bnd[ibm + 15] * val(irr);
  };
  auto laysol_lower = [&](int layreffr) {
    int laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; ++lay)
      if (c.jp[lay] < layreffr && c.jp[lay + 1] >= layreffr) laysolfr = std::min(lay + 1, laytrop);
    return laysolfr;
  };
  auto laysol_upper = [&](int layreffr) {
    int laysolfr = nlayers;
    for (int lay = laytrop + 1; lay <= nlayers; ++lay)
      if (c.jp[lay - 1] < layreffr && c.jp[lay] >= layreffr) laysolfr = lay;
    return laysolfr;
  };
  // NOTE on laysolfr: the Fortran updates laysolfr inside the layer loop and assigns the solar source when
  // `lay == laysolfr`; the last assignment wins, which is the layer returned by laysol_lower/upper (taumol.f90:586-590).

  // ---- band 16: 2600-3250 (low - h2o,ch4; high - ch4) :275-389
  {
    const SwBand& B = S.band[1];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const double strrat1 = 252.131;
    const int layreffr = 18, ngs = 0;
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin b = binspec(c.colh2o(lay), strrat1, c.colch4(lay), 8., lay);
      int ind0 = ind0a(lay, 1) + b.js, ind1 = ind1a(lay, 1) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absa, ind0, ind1, 9, ig) +
                                 c.colh2o(lay) * (tself(selfref, lay, ig) + tfor(forref, lay, ig));
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    int laysolfr = laysol_upper(layreffr);
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 1) + 1, ind1 = ind1b(lay, 1) + 1;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = c.colch4(lay) * simple4(absb, ind0, ind1, lay, ig);
        if (lay == laysolfr) solar(B, 1, ngs + ig, ig, 0, 0., 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- band 17: 3250-4000 (low - h2o,co2; high - h2o,co2) :392-531
  {
    const SwBand& B = S.band[2];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const double strrat = 0.364641;
    const int layreffr = 30, ngs = 6;
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin b = binspec(c.colh2o(lay), strrat, c.colco2(lay), 8., lay);
      int ind0 = ind0a(lay, 2) + b.js, ind1 = ind1a(lay, 2) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absa, ind0, ind1, 9, ig) +
                                 c.colh2o(lay) * (tself(selfref, lay, ig) + tfor(forref, lay, ig));
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    int laysolfr = laysol_upper(layreffr);
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      Bin b = binspec(c.colh2o(lay), strrat, c.colco2(lay), 4., lay);
      int ind0 = ind0b(lay, 2) + b.js, ind1 = ind1b(lay, 2) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absb, ind0, ind1, 5, ig) + c.colh2o(lay) * tfor(forref, lay, ig);
        if (lay == laysolfr) solar(B, 2, ngs + ig, ig, b.js, b.fs, 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- bands 18 (h2o,ch4 / ch4), 19 (h2o,co2 / co2): :534-782
  for (int ibm = 3; ibm <= 4; ++ibm) {
    const SwBand& B = S.band[ibm];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const double strrat = ibm == 3 ? 38.9589 : 5.49281;
    const int layreffr = ibm == 3 ? 6 : 3, ngs = ngs_[ibm - 2];
    const A1& colb = ibm == 3 ? c.colch4 : c.colco2;
    int laysolfr = laysol_lower(layreffr);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin b = binspec(c.colh2o(lay), strrat, colb(lay), 8., lay);
      int ind0 = ind0a(lay, ibm) + b.js, ind1 = ind1a(lay, ibm) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absa, ind0, ind1, 9, ig) +
                                 c.colh2o(lay) * (tself(selfref, lay, ig) + tfor(forref, lay, ig));
        if (lay == laysolfr) solar(B, ibm, ngs + ig, ig, b.js, b.fs, 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, ibm) + 1, ind1 = ind1b(lay, ibm) + 1;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = colb(lay) * simple4(absb, ind0, ind1, lay, ig);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- band 20: 5150-6150 (h2o / h2o; + ch4) :785-869
  {
    const SwBand& B = S.band[5];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"], &absch4 = B["absch4"];
    const int layreffr = 3, ngs = 34;
    int laysolfr = laysol_lower(layreffr);
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 5) + 1, ind1 = ind1a(lay, 5) + 1;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = c.colh2o(lay) * ((simple4(absa, ind0, ind1, lay, ig)) + tself(selfref, lay, ig) +
                                                  tfor(forref, lay, ig)) +
                                 c.colch4(lay) * absch4(ig, 1);
        if (lay == laysolfr) solar(B, 5, ngs + ig, ig, 0, 0., 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 5) + 1, ind1 = ind1b(lay, 5) + 1;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) =
            c.colh2o(lay) * (c.fac00(lay) * absb(ind0, ig) + c.fac10(lay) * absb(ind0 + 1, ig) +
                             c.fac01(lay) * absb(ind1, ig) + c.fac11(lay) * absb(ind1 + 1, ig) + tfor(forref, lay, ig)) +
            c.colch4(lay) * absch4(ig, 1);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- band 21: 6150-7700 (h2o,co2 / h2o,co2) :872-1009
  {
    const SwBand& B = S.band[6];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const double strrat = 0.0045321;
    const int layreffr = 8, ngs = 44;
    int laysolfr = laysol_lower(layreffr);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin b = binspec(c.colh2o(lay), strrat, c.colco2(lay), 8., lay);
      int ind0 = ind0a(lay, 6) + b.js, ind1 = ind1a(lay, 6) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absa, ind0, ind1, 9, ig) +
                                 c.colh2o(lay) * (tself(selfref, lay, ig) + tfor(forref, lay, ig));
        if (lay == laysolfr) solar(B, 6, ngs + ig, ig, b.js, b.fs, 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      Bin b = binspec(c.colh2o(lay), strrat, c.colco2(lay), 4., lay);
      int ind0 = ind0b(lay, 6) + b.js, ind1 = ind1b(lay, 6) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absb, ind0, ind1, 5, ig) + c.colh2o(lay) * tfor(forref, lay, ig);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- band 22: 7700-8050 (h2o,o2 / o2) :1012-1135
  {
    const SwBand& B = S.band[7];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"];
    const double o2adj = 1.6, strrat = 0.022708;
    const int layreffr = 2, ngs = 54;
    int laysolfr = laysol_lower(layreffr);
    for (int lay = 1; lay <= laytrop; ++lay) {
      double o2cont = 4.35e-4 * c.colo2(lay) / (350.0 * 2.0);
      Bin b = binspec(c.colh2o(lay), o2adj * strrat, c.colo2(lay), 8., lay);
      int ind0 = ind0a(lay, 7) + b.js, ind1 = ind1a(lay, 7) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absa, ind0, ind1, 9, ig) +
                                 c.colh2o(lay) * (tself(selfref, lay, ig) + tfor(forref, lay, ig)) + o2cont;
        if (lay == laysolfr) solar(B, 7, ngs + ig, ig, b.js, b.fs, 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      double o2cont = 4.35e-4 * c.colo2(lay) / (350.0 * 2.0);
      int ind0 = ind0b(lay, 7) + 1, ind1 = ind1b(lay, 7) + 1;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = c.colo2(lay) * o2adj * simple4(absb, ind0, ind1, lay, ig) + o2cont;
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- band 23: 8050-12850 (h2o / nothing) :1138-1216
  {
    const SwBand& B = S.band[8];
    const A2 &absa = B["ka"], &selfref = B["selfref"], &forref = B["forref"], &rayl = B["rayl"];
    const double givfac = 1.029;
    const int layreffr = 6, ngs = 56;
    int laysolfr = laysol_lower(layreffr);
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 8) + 1, ind1 = ind1a(lay, 8) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauray = c.colmol(lay) * rayl(ig, 1);
        sp.taug(lay, ngs + ig) = c.colh2o(lay) * (givfac * simple4(absa, ind0, ind1, lay, ig) + tself(selfref, lay, ig) +
                                                  tfor(forref, lay, ig));
        if (lay == laysolfr) solar(B, 8, ngs + ig, ig, 0, 0., 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay)
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = 0.;
        sp.taur(lay, ngs + ig) = c.colmol(lay) * rayl(ig, 1);
      }
  }
  // ---- band 24: 12850-16000 (h2o,o2 / o2) :1219-1344
  {
    const SwBand& B = S.band[9];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"], &abso3a = B["abso3a"],
             &abso3b = B["abso3b"], &rayla = B["rayla"], &raylb = B["raylb"];
    const double strrat = 0.124692;
    const int layreffr = 1, ngs = 66;
    int laysolfr = laysol_lower(layreffr);
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin b = binspec(c.colh2o(lay), strrat, c.colo2(lay), 8., lay);
      int ind0 = ind0a(lay, 9) + b.js, ind1 = ind1a(lay, 9) + b.js;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauray = c.colmol(lay) * (rayla(ig, b.js) + b.fs * (rayla(ig, b.js + 1) - rayla(ig, b.js)));
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absa, ind0, ind1, 9, ig) + c.colo3(lay) * abso3a(ig, 1) +
                                 c.colh2o(lay) * (tself(selfref, lay, ig) + tfor(forref, lay, ig));
        if (lay == laysolfr) solar(B, 9, ngs + ig, ig, b.js, b.fs, 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 9) + 1, ind1 = ind1b(lay, 9) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauray = c.colmol(lay) * raylb(ig, 1);
        sp.taug(lay, ngs + ig) = c.colo2(lay) * simple4(absb, ind0, ind1, lay, ig) + c.colo3(lay) * abso3b(ig, 1);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- band 25: 16000-22650 (h2o / nothing; + o3) :1347-1424
  {
    const SwBand& B = S.band[10];
    const A2 &absa = B["ka"], &abso3a = B["abso3a"], &abso3b = B["abso3b"], &rayl = B["rayl"];
    const int layreffr = 2, ngs = 74;
    int laysolfr = laysol_lower(layreffr);
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 10) + 1, ind1 = ind1a(lay, 10) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauray = c.colmol(lay) * rayl(ig, 1);
        sp.taug(lay, ngs + ig) = c.colh2o(lay) * simple4(absa, ind0, ind1, lay, ig) + c.colo3(lay) * abso3a(ig, 1);
        if (lay == laysolfr) solar(B, 10, ngs + ig, ig, 0, 0., 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay)
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauray = c.colmol(lay) * rayl(ig, 1);
        sp.taug(lay, ngs + ig) = c.colo3(lay) * abso3b(ig, 1);
        sp.taur(lay, ngs + ig) = tauray;
      }
  }
  // ---- band 26: 22650-29000 (nothing) :1427-1489
  {
    const SwBand& B = S.band[11];
    const A2& rayl = B["rayl"];
    const int ngs = 80;
    int laysolfr = laytrop;
    for (int lay = 1; lay <= laytrop; ++lay)
      for (int ig = 1; ig <= B.ng; ++ig) {
        if (lay == laysolfr) solar(B, 11, ngs + ig, ig, 0, 0., 1.);
        sp.taug(lay, ngs + ig) = 0.;
        sp.taur(lay, ngs + ig) = c.colmol(lay) * rayl(ig, 1);
      }
    for (int lay = laytrop + 1; lay <= nlayers; ++lay)
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = 0.;
        sp.taur(lay, ngs + ig) = c.colmol(lay) * rayl(ig, 1);
      }
  }
  // ---- band 27: 29000-38000 (o3 / o3) :1492-1575
  {
    const SwBand& B = S.band[12];
    const A2 &absa = B["ka"], &absb = B["kb"], &rayl = B["rayl"];
    const double scalekur = 50.15 / 48.37;
    const int layreffr = 32, ngs = 86;
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 12) + 1, ind1 = ind1a(lay, 12) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauray = c.colmol(lay) * rayl(ig, 1);
        sp.taug(lay, ngs + ig) = c.colo3(lay) * simple4(absa, ind0, ind1, lay, ig);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    int laysolfr = laysol_upper(layreffr);
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 12) + 1, ind1 = ind1b(lay, 12) + 1;
      for (int ig = 1; ig <= B.ng; ++ig) {
        double tauray = c.colmol(lay) * rayl(ig, 1);
        sp.taug(lay, ngs + ig) = c.colo3(lay) * simple4(absb, ind0, ind1, lay, ig);
        if (lay == laysolfr) solar(B, 12, ngs + ig, ig, 0, 0., scalekur);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- band 28: 38000-50000 (o3,o2 / o3,o2) :1578-1692
  {
    const SwBand& B = S.band[13];
    const A2 &absa = B["ka"], &absb = B["kb"];
    const double strrat = 6.67029e-07;
    const int layreffr = 58, ngs = 94;
    for (int lay = 1; lay <= laytrop; ++lay) {
      Bin b = binspec(c.colo3(lay), strrat, c.colo2(lay), 8., lay);
      int ind0 = ind0a(lay, 13) + b.js, ind1 = ind1a(lay, 13) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absa, ind0, ind1, 9, ig);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    int laysolfr = laysol_upper(layreffr);
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      Bin b = binspec(c.colo3(lay), strrat, c.colo2(lay), 4., lay);
      int ind0 = ind0b(lay, 13) + b.js, ind1 = ind1b(lay, 13) + b.js;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = b.speccomb * major8(b, absb, ind0, ind1, 5, ig);
        if (lay == laysolfr) solar(B, 13, ngs + ig, ig, b.js, b.fs, 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
  // ---- band 29: 820-2600 (h2o / co2; + co2, h2o minor) :1695-1787
  {
    const SwBand& B = S.band[14];
    const A2 &absa = B["ka"], &absb = B["kb"], &selfref = B["selfref"], &forref = B["forref"], &absco2 = B["absco2"],
             &absh2o = B["absh2o"];
    const int layreffr = 49, ngs = 100;
    for (int lay = 1; lay <= laytrop; ++lay) {
      int ind0 = ind0a(lay, 14) + 1, ind1 = ind1a(lay, 14) + 1;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = c.colh2o(lay) * ((simple4(absa, ind0, ind1, lay, ig)) + tself(selfref, lay, ig) +
                                                  tfor(forref, lay, ig)) +
                                 c.colco2(lay) * absco2(ig, 1);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
    int laysolfr = laysol_upper(layreffr);
    for (int lay = laytrop + 1; lay <= nlayers; ++lay) {
      int ind0 = ind0b(lay, 14) + 1, ind1 = ind1b(lay, 14) + 1;
      double tauray = c.colmol(lay) * B.rayl;
      for (int ig = 1; ig <= B.ng; ++ig) {
        sp.taug(lay, ngs + ig) = c.colco2(lay) * simple4(absb, ind0, ind1, lay, ig) + c.colh2o(lay) * absh2o(ig, 1);
        if (lay == laysolfr) solar(B, 14, ngs + ig, ig, 0, 0., 1.);
        sp.taur(lay, ngs + ig) = tauray;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// reftra_sw — rrtmg_sw_reftra.f90:48-324 (kmodts = 2, PIFM)
static void reftra_sw(int nlayers, const std::vector<char>& lrtchk, const A1& pgg, double prmuz, const A1& ptau,
                      const A1& pw, A1& pref, A1& prefd, A1& ptra, A1& ptrad) {
  const double eps = 1.e-08, od_lo = 0.06, tblint = 10000.0, bpade = S.bpade;
  const std::vector<double>& exp_tbl = S.exp_tbl;
  const double zwcrit = 0.9999995;
  for (int jk = 1; jk <= nlayers; ++jk) {
    if (!lrtchk[jk]) {
      pref(jk) = 0.; ptra(jk) = 1.; prefd(jk) = 0.; ptrad(jk) = 1.;
    } else {
      double zto1 = ptau(jk), zw = pw(jk), zg = pgg(jk);
      double zg3 = 3. * zg;
      double zgamma1 = (8. - zw * (5. + zg3)) * 0.25;
      double zgamma2 = 3. * (zw * (1. - zg)) * 0.25;
      double zgamma3 = (2. - zg3 * prmuz) * 0.25;
      double zgamma4 = 1. - zgamma3;
      double r = zg / (1. - zg);
      double zwo = zw / (1. - (1. - zw) * (r * r));
      if (zwo >= zwcrit) {
        double za = zgamma1 * prmuz;
        double za1 = za - zgamma3;
        double zgt = zgamma1 * zto1;
        double ze1 = std::min(zto1 / prmuz, 500.);
        double ze2;
        if (ze1 <= od_lo) ze2 = 1. - ze1 + 0.5 * ze1 * ze1;
        else {
          double tblind = ze1 / (bpade + ze1);
          int itind = (int)(tblint * tblind + 0.5);
          ze2 = exp_tbl[itind];
        }
        pref(jk) = (zgt - za1 * (1. - ze2)) / (1. + zgt);
        ptra(jk) = 1. - pref(jk);
        prefd(jk) = zgt / (1. + zgt);
        ptrad(jk) = 1. - prefd(jk);
        if (ze2 == 1.0) { pref(jk) = 0.0; ptra(jk) = 1.0; prefd(jk) = 0.0; ptrad(jk) = 1.0; }
      } else {
        double za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3;
        double za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4;
        double zrk = std::sqrt(zgamma1 * zgamma1 - zgamma2 * zgamma2);
        double zrp = zrk * prmuz;
        double zrp1 = 1. + zrp, zrm1 = 1. - zrp, zrk2 = 2. * zrk, zrpp = 1. - zrp * zrp, zrkg = zrk + zgamma1;
        double zr1 = zrm1 * (za2 + zrk * zgamma3);
        double zr2 = zrp1 * (za2 - zrk * zgamma3);
        double zr3 = zrk2 * (zgamma3 - za2 * prmuz);
        double zr4 = zrpp * zrkg;
        double zr5 = zrpp * (zrk - zgamma1);
        double zt1 = zrp1 * (za1 + zrk * zgamma4);
        double zt2 = zrm1 * (za1 - zrk * zgamma4);
        double zt3 = zrk2 * (zgamma4 + za1 * prmuz);
        double zt4 = zr4, zt5 = zr5;
        double zbeta = (zgamma1 - zrk) / zrkg;
        double ze1 = std::min(zrk * zto1, 500.);
        double ze2 = std::min(zto1 / prmuz, 500.);
        double zem1, zep1, zem2, zep2;
        if (ze1 <= od_lo) { zem1 = 1. - ze1 + 0.5 * ze1 * ze1; zep1 = 1. / zem1; }
        else {
          double tblind = ze1 / (bpade + ze1);
          int itind = (int)(tblint * tblind + 0.5);
          zem1 = exp_tbl[itind];
          zep1 = 1. / zem1;
        }
        if (ze2 <= od_lo) { zem2 = 1. - ze2 + 0.5 * ze2 * ze2; zep2 = 1. / zem2; }
        else {
          double tblind = ze2 / (bpade + ze2);
          int itind = (int)(tblint * tblind + 0.5);
          zem2 = exp_tbl[itind];
          zep2 = 1. / zem2;
        }
        double zdenr = zr4 * zep1 + zr5 * zem1;
        double zdent = zt4 * zep1 + zt5 * zem1;
        if (zdenr >= -eps && zdenr <= eps) {
          pref(jk) = eps;
          ptra(jk) = zem2;
        } else {
          pref(jk) = zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) / zdenr;
          ptra(jk) = zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) / zdent;
        }
        double zemm = zem1 * zem1;
        double zdend = 1. / ((1. - zbeta * zemm) * zrkg);
        prefd(jk) = zgamma2 * (1. - zemm) * zdend;
        ptrad(jk) = zrk2 * zem1 * zdend;
      }
    }
  }
}

// vrtqdr_sw — rrtmg_sw_vrtqdr.f90:47-171
static void vrtqdr_sw(int klev, const A1& pref, const A1& prefd, const A1& ptra, const A1& ptrad, const A1& pdbt,
                      A1& prdnd, A1& prup, A1& prupd, const A1& ptdbt, A1& pfd, A1& pfu) {
  A1 ztdn(klev + 2);
  double zreflect = 1. / (1. - prefd(klev + 1) * prefd(klev));
  prup(klev) = pref(klev) + (ptrad(klev) * ((ptra(klev) - pdbt(klev)) * prefd(klev + 1) + pdbt(klev) * pref(klev + 1))) * zreflect;
  prupd(klev) = prefd(klev) + ptrad(klev) * ptrad(klev) * prefd(klev + 1) * zreflect;
  for (int jk = 1; jk <= klev - 1; ++jk) {
    int ikp = klev + 1 - jk, ikx = ikp - 1;
    zreflect = 1. / (1. - prupd(ikp) * prefd(ikx));
    prup(ikx) = pref(ikx) + (ptrad(ikx) * ((ptra(ikx) - pdbt(ikx)) * prupd(ikp) + pdbt(ikx) * prup(ikp))) * zreflect;
    prupd(ikx) = prefd(ikx) + ptrad(ikx) * ptrad(ikx) * prupd(ikp) * zreflect;
  }
  ztdn(1) = 1.;
  prdnd(1) = 0.;
  ztdn(2) = ptra(1);
  prdnd(2) = prefd(1);
  for (int jk = 2; jk <= klev; ++jk) {
    int ikp = jk + 1;
    zreflect = 1. / (1. - prefd(jk) * prdnd(jk));
    ztdn(ikp) = ptdbt(jk) * ptra(jk) + (ptrad(jk) * ((ztdn(jk) - ptdbt(jk)) + ptdbt(jk) * pref(jk) * prdnd(jk))) * zreflect;
    prdnd(ikp) = prefd(jk) + ptrad(jk) * ptrad(jk) * prdnd(jk) * zreflect;
  }
  for (int jk = 1; jk <= klev + 1; ++jk) {
    zreflect = 1. / (1. - prdnd(jk) * prupd(jk));
    pfu(jk) = (ptdbt(jk) * prup(jk) + (ztdn(jk) - ptdbt(jk)) * prupd(jk)) * zreflect;
    pfd(jk) = ptdbt(jk) + (ztdn(jk) - ptdbt(jk) + ptdbt(jk) * prup(jk) * prdnd(jk)) * zreflect;
  }
}

// spcvrt_sw — rrtmg_sw_spcvrt.f90:53-667 (idelm = 1, icpr = 1 as set by the driver)
struct SwFlux {
  A1 bbfd, bbfu, bbcd, bbcu;
  explicit SwFlux(int nlay) : bbfd(nlay + 2), bbfu(nlay + 2), bbcd(nlay + 2), bbcu(nlay + 2) {}
};
// With mc != nullptr this is spcvmc_sw (rrtmg_sw_spcvmc.f90:53-663): cloud fraction / optics per g-point
// (pcldfmc, ptaucmc, pasycmc, pomgcmc dimensioned (nlay, 112)) instead of per band.
struct McCloud {
  const A2 *cldfmc, *taucmc, *asycmc, *omgcmc;
};
static void spcvrt_sw(Col& c, int icpr, const double* palbd, const double* palbp, const A2& ptauc, const A2& pasyc,
                      const A2& pomgc, const A2& ptaua, const A2& pasya, const A2& pomga, double prmu0, int isolvar,
                      SwFlux& F, const McCloud* mc = nullptr) {
  const int klev = c.nlayers;
  const double repclc = 1.e-12, od_lo = 0.06, tblint = 10000.0, bpade = S.bpade;
  const std::vector<double>& exp_tbl = S.exp_tbl;
  for (int jk = 1; jk <= klev + 1; ++jk) { F.bbcd(jk) = 0.; F.bbcu(jk) = 0.; F.bbfd(jk) = 0.; F.bbfu(jk) = 0.; }
  Spectral sp(klev);
  taumol_sw(c, isolvar, sp);
  int n = klev + 2;
  A1 zdbt(n), zdbtc(n), zgcc(n), zgco(n), zomcc(n), zomco(n), zrdnd(n), zrdndc(n), zref(n), zrefc(n), zrefo(n), zrefd(n),
      zrefdc(n), zrefdo(n), zrup(n), zrupd(n), zrupc(n), zrupdc(n), ztauc(n), ztauo(n), ztdbt(n), ztdbtc(n), ztra(n),
      ztrac(n), ztrao(n), ztrad(n), ztradc(n), ztrado(n), zcd(n), zcu(n), zfd(n), zfu(n);
  std::vector<char> lrtchkclr(n, 0), lrtchkcld(n, 0);
  int iw = 0;
  for (int jb = jpb1; jb <= jpb2; ++jb) {
    int ibm = jb - 15;
    int igt = ngc_[ibm - 1];
    for (int jg = 1; jg <= igt; ++jg) {
      iw = iw + 1;
      double zincflx = 0.;
      if (isolvar < 0) zincflx = c.adjflux[jb] * sp.sfluxzen[iw] * prmu0;
      if (isolvar >= 0) zincflx = c.adjflux[jb] * sp.ssi[iw] * prmu0;
      ztdbtc(1) = 1.0;
      zdbtc(klev + 1) = 0.0; ztrac(klev + 1) = 0.0; ztradc(klev + 1) = 0.0;
      zrefc(klev + 1) = palbp[ibm]; zrefdc(klev + 1) = palbd[ibm]; zrupc(klev + 1) = palbp[ibm]; zrupdc(klev + 1) = palbd[ibm];
      ztrao(klev + 1) = 0.0; ztrado(klev + 1) = 0.0; zrefo(klev + 1) = palbp[ibm]; zrefdo(klev + 1) = palbd[ibm];
      ztdbt(1) = 1.0;
      zdbt(klev + 1) = 0.0; ztra(klev + 1) = 0.0; ztrad(klev + 1) = 0.0;
      zref(klev + 1) = palbp[ibm]; zrefd(klev + 1) = palbd[ibm]; zrup(klev + 1) = palbp[ibm]; zrupd(klev + 1) = palbd[ibm];
      for (int jk = 1; jk <= klev; ++jk) {
        int ikl = klev + 1 - jk;
        lrtchkclr[jk] = 1;
        const double clf = mc ? (*mc->cldfmc)(ikl, iw) : c.cldfrac(ikl);
        const double tcl = mc ? (*mc->taucmc)(ikl, iw) : ptauc(ikl, ibm);
        const double ocl = mc ? (*mc->omgcmc)(ikl, iw) : pomgc(ikl, ibm);
        const double gcl = mc ? (*mc->asycmc)(ikl, iw) : pasyc(ikl, ibm);
        lrtchkcld[jk] = (clf > repclc);
        ztauc(jk) = sp.taur(ikl, iw) + sp.taug(ikl, iw) + ptaua(ikl, ibm);
        zomcc(jk) = sp.taur(ikl, iw) * 1.0 + ptaua(ikl, ibm) * pomga(ikl, ibm);
        zgcc(jk) = pasya(ikl, ibm) * pomga(ikl, ibm) * ptaua(ikl, ibm) / zomcc(jk);
        zomcc(jk) = zomcc(jk) / ztauc(jk);
        double zf = zgcc(jk) * zgcc(jk);
        double zwf = zomcc(jk) * zf;
        ztauc(jk) = (1.0 - zwf) * ztauc(jk);
        zomcc(jk) = (zomcc(jk) - zwf) / (1.0 - zwf);
        zgcc(jk) = (zgcc(jk) - zf) / (1.0 - zf);
        if (icpr >= 1) {
          ztauo(jk) = ztauc(jk) + tcl;
          zomco(jk) = ztauc(jk) * zomcc(jk) + tcl * ocl;
          zgco(jk) = (tcl * ocl * gcl + ztauc(jk) * zomcc(jk) * zgcc(jk)) / zomco(jk);
          zomco(jk) = zomco(jk) / ztauo(jk);
        } else {
          ztauo(jk) = sp.taur(ikl, iw) + sp.taug(ikl, iw) + ptaua(ikl, ibm) + tcl;
          zomco(jk) = ptaua(ikl, ibm) * pomga(ikl, ibm) + tcl * ocl + sp.taur(ikl, iw) * 1.0;
          zgco(jk) = (tcl * ocl * gcl + ptaua(ikl, ibm) * pomga(ikl, ibm) * pasya(ikl, ibm)) / zomco(jk);
          zomco(jk) = zomco(jk) / ztauo(jk);
          zf = zgco(jk) * zgco(jk);
          zwf = zomco(jk) * zf;
          ztauo(jk) = (1. - zwf) * ztauo(jk);
          zomco(jk) = (zomco(jk) - zwf) / (1.0 - zwf);
          zgco(jk) = (zgco(jk) - zf) / (1.0 - zf);
        }
      }
      reftra_sw(klev, lrtchkclr, zgcc, prmu0, ztauc, zomcc, zrefc, zrefdc, ztrac, ztradc);
      reftra_sw(klev, lrtchkcld, zgco, prmu0, ztauo, zomco, zrefo, zrefdo, ztrao, ztrado);
      for (int jk = 1; jk <= klev; ++jk) {
        int ikl = klev + 1 - jk;
        const double clf2 = mc ? (*mc->cldfmc)(ikl, iw) : c.cldfrac(ikl);
        double zclear = 1.0 - clf2, zcloud = clf2;
        zref(jk) = zclear * zrefc(jk) + zcloud * zrefo(jk);
        zrefd(jk) = zclear * zrefdc(jk) + zcloud * zrefdo(jk);
        ztra(jk) = zclear * ztrac(jk) + zcloud * ztrao(jk);
        ztrad(jk) = zclear * ztradc(jk) + zcloud * ztrado(jk);
        double ze1 = ztauc(jk) / prmu0, zdbtmc, zdbtmo;
        if (ze1 <= od_lo) zdbtmc = 1. - ze1 + 0.5 * ze1 * ze1;
        else {
          double tblind = ze1 / (bpade + ze1);
          int itind = (int)(tblint * tblind + 0.5);
          zdbtmc = exp_tbl[itind];
        }
        zdbtc(jk) = zdbtmc;
        ztdbtc(jk + 1) = zdbtc(jk) * ztdbtc(jk);
        ze1 = ztauo(jk) / prmu0;
        if (ze1 <= od_lo) zdbtmo = 1. - ze1 + 0.5 * ze1 * ze1;
        else {
          double tblind = ze1 / (bpade + ze1);
          int itind = (int)(tblint * tblind + 0.5);
          zdbtmo = exp_tbl[itind];
        }
        zdbt(jk) = zclear * zdbtmc + zcloud * zdbtmo;
        ztdbt(jk + 1) = zdbt(jk) * ztdbt(jk);
      }
      vrtqdr_sw(klev, zrefc, zrefdc, ztrac, ztradc, zdbtc, zrdndc, zrupc, zrupdc, ztdbtc, zcd, zcu);
      vrtqdr_sw(klev, zref, zrefd, ztra, ztrad, zdbt, zrdnd, zrup, zrupd, ztdbt, zfd, zfu);
      for (int jk = 1; jk <= klev + 1; ++jk) {
        int ikl = klev + 2 - jk;
        F.bbfu(ikl) = F.bbfu(ikl) + zincflx * zfu(jk);
        F.bbfd(ikl) = F.bbfd(ikl) + zincflx * zfd(jk);
        F.bbcu(ikl) = F.bbcu(ikl) + zincflx * zcu(jk);
        F.bbcd(ikl) = F.bbcd(ikl) + zincflx * zcd(jk);
      }
    }
  }
}

}  // namespace orcsw

// =============================================================================================
using namespace orcsw;
static std::string g_err_sw;
extern "C" const char* orc_sw_last_error() { return g_err_sw.c_str(); }

// rrtmg_sw_set_constants (rrtmg_sw_c_binder.f90:19-46)
extern "C" void orc_sw_set_constants(double pi, double grav, double planck, double boltz, double clight, double avogad,
                                     double alosmt, double gascon, double sbcnst, double secdy) {
  (void)planck; (void)boltz; (void)clight; (void)alosmt; (void)gascon; (void)sbcnst;
  S.pi = pi; S.grav = grav; S.avogad = avogad; S.secdy = secdy;
}
extern "C" int orc_sw_ini(const char* raw_blob, double cpdair) {
  try {
    Blob b(raw_blob);
    sw_ini(b, cpdair);
  } catch (std::exception& e) {
    g_err_sw = e.what();
    return 1;
  }
  return 0;
}
extern "C" int orc_sw_get_reduced(int band /*1..14*/, const char* name, double* out, int64_t cap) {
  auto it = S.band[band].t.find(name);
  if (it == S.band[band].t.end()) return -1;
  int64_t n = (int64_t)it->second.d.size();
  if (out && cap >= n) std::memcpy(out, it->second.d.data(), sizeof(double) * (size_t)n);
  return (int)n;
}

// rrtmg_sw_nomcica_wrapper (rrtmg_sw_c_binder.f90:203-296) -> rrtmg_sw (rrtmg_sw_rad.nomcica.f90:97-816)
extern "C" int orc_sw_nomcica(int ncol, int nlay, int* icld, int* iaer, const double* play, const double* plev,
                              const double* tlay, const double* tlev, const double* tsfc, const double* h2ovmr,
                              const double* o3vmr, const double* co2vmr, const double* ch4vmr, const double* n2ovmr,
                              const double* o2vmr, const double* asdir, const double* asdif, const double* aldir,
                              const double* aldif, const double* coszen, double adjes, int dyofyr, double scon,
                              int isolvar, int inflgsw, int iceflgsw, int liqflgsw, const double* cldfr,
                              const double* taucld, const double* ssacld, const double* asmcld, const double* fsfcld,
                              const double* cicewp, const double* cliqwp, const double* reice, const double* reliq,
                              const double* tauaer, const double* ssaaer, const double* asmaer, const double* ecaer,
                              double* swuflx, double* swdflx, double* swhr, double* swuflxc, double* swdflxc,
                              double* swhrc, const double* bndsolvar, double* indsolvar, double solcycfrac) {
  if (!S.ready) { g_err_sw = "orc_sw_ini not called"; return 1; }
  const double zepsec = 1.e-06, zepzen = 1.e-10;
  S.oneminus = 1.0 - zepsec;
  S.pi = 2. * std::asin(1.);
  if (*icld < 0 || *icld > 3) *icld = 2;
  if (*iaer != 0 && *iaer != 6 && *iaer != 10) *iaer = 0;
  SwIn in{ncol, nlay, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, asdir, asdif, aldir, aldif,
          coszen, cldfr, taucld, ssacld, asmcld, fsfcld, cicewp, cliqwp, reice, reliq, tauaer, ssaaer, asmaer, ecaer};
  Col c(nlay);
  SwFlux F(nlay);
  A2 ztauc(nlay + 1, 14), zasyc(nlay + 1, 14), zomgc(nlay + 1, 14), ztaua(nlay + 1, 14), zasya(nlay + 1, 14), zomga(nlay + 1, 14);
  for (int iplon = 1; iplon <= ncol; ++iplon) {
    inatm_sw(in, iplon, *icld, *iaer, adjes, dyofyr, scon, isolvar, inflgsw, iceflgsw, liqflgsw, bndsolvar, indsolvar,
             solcycfrac, c);
    for (int i = 1; i <= nlay; ++i)
      if (c.cldfrac(i) > zepsec && c.cldfrac(i) < S.oneminus) { g_err_sw = "PARTIAL CLOUD NOT ALLOWED"; return 2; }
    if (cldprop_sw(c, g_err_sw)) return 2;
    setcoef_sw(c);
    double cossza = coszen[iplon - 1];
    if (cossza < zepzen) cossza = zepzen;
    double albdir[15], albdif[15];
    for (int ib = 1; ib <= 9; ++ib) { albdir[ib] = aldir[iplon - 1]; albdif[ib] = aldif[iplon - 1]; }
    albdir[nbndsw] = aldir[iplon - 1]; albdif[nbndsw] = aldif[iplon - 1];
    for (int ib = 10; ib <= 13; ++ib) { albdir[ib] = asdir[iplon - 1]; albdif[ib] = asdif[iplon - 1]; }
    if (*icld == 0) {
      std::fill(ztauc.d.begin(), ztauc.d.end(), 0.);
      std::fill(zasyc.d.begin(), zasyc.d.end(), 0.);
      std::fill(zomgc.d.begin(), zomgc.d.end(), 1.);
    } else {
      for (int i = 1; i <= nlay; ++i)
        for (int ib = 1; ib <= nbndsw; ++ib) {
          ztauc(i, ib) = c.taucloud(i, jpb1 - 1 + ib);
          zasyc(i, ib) = c.asmcloud(i, jpb1 - 1 + ib);
          zomgc(i, ib) = c.ssacloud(i, jpb1 - 1 + ib);
        }
    }
    if (*iaer == 0) {
      std::fill(ztaua.d.begin(), ztaua.d.end(), 0.);
      std::fill(zasya.d.begin(), zasya.d.end(), 0.);
      std::fill(zomga.d.begin(), zomga.d.end(), 1.);
    } else if (*iaer == 6) {
      for (int i = 1; i <= nlay; ++i)
        for (int ib = 1; ib <= nbndsw; ++ib) {
          ztaua(i, ib) = 0.; zasya(i, ib) = 0.; zomga(i, ib) = 0.;
          for (int ia = 1; ia <= naerec; ++ia) {
            double ec = ecaer[(size_t)(iplon - 1) + (size_t)ncol * ((i - 1) + (size_t)nlay * (ia - 1))];
            ztaua(i, ib) = ztaua(i, ib) + S.rsrtaua(ib, ia) * ec;
            zomga(i, ib) = zomga(i, ib) + S.rsrtaua(ib, ia) * ec * S.rsrpiza(ib, ia);
            zasya(i, ib) = zasya(i, ib) + S.rsrtaua(ib, ia) * ec * S.rsrpiza(ib, ia) * S.rsrasya(ib, ia);
          }
          if (ztaua(i, ib) == 0.) { ztaua(i, ib) = 0.; zasya(i, ib) = 0.; zomga(i, ib) = 1.; }
          else {
            if (zomga(i, ib) != 0.) zasya(i, ib) = zasya(i, ib) / zomga(i, ib);
            if (ztaua(i, ib) != 0.) zomga(i, ib) = zomga(i, ib) / ztaua(i, ib);
          }
        }
    } else {
      for (int i = 1; i <= nlay; ++i)
        for (int ib = 1; ib <= nbndsw; ++ib) { ztaua(i, ib) = c.taua(i, ib); zasya(i, ib) = c.asma(i, ib); zomga(i, ib) = c.ssaa(i, ib); }
    }
    spcvrt_sw(c, 1, albdif, albdir, ztauc, zasyc, zomgc, ztaua, zasya, zomga, cossza, isolvar, F);
    for (int i = 1; i <= nlay + 1; ++i) {
      size_t o = (size_t)(iplon - 1) + (size_t)ncol * (i - 1);
      swuflxc[o] = F.bbcu(i); swdflxc[o] = F.bbcd(i); swuflx[o] = F.bbfu(i); swdflx[o] = F.bbfd(i);
    }
    for (int i = 1; i <= nlay; ++i) {
      size_t o = (size_t)(iplon - 1) + (size_t)ncol * (i - 1);
      double zdpgcp = S.heatfac / c.pdp(i);
      double n1 = F.bbcd(i + 1) - F.bbcu(i + 1), n0 = F.bbcd(i) - F.bbcu(i);
      swhrc[o] = (n1 - n0) * zdpgcp;
      n1 = F.bbfd(i + 1) - F.bbfu(i + 1); n0 = F.bbfd(i) - F.bbfu(i);
      swhr[o] = (n1 - n0) * zdpgcp;
    }
  }
  return 0;
}


// =============================================================================================
// McICA shortwave: mcica_subcol_sw + cldprmc_sw + spcvmc_sw (rrtmg_sw_rad.f90:100-850)
namespace orcsw {
static int ngb_sw_[112];
static void init_ngb_sw() {
  int ib = 1;
  for (int ig = 1; ig <= 112; ++ig) {
    while (ig > ngs_[ib - 1]) ++ib;
    ngb_sw_[ig - 1] = ib + 15;
  }
}

// cldprmc_sw — rrtmg_sw_cldprmc.f90:53-349; arrays (112, nlay)
static int cldprmc_sw(int nlayers, int inflag, int iceflag, int liqflag, const A2& cldfmc, const A2& ciwpmc, const A2& clwpmc,
                      const A1& reicmc, const A1& relqmc, A2& taormc, A2& taucmc, A2& ssacmc, A2& asmcmc, A2& fsfcmc,
                      std::string& err) {
  const double eps = 1.e-06, cldmin = 1.e-20;
  double extcoice[113] = {0}, gice[113] = {0}, ssacoice[113] = {0}, forwice[113] = {0}, extcoliq[113] = {0},
         gliq[113] = {0}, ssacoliq[113] = {0}, forwliq[113] = {0}, fdelta[113] = {0};
  for (int lay = 1; lay <= nlayers; ++lay)
    for (int ig = 1; ig <= ngptsw; ++ig) taormc(ig, lay) = taucmc(ig, lay);
  for (int lay = 1; lay <= nlayers; ++lay)
    for (int ig = 1; ig <= ngptsw; ++ig) {
      double cwp = ciwpmc(ig, lay) + clwpmc(ig, lay);
      if (cldfmc(ig, lay) >= cldmin && (cwp >= cldmin || taucmc(ig, lay) >= cldmin)) {
        if (inflag == 0) {
          double taucldorig_a = taucmc(ig, lay);
          double ffp = fsfcmc(ig, lay), ffp1 = 1.0 - ffp, ffpssa = 1.0 - ffp * ssacmc(ig, lay);
          double ssacloud_a = ffp1 * ssacmc(ig, lay) / ffpssa;
          double taucloud_a = ffpssa * taucldorig_a;
          taormc(ig, lay) = taucldorig_a;
          ssacmc(ig, lay) = ssacloud_a;
          taucmc(ig, lay) = taucloud_a;
          asmcmc(ig, lay) = (asmcmc(ig, lay) - ffp) / (ffp1);
        } else if (inflag == 1) {
          err = "INFLAG = 1 OPTION NOT AVAILABLE WITH MCICA"; return 1;
        } else if (inflag == 2) {
          double radice = reicmc(lay);
          int ib = ngb_sw_[ig - 1];
          int k = ib - 15;
          if (ciwpmc(ig, lay) == 0.0) { extcoice[ig] = 0.; ssacoice[ig] = 0.; gice[ig] = 0.; forwice[ig] = 0.; }
          else if (iceflag == 1) {
            if (radice < 13.0 || radice > 130.) { err = "ICE RADIUS OUT OF BOUNDS"; return 1; }
            int icx = 5;
            double w2 = wavenum2_[ib - 16];
            if (w2 > 1.43e04) icx = 1; else if (w2 > 7.7e03) icx = 2; else if (w2 > 5.3e03) icx = 3; else if (w2 > 4.0e03) icx = 4;
            extcoice[ig] = (S.abari(icx) + S.bbari(icx) / radice);
            ssacoice[ig] = 1. - S.cbari(icx) - S.dbari(icx) * radice;
            gice[ig] = S.ebari(icx) + S.fbari(icx) * radice;
            if (gice[ig] >= 1.) gice[ig] = 1. - eps;
            forwice[ig] = gice[ig] * gice[ig];
          } else if (iceflag == 2) {
            if (radice < 5.0 || radice > 131.0) { err = "ICE RADIUS OUT OF BOUNDS"; return 1; }
            double factor = (radice - 2.) / 3.;
            int index = (int)factor;
            if (index == 43) index = 42;
            double fint = factor - (double)index;
            extcoice[ig] = S.extice2(index, k) + fint * (S.extice2(index + 1, k) - S.extice2(index, k));
            ssacoice[ig] = S.ssaice2(index, k) + fint * (S.ssaice2(index + 1, k) - S.ssaice2(index, k));
            gice[ig] = S.asyice2(index, k) + fint * (S.asyice2(index + 1, k) - S.asyice2(index, k));
            forwice[ig] = gice[ig] * gice[ig];
          } else if (iceflag == 3) {
            if (radice < 5.0 || radice > 140.0) { err = "ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS"; return 1; }
            double factor = (radice - 2.) / 3.;
            int index = (int)factor;
            if (index == 46) index = 45;
            double fint = factor - (double)index;
            extcoice[ig] = S.extice3(index, k) + fint * (S.extice3(index + 1, k) - S.extice3(index, k));
            ssacoice[ig] = S.ssaice3(index, k) + fint * (S.ssaice3(index + 1, k) - S.ssaice3(index, k));
            gice[ig] = S.asyice3(index, k) + fint * (S.asyice3(index + 1, k) - S.asyice3(index, k));
            fdelta[ig] = S.fdlice3(index, k) + fint * (S.fdlice3(index + 1, k) - S.fdlice3(index, k));
            forwice[ig] = fdelta[ig] + 0.5 / ssacoice[ig];
            if (forwice[ig] > gice[ig]) forwice[ig] = gice[ig];
          }
          if (clwpmc(ig, lay) == 0.0) { extcoliq[ig] = 0.; ssacoliq[ig] = 0.; gliq[ig] = 0.; forwliq[ig] = 0.; }
          else if (liqflag == 1) {
            double radliq = relqmc(lay);
            if (radliq < 2.5 || radliq > 60.) { err = "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS"; return 1; }
            int index = (int)(radliq - 1.5);
            if (index == 0) index = 1;
            if (index == 58) index = 57;
            double fint = radliq - 1.5 - (double)index;
            extcoliq[ig] = S.extliq1(index, k) + fint * (S.extliq1(index + 1, k) - S.extliq1(index, k));
            ssacoliq[ig] = S.ssaliq1(index, k) + fint * (S.ssaliq1(index + 1, k) - S.ssaliq1(index, k));
            if (fint < 0. && ssacoliq[ig] > 1.) ssacoliq[ig] = S.ssaliq1(index, k);
            gliq[ig] = S.asyliq1(index, k) + fint * (S.asyliq1(index + 1, k) - S.asyliq1(index, k));
            forwliq[ig] = gliq[ig] * gliq[ig];
          }
          double tauliqorig = clwpmc(ig, lay) * extcoliq[ig], tauiceorig = ciwpmc(ig, lay) * extcoice[ig];
          taormc(ig, lay) = tauliqorig + tauiceorig;
          double ssaliq = ssacoliq[ig] * (1. - forwliq[ig]) / (1. - forwliq[ig] * ssacoliq[ig]);
          double tauliq = (1. - forwliq[ig] * ssacoliq[ig]) * tauliqorig;
          double ssaice = ssacoice[ig] * (1. - forwice[ig]) / (1. - forwice[ig] * ssacoice[ig]);
          double tauice = (1. - forwice[ig] * ssacoice[ig]) * tauiceorig;
          double scatliq = ssaliq * tauliq, scatice = ssaice * tauice;
          taucmc(ig, lay) = tauliq + tauice;
          if (taucmc(ig, lay) == 0.) taucmc(ig, lay) = cldmin;
          if (scatice == 0.) scatice = cldmin;
          ssacmc(ig, lay) = (scatliq + scatice) / taucmc(ig, lay);
          if (iceflag == 3)
            asmcmc(ig, lay) = (1.0 / (scatliq + scatice)) * (scatliq * (gliq[ig] - forwliq[ig]) / (1.0 - forwliq[ig]) +
                                                             scatice * ((gice[ig] - forwice[ig]) / (1.0 - forwice[ig])));
          else
            asmcmc(ig, lay) = (scatliq * (gliq[ig] - forwliq[ig]) / (1.0 - forwliq[ig]) +
                               scatice * (gice[ig] - forwice[ig]) / (1.0 - forwice[ig])) / (scatliq + scatice);
        }
      }
    }
  return 0;
}
}  // namespace orcsw

// mcica_subcol_sw_wrapper + rrtmg_sw_mcica_wrapper (rrtmg_sw_c_binder.f90:59-201) as chained by _rrtmg_sw.pyx:284-417
extern "C" int orc_sw_mcica(int ncol, int nlay, int* icld, int* iaer, int permuteseed, int irng, const double* play,
                            const double* plev, const double* tlay, const double* tlev, const double* tsfc,
                            const double* h2ovmr, const double* o3vmr, const double* co2vmr, const double* ch4vmr,
                            const double* n2ovmr, const double* o2vmr, const double* asdir, const double* asdif,
                            const double* aldir, const double* aldif, const double* coszen, double adjes, int dyofyr,
                            double scon, int isolvar, int inflgsw, int iceflgsw, int liqflgsw, const double* cldfr,
                            const double* taucld, const double* ssacld, const double* asmcld, const double* fsfcld,
                            const double* cicewp, const double* cliqwp, const double* reice, const double* reliq,
                            const double* tauaer, const double* ssaaer, const double* asmaer, const double* ecaer,
                            double* swuflx, double* swdflx, double* swhr, double* swuflxc, double* swdflxc, double* swhrc,
                            const double* bndsolvar, double* indsolvar, double solcycfrac) {
  if (!S.ready) { g_err_sw = "orc_sw_ini not called"; return 1; }
  init_ngb_sw();
  const double zepzen = 1.e-10;
  S.oneminus = 1.0 - 1.e-06;
  S.pi = 2. * std::asin(1.);
  if (*icld < 0 || *icld > 3) { g_err_sw = "MCICA_SUBCOL: INVALID ICLD"; return 2; }
  if (*iaer != 0 && *iaer != 6 && *iaer != 10) *iaer = 0;
  const size_t nsub = 112, tot = nsub * (size_t)ncol * nlay;
  std::vector<double> cldfmcl(tot, 0.), ciwpmcl(tot, 0.), clwpmcl(tot, 0.), taucmcl(tot, 0.), ssacmcl(tot, 1.),
      asmcmcl(tot, 0.), fsfcmcl(tot, 0.);
  if (*icld != 0) {
    std::vector<double> pmid((size_t)ncol * nlay);
    for (size_t i = 0; i < pmid.size(); ++i) pmid[i] = play[i] * 1.e2;
    if (orc::generate_stochastic_clouds(ncol, nlay, 112, *icld, irng, pmid.data(), cldfr, cliqwp, cicewp, taucld, 14,
                                        ngb_sw_, 15, cldfmcl.data(), clwpmcl.data(), ciwpmcl.data(), taucmcl.data(),
                                        permuteseed, g_err_sw, ssacld, asmcld, fsfcld, ssacmcl.data(), asmcmcl.data(),
                                        fsfcmcl.data()))
      return 2;
  }
  SwIn in{ncol, nlay, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr, asdir, asdif, aldir, aldif,
          coszen, cldfr, taucld, ssacld, asmcld, fsfcld, cicewp, cliqwp, reice, reliq, tauaer, ssaaer, asmaer, ecaer};
  Col c(nlay);
  SwFlux F(nlay);
  A2 dummy(nlay + 1, 14), ztaua(nlay + 1, 14), zasya(nlay + 1, 14), zomga(nlay + 1, 14);
  A2 cldfmc(112, nlay + 1), taucmc(112, nlay + 1), ciwpmc(112, nlay + 1), clwpmc(112, nlay + 1), ssacmc(112, nlay + 1),
      asmcmc(112, nlay + 1), fsfcmc(112, nlay + 1), taormc(112, nlay + 1);
  A2 zcldfmc(nlay + 1, 112), ztaucmc(nlay + 1, 112), zasycmc(nlay + 1, 112), zomgcmc(nlay + 1, 112);
  A1 reicmc(nlay + 1), relqmc(nlay + 1);
  for (int iplon = 1; iplon <= ncol; ++iplon) {
    inatm_sw(in, iplon, 0, *iaer, adjes, dyofyr, scon, isolvar, inflgsw, iceflgsw, liqflgsw, bndsolvar, indsolvar,
             solcycfrac, c);
    std::fill(cldfmc.d.begin(), cldfmc.d.end(), 0.); std::fill(taucmc.d.begin(), taucmc.d.end(), 0.);
    std::fill(ssacmc.d.begin(), ssacmc.d.end(), 1.); std::fill(asmcmc.d.begin(), asmcmc.d.end(), 0.);
    std::fill(fsfcmc.d.begin(), fsfcmc.d.end(), 0.); std::fill(ciwpmc.d.begin(), ciwpmc.d.end(), 0.);
    std::fill(clwpmc.d.begin(), clwpmc.d.end(), 0.);
    if (*icld >= 1)
      for (int l = 1; l <= nlay; ++l) {
        for (int ig = 1; ig <= 112; ++ig) {
          size_t o = (size_t)(ig - 1) + nsub * ((size_t)(iplon - 1) + (size_t)ncol * (l - 1));
          cldfmc(ig, l) = cldfmcl[o]; taucmc(ig, l) = taucmcl[o]; ssacmc(ig, l) = ssacmcl[o]; asmcmc(ig, l) = asmcmcl[o];
          fsfcmc(ig, l) = fsfcmcl[o]; ciwpmc(ig, l) = ciwpmcl[o]; clwpmc(ig, l) = clwpmcl[o];
        }
        reicmc(l) = reice[(size_t)(iplon - 1) + (size_t)ncol * (l - 1)];
        relqmc(l) = reliq[(size_t)(iplon - 1) + (size_t)ncol * (l - 1)];
      }
    if (cldprmc_sw(nlay, inflgsw, iceflgsw, liqflgsw, cldfmc, ciwpmc, clwpmc, reicmc, relqmc, taormc, taucmc, ssacmc, asmcmc,
                   fsfcmc, g_err_sw))
      return 2;
    setcoef_sw(c);
    double cossza = coszen[iplon - 1];
    if (cossza < zepzen) cossza = zepzen;
    double albdir[15], albdif[15];
    for (int ib = 1; ib <= 9; ++ib) { albdir[ib] = aldir[iplon - 1]; albdif[ib] = aldif[iplon - 1]; }
    albdir[nbndsw] = aldir[iplon - 1]; albdif[nbndsw] = aldif[iplon - 1];
    for (int ib = 10; ib <= 13; ++ib) { albdir[ib] = asdir[iplon - 1]; albdif[ib] = asdif[iplon - 1]; }
    if (*icld == 0) {
      std::fill(zcldfmc.d.begin(), zcldfmc.d.end(), 0.); std::fill(ztaucmc.d.begin(), ztaucmc.d.end(), 0.);
      std::fill(zasycmc.d.begin(), zasycmc.d.end(), 0.); std::fill(zomgcmc.d.begin(), zomgcmc.d.end(), 1.);
    } else {
      for (int i = 1; i <= nlay; ++i)
        for (int ig = 1; ig <= 112; ++ig) {
          zcldfmc(i, ig) = cldfmc(ig, i); ztaucmc(i, ig) = taucmc(ig, i); zasycmc(i, ig) = asmcmc(ig, i); zomgcmc(i, ig) = ssacmc(ig, i);
        }
    }
    if (*iaer == 0) {
      std::fill(ztaua.d.begin(), ztaua.d.end(), 0.); std::fill(zasya.d.begin(), zasya.d.end(), 0.); std::fill(zomga.d.begin(), zomga.d.end(), 1.);
    } else if (*iaer == 6) {
      for (int i = 1; i <= nlay; ++i)
        for (int ib = 1; ib <= nbndsw; ++ib) {
          ztaua(i, ib) = 0.; zasya(i, ib) = 0.; zomga(i, ib) = 0.;
          for (int ia = 1; ia <= naerec; ++ia) {
            double ec = ecaer[(size_t)(iplon - 1) + (size_t)ncol * ((i - 1) + (size_t)nlay * (ia - 1))];
            ztaua(i, ib) = ztaua(i, ib) + S.rsrtaua(ib, ia) * ec;
            zomga(i, ib) = zomga(i, ib) + S.rsrtaua(ib, ia) * ec * S.rsrpiza(ib, ia);
            zasya(i, ib) = zasya(i, ib) + S.rsrtaua(ib, ia) * ec * S.rsrpiza(ib, ia) * S.rsrasya(ib, ia);
          }
          if (ztaua(i, ib) == 0.) { ztaua(i, ib) = 0.; zasya(i, ib) = 0.; zomga(i, ib) = 1.; }
          else {
            if (zomga(i, ib) != 0.) zasya(i, ib) = zasya(i, ib) / zomga(i, ib);
            if (ztaua(i, ib) != 0.) zomga(i, ib) = zomga(i, ib) / ztaua(i, ib);
          }
        }
    } else {
      for (int i = 1; i <= nlay; ++i)
        for (int ib = 1; ib <= nbndsw; ++ib) { ztaua(i, ib) = c.taua(i, ib); zasya(i, ib) = c.asma(i, ib); zomga(i, ib) = c.ssaa(i, ib); }
    }
    McCloud mc{&zcldfmc, &ztaucmc, &zasycmc, &zomgcmc};
    spcvrt_sw(c, 1, albdif, albdir, dummy, dummy, dummy, ztaua, zasya, zomga, cossza, isolvar, F, &mc);
    for (int i = 1; i <= nlay + 1; ++i) {
      size_t o = (size_t)(iplon - 1) + (size_t)ncol * (i - 1);
      swuflxc[o] = F.bbcu(i); swdflxc[o] = F.bbcd(i); swuflx[o] = F.bbfu(i); swdflx[o] = F.bbfd(i);
    }
    for (int i = 1; i <= nlay; ++i) {
      size_t o = (size_t)(iplon - 1) + (size_t)ncol * (i - 1);
      double zdpgcp = S.heatfac / c.pdp(i);
      double n1 = F.bbcd(i + 1) - F.bbcu(i + 1), n0 = F.bbcd(i) - F.bbcu(i);
      swhrc[o] = (n1 - n0) * zdpgcp;
      n1 = F.bbfd(i + 1) - F.bbfu(i + 1); n0 = F.bbfd(i) - F.bbfu(i);
      swhr[o] = (n1 - n0) * zdpgcp;
    }
  }
  return 0;
}
