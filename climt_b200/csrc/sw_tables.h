// climt_b200 -- host-side table loader + solar set-up for the SW engine (host code only).
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "lw_tables.h"  // BlobView, lw::Constants
#include "sw_core.cuh"

namespace cb {
namespace sw {

using Constants = lw::Constants;

inline void build_tables(const std::string& blob_path, const Constants& k, std::vector<double>& img, Tables& T) {
  BlobView b;
  b.load(blob_path);
  img.clear();
  auto put = [&](const double* p, size_t n) {
    if (img.size() & 1) img.push_back(0.0);
    int off = (int)img.size();
    img.insert(img.end(), p, p + n);
    return off;
  };
  auto putk = [&](const std::string& key) {
    const BlobView::Ent& e = b.get(key);
    return put(e.p, (size_t)e.count);
  };
  std::memset(&T, 0, sizeof(T));
  for (int ib = 0; ib < 14; ++ib) {
    char pre[8];
    std::snprintf(pre, sizeof pre, "b%02d.", ib + 16);
    BandOff& O = T.b[ib];
    auto opt = [&](const char* nm) { std::string key = std::string(pre) + nm; return b.has(key) ? putk(key) : -1; };
    // (rows, ng) table -> one (rows, 4) block per group of 4 g-points (layout: BandOff in sw_core.cuh); returns offset, sets stride
    const int ng = kNG[ib];
    auto grouped = [&](const char* nm, int& gs) {
      const std::string key = std::string(pre) + nm;
      gs = 0;
      if (!b.has(key)) return -1;
      const BlobView::Ent& e = b.get(key);
      const int rows = (int)(e.count / ng);
      while (img.size() & 15) img.push_back(0.0);
      const int off = (int)img.size();
      gs = rows * 4;
      for (int q = 0; q < (ng + 3) / 4; ++q)
        for (int r = 0; r < rows; ++r)
          for (int j = 0; j < 4; ++j) img.push_back(q * 4 + j < ng ? e.p[(size_t)r * ng + q * 4 + j] : 0.0);
      return off;
    };
    O.absa = grouped("absa", O.gs_absa); O.absb = grouped("absb", O.gs_absb);
    O.selfref = grouped("selfref", O.gs_selfref); O.forref = grouped("forref", O.gs_forref);
    O.sfluxref = opt("sfluxref"); O.irradnce = opt("irradnce"); O.facbrght = opt("facbrght"); O.snsptdrk = opt("snsptdrk");
    O.raylv = -1; O.raylb = opt("raylb"); O.rayl = 0.;
    std::string rk = std::string(pre) + "rayl";
    if (b.has(rk)) {
      const BlobView::Ent& e = b.get(rk);
      if (e.count == 1) O.rayl = e.p[0];
      else O.raylv = put(e.p, (size_t)e.count);
    }
    if (b.has(std::string(pre) + "rayla")) O.raylv = putk(std::string(pre) + "rayla");
    O.x0 = O.x1 = -1;
    const int band = ib + 16;
    if (band == 20) O.x0 = opt("absch4");
    if (band == 24 || band == 25) { O.x0 = opt("abso3a"); O.x1 = opt("abso3b"); }
    if (band == 29) { O.x0 = opt("absco2"); O.x1 = opt("absh2o"); }
  }
  T.preflog = putk("preflog");
  T.tref = putk("tref");
  T.extliq1 = putk("cld.extliq1"); T.ssaliq1 = putk("cld.ssaliq1"); T.asyliq1 = putk("cld.asyliq1");
  T.extice2 = putk("cld.extice2"); T.ssaice2 = putk("cld.ssaice2"); T.asyice2 = putk("cld.asyice2");
  T.extice3 = putk("cld.extice3"); T.ssaice3 = putk("cld.ssaice3"); T.asyice3 = putk("cld.asyice3");
  T.fdlice3 = putk("cld.fdlice3");
  T.abari = putk("cld.abari"); T.bbari = putk("cld.bbari"); T.cbari = putk("cld.cbari");
  T.dbari = putk("cld.dbari"); T.ebari = putk("cld.ebari"); T.fbari = putk("cld.fbari");
  T.rsrtaua = putk("aer.rsrtaua"); T.rsrpiza = putk("aer.rsrpiza"); T.rsrasya = putk("aer.rsrasya");
  {  // exp table, rrtmg_sw_init.f90:113-123 (kind=rb abscissa here, unlike the LW table)
    const double pade = 0.278, expeps = 1.e-20;
    T.bpade = 1.0 / pade;
    std::vector<double> ex(NTBL + 1);
    ex[0] = 1.0;
    ex[NTBL] = expeps;
    for (int itr = 1; itr <= NTBL - 1; ++itr) {
      const double tfn = (double)itr / (double)NTBL;
      const double tau = T.bpade * tfn / (1. - tfn);
      ex[itr] = std::exp(-tau);
      if (ex[itr] <= expeps) ex[itr] = expeps;
    }
    T.exp_tbl = put(ex.data(), ex.size());
  }
  T.oneminus = 1.0 - 1.e-06;
  T.heatfac = k.grav * k.secdy / (k.cpdair * 1.e2);
  T.avogad = k.avogad;
  T.grav = k.grav;
}

// Solar constant / variability / earth-sun distance set-up of inatm_sw (rrtmg_sw_rad.nomcica.f90:1195-1411);
// identical for every column, so it runs once per call on the host.  NRLSSI2 cycle tables :1122-1167.
struct SolarOptions {
  int isolvar = 0;
  double scon = 1367.0;
  double indsolvar[2] = {1.0, 1.0};
  double bndsolvar[14] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
};
namespace detail {
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
}  // namespace detail

inline Solar compute_solar(const SolarOptions& opt_in, double adjes, int dyofyr, double solcycfrac) {
  using namespace detail;
  SolarOptions opt = opt_in;  // indsolvar is intent(inout) in the reference; we work on a copy
  const double rrsw_scon = (double)1.36822e+03f;  // default-real parameter, parrrsw.f90:115
  const double Iint = 1360.37, Fint = 0.996047, Sint = -0.511590, Foffset = 0.14959542, Soffset = 0.00066696,
               svar_f_avg = 0.1568113, svar_s_avg = 909.21910;
  const int nsolfrac = 132;
  Solar s;
  s.isolvar = opt.isolvar;
  double solvar[14];
  for (int i = 0; i < 14; ++i) { solvar[i] = 1.0; s.adjflux[i] = 1.0; s.svar_f_bnd[i] = 1.0; s.svar_s_bnd[i] = 1.0; s.svar_i_bnd[i] = 1.0; }
  s.svar_f = 1.0; s.svar_s = 1.0; s.svar_i = 1.0;
  const double scon = opt.scon;
  double* ind = opt.indsolvar;
  if (opt.isolvar == 1 && (ind[0] != 1.0 || ind[1] != 1.0)) {
    double wgt;
    if (solcycfrac >= 0.0 && solcycfrac < 0.0229) {
      wgt = (solcycfrac + 1.0 - 0.3817) / (1.0229 - 0.3817);
      ind[0] = ind[0] + wgt * (1.0 - ind[0]);
      ind[1] = ind[1] + wgt * (1.0 - ind[1]);
    }
    if (solcycfrac >= 0.0229 && solcycfrac <= 0.3817) {
      wgt = (solcycfrac - 0.0229) / (0.3817 - 0.0229);
      ind[0] = 1.0 + wgt * (ind[0] - 1.0);
      ind[1] = 1.0 + wgt * (ind[1] - 1.0);
    }
    if (solcycfrac > 0.3817 && solcycfrac <= 1.0) {
      wgt = (solcycfrac - 0.3817) / (1.0229 - 0.3817);
      ind[0] = ind[0] + wgt * (1.0 - ind[0]);
      ind[1] = ind[1] + wgt * (1.0 - ind[1]);
    }
  }
  double adjflx = adjes;
  if (dyofyr > 0) {  // earth_sun, :818-843
    const double pi = 2. * std::asin(1.);
    const double gamma = 2. * pi * (dyofyr - 1) / 365.;
    adjflx = 1.000110 + .034221 * std::cos(gamma) + .001289 * std::sin(gamma) + .000719 * std::cos(2. * gamma) +
             .000077 * std::sin(2. * gamma);
  }
  auto cycle = [&](double& a0, double& b0) {
    if (solcycfrac <= 0.0) { a0 = mgavgcyc[0]; b0 = sbavgcyc[0]; }
    else if (solcycfrac >= 1.0) { a0 = mgavgcyc[nsolfrac - 1]; b0 = sbavgcyc[nsolfrac - 1]; }
    else {
      const int sfid = (int)std::floor(solcycfrac * (nsolfrac - 1)) + 1;
      const double inv = 1.0 / (nsolfrac - 1);
      const double fraclo = (sfid - 1) * inv, frachi = sfid * inv;
      const double f = (solcycfrac - fraclo) / (frachi - fraclo);
      a0 = mgavgcyc[sfid - 1] + f * (mgavgcyc[sfid] - mgavgcyc[sfid - 1]);
      b0 = sbavgcyc[sfid - 1] + f * (sbavgcyc[sfid] - sbavgcyc[sfid - 1]);
    }
  };
  if (scon == 0.0) {
    if (opt.isolvar == -1) for (int i = 0; i < 14; ++i) solvar[i] = opt.bndsolvar[i];
    if (opt.isolvar == 1) {
      double a0, b0;
      cycle(a0, b0);
      s.svar_f = ind[0] * (a0 - Foffset) / (svar_f_avg - Foffset);
      s.svar_s = ind[1] * (b0 - Soffset) / (svar_s_avg - Soffset);
    }
    if (opt.isolvar == 2) {
      s.svar_f = (ind[0] - Foffset) / (svar_f_avg - Foffset);
      s.svar_s = (ind[1] - Soffset) / (svar_s_avg - Soffset);
    }
    if (opt.isolvar == 3)
      for (int i = 0; i < 14; ++i) { solvar[i] = opt.bndsolvar[i]; s.svar_f_bnd[i] = s.svar_s_bnd[i] = s.svar_i_bnd[i] = solvar[i]; }
  } else if (scon > 0.0) {
    if (opt.isolvar == -1) for (int i = 0; i < 14; ++i) solvar[i] = opt.bndsolvar[i] * scon / rrsw_scon;
    if (opt.isolvar == 0) {
      const double r = scon / (Fint + Sint + Iint);
      s.svar_f = r; s.svar_s = r; s.svar_i = r;
    }
    if (opt.isolvar == 1) {
      double a0, b0;
      cycle(a0, b0);
      s.svar_i = (scon - (ind[0] * Fint + ind[1] * Sint)) / Iint;
      s.svar_f = ind[0] * (a0 - Foffset) / (svar_f_avg - Foffset);
      s.svar_s = ind[1] * (b0 - Soffset) / (svar_s_avg - Soffset);
    }
    if (opt.isolvar == 3) {
      const double cp = Fint + Sint + Iint;
      for (int i = 0; i < 14; ++i) { solvar[i] = opt.bndsolvar[i] * scon / cp; s.svar_f_bnd[i] = s.svar_s_bnd[i] = s.svar_i_bnd[i] = solvar[i]; }
    }
  }
  for (int i = 0; i < 14; ++i) s.adjflux[i] = opt.isolvar < 0 ? adjflx * solvar[i] : adjflx;
  return s;
}

}  // namespace sw
}  // namespace cb
