// climt_b200 -- host-side table loader for the LW engine: reads the reduced-table blob written by
// climt_b200/rrtmg_tables.py, appends the exp/tau/tfn lookup tables, and fills cb::lw::Tables offsets.
// Host code only (used by lw_engine.cu and by the test-only host emulation).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "lw_core.cuh"

namespace cb {

struct BlobView {
  std::vector<char> raw;
  struct Ent { std::vector<int64_t> shape; const double* p; int64_t count; };
  std::map<std::string, Ent> ent;
  void load(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open table blob: " + path);
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    raw.resize((size_t)sz);
    size_t got = std::fread(raw.data(), 1, (size_t)sz, f);
    std::fclose(f);
    if (got != (size_t)sz || sz < 16 || std::memcmp(raw.data(), "CB2TBL01", 8) != 0)
      throw std::runtime_error("bad table blob: " + path);
    int64_t n;
    std::memcpy(&n, raw.data() + 8, 8);
    const size_t esz = 56 + 8 + 48 + 8 + 8;
    const char* data0 = raw.data() + 16 + (size_t)n * esz;
    for (int64_t i = 0; i < n; ++i) {
      const char* e = raw.data() + 16 + (size_t)i * esz;
      char name[57];
      std::memcpy(name, e, 56);
      name[56] = 0;
      int64_t ndim, shp[6], off, cnt;
      std::memcpy(&ndim, e + 56, 8);
      std::memcpy(shp, e + 64, 48);
      std::memcpy(&off, e + 112, 8);
      std::memcpy(&cnt, e + 120, 8);
      Ent en;
      en.shape.assign(shp, shp + ndim);
      en.p = reinterpret_cast<const double*>(data0) + off;
      en.count = cnt;
      ent[name] = en;
    }
  }
  const Ent& get(const std::string& k) const {
    auto it = ent.find(k);
    if (it == ent.end()) throw std::runtime_error("table blob has no entry '" + k + "'");
    return it->second;
  }
  bool has(const std::string& k) const { return ent.count(k) != 0; }
};

namespace lw {

// The 10 constants of rrtmg_set_constants (rrlw_con.f90:46-71) + cp of dry air (rrtmg_lw_ini_wrapper).
struct Constants {
  double pi, grav, planck, boltz, clight, avogad, alosmt, gascon, sbcnst, secdy, cpdair;
};

// minor-gas / cross-section slot names per band (slot order is what region<B,LOWER>() refers to)
static const char* const kMinorNames[16][5] = {
    {"ka_mn2", "kb_mn2", 0, 0, 0},      {0, 0, 0, 0, 0},
    {"ka_mn2o", "kb_mn2o", 0, 0, 0},    {0, 0, 0, 0, 0},
    {"ka_mo3", 0, 0, 0, 0},             {"ka_mco2", 0, 0, 0, 0},
    {"ka_mco2", "kb_mco2", 0, 0, 0},    {"ka_mco2", "ka_mo3", "ka_mn2o", "kb_mco2", "kb_mn2o"},
    {"ka_mn2o", "kb_mn2o", 0, 0, 0},    {0, 0, 0, 0, 0},
    {"ka_mo2", "kb_mo2", 0, 0, 0},      {0, 0, 0, 0, 0},
    {"ka_mco2", "ka_mco", "kb_mo3", 0, 0}, {0, 0, 0, 0, 0},
    {"ka_mn2", 0, 0, 0, 0},             {0, 0, 0, 0, 0}};
static const char* const kXsecNames[16][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {"ccl4", 0}, {"cfc11adj", "cfc12"},
                                              {0, 0}, {"cfc12", "cfc22adj"}, {0, 0}, {0, 0}, {0, 0}, {0, 0},
                                              {0, 0}, {0, 0}, {0, 0}, {0, 0}};

// Builds the flat table image (host) and the offset struct.  `T.base` is left null; the caller points it
// at the host image (emulation) or at its HBM copy (engine).
inline void build_tables(const std::string& blob_path, const Constants& k, std::vector<double>& img, Tables& T) {
  BlobView b;
  b.load(blob_path);
  img.clear();
  auto put = [&](const double* p, size_t n) {
    // keep every table 16-byte aligned so paired g-point loads can be vectorised
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
  const BlobView::Ent& chi = b.get("chi_mls");
  auto CHI = [&](int imol, int jp) { return chi.p[(imol - 1) * 59 + (jp - 1)]; };
  for (int ib = 1; ib <= 16; ++ib) {
    char pre[8];
    std::snprintf(pre, sizeof pre, "b%02d.", ib);
    BandOff& O = T.b[ib - 1];
    const int ng = kNG[ib - 1], ngrp = (ng + TGW - 1) / TGW;
    // Every table of the band is (rows, ng) in the blob, g-point fastest.  Collect them, then write one block per group of TGW
    // g-points holding all of them row after row (layout: see BandOff in lw_core.cuh).
    std::vector<const BlobView::Ent*> ents;
    int nrows = 0;
    auto opt = [&](const char* nm) {
      const std::string key = std::string(pre) + nm;
      if (!b.has(key)) return -1;
      const BlobView::Ent& e = b.get(key);
      if (e.count % ng) throw std::runtime_error("table " + key + ": size is not a multiple of the band's g-point count");
      ents.push_back(&e);
      const int r0 = nrows;
      nrows += (int)(e.count / ng);
      return r0;
    };
    O.absa = opt("absa"); O.absb = opt("absb"); O.selfref = opt("selfref"); O.forref = opt("forref");
    O.fracrefa = opt("fracrefa"); O.fracrefb = opt("fracrefb");
    for (int s = 0; s < 5; ++s) O.m[s] = kMinorNames[ib - 1][s] ? opt(kMinorNames[ib - 1][s]) : -1;
    for (int s = 0; s < 2; ++s) O.x[s] = kXsecNames[ib - 1][s] ? opt(kXsecNames[ib - 1][s]) : -1;
    while (img.size() & 15) img.push_back(0.0);  // 128-byte aligned blocks (cp.async.bulk needs 16)
    O.base = (int)img.size();
    O.rows = nrows;
    for (int q = 0; q < ngrp; ++q)
      for (const BlobView::Ent* e : ents) {
        const int rows = (int)(e->count / ng);
        for (int r = 0; r < rows; ++r)
          for (int j = 0; j < TGW; ++j) {
            const int g = q * TGW + j;
            img.push_back(g < ng ? e->p[(size_t)r * ng + g] : 0.0);
          }
      }
  }
  // reference ratios (rrtmg_lw_taumol.f90:506-509, 777-778, 1040-1042, 1410-1411, 1806-1807, 2211, 2420-2422,
  // 2737-2738, 2951)
  auto& B = T.b;
  B[2].refrat_planck_a = CHI(1, 9) / CHI(2, 9);   B[2].refrat_planck_b = CHI(1, 13) / CHI(2, 13);
  B[2].refrat_m_a = CHI(1, 3) / CHI(2, 3);        B[2].refrat_m_b = CHI(1, 13) / CHI(2, 13);
  B[3].refrat_planck_a = CHI(1, 11) / CHI(2, 11); B[3].refrat_planck_b = CHI(3, 13) / CHI(2, 13);
  B[4].refrat_planck_a = CHI(1, 5) / CHI(2, 5);   B[4].refrat_planck_b = CHI(3, 43) / CHI(2, 43);
  B[4].refrat_m_a = CHI(1, 7) / CHI(2, 7);
  B[6].refrat_planck_a = CHI(1, 3) / CHI(3, 3);   B[6].refrat_m_a = CHI(1, 3) / CHI(3, 3);
  B[8].refrat_planck_a = CHI(1, 9) / CHI(6, 9);   B[8].refrat_m_a = CHI(1, 3) / CHI(6, 3);
  B[11].refrat_planck_a = CHI(1, 10) / CHI(2, 10);
  B[12].refrat_planck_a = CHI(1, 5) / CHI(4, 5);  B[12].refrat_m_a = CHI(1, 1) / CHI(4, 1);
  B[12].refrat_m_a3 = CHI(1, 3) / CHI(4, 3);
  B[14].refrat_planck_a = CHI(4, 1) / CHI(2, 1);  B[14].refrat_m_a = CHI(4, 1) / CHI(2, 1);
  B[15].refrat_planck_a = CHI(1, 6) / CHI(6, 6);
  T.chi_mls = putk("chi_mls");
  T.preflog = putk("preflog");
  T.tref = putk("tref");
  T.rat = putk("rat");
  T.totplnk = putk("totplnk");
  T.totplnkderiv = putk("totplnkderiv");
  T.delwave = putk("delwave");
  T.absice0 = putk("cld.absice0");
  T.absice1 = putk("cld.absice1");
  T.absice2 = putk("cld.absice2");
  T.absice3 = putk("cld.absice3");
  T.absliq1 = putk("cld.absliq1");
  T.abscld1 = b.get("cld.abscld1").p[0];
  T.absliq0 = b.get("cld.absliq0").p[0];
  // exp / tau / transmittance-function lookup tables (rrtmg_lw_init.f90:97-123); the abscissa is a
  // default-real (float32) quotient in the reference and is kept so.
  {
    const double pade = 0.278, expeps = 1.e-20;
    T.bpade = 1.0 / pade;
    std::vector<double> tau(NTBL + 1), ex(NTBL + 1), tfn(NTBL + 1);
    tau[0] = 0.0; tau[NTBL] = 1.e10; ex[0] = 1.0; ex[NTBL] = expeps; tfn[0] = 0.0; tfn[NTBL] = 1.0;
    for (int itr = 1; itr <= NTBL - 1; ++itr) {
      const double x = (double)((float)itr / (float)NTBL);
      tau[itr] = T.bpade * x / (1. - x);
      ex[itr] = std::exp(-tau[itr]);
      if (ex[itr] <= expeps) ex[itr] = expeps;
      if (tau[itr] < 0.06) tfn[itr] = tau[itr] / 6.;
      else tfn[itr] = 1. - 2. * ((1. / tau[itr]) - (ex[itr] / (1. - ex[itr])));
    }
    T.tau_tbl = put(tau.data(), tau.size());
    T.exp_tbl = put(ex.data(), ex.size());
    T.tfn_tbl = put(tfn.data(), tfn.size());
    std::vector<double> et(2 * (NTBL + 1));
    for (int i = 0; i <= NTBL; ++i) { et[2 * i] = ex[i]; et[2 * i + 1] = tfn[i]; }
    T.et_tbl = put(et.data(), et.size());
  }
  // rrtmg_lw_rad.nomcica.f90:419-421, rrtmg_lw_init.f90:279
  T.oneminus = 1. - 1.e-6;
  const double pi = 2. * std::asin(1.);
  T.fluxfac = pi * 2.e4;
  T.heatfac = k.grav * k.secdy / (k.cpdair * 1.e2);
  T.avogad = k.avogad;
  T.grav = k.grav;
}

}  // namespace lw
}  // namespace cb
