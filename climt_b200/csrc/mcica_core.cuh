// climt_b200 -- McICA sub-column cloud mask generation (shared by the LW and SW engines).
// Reference: generate_stochastic_clouds, climt/_lib/rrtmg_lw/mcica_subcol_gen_lw.f90:156-522 (identical logic in
// rrtmg_sw/mcica_subcol_gen_sw.f90:172-555).  The only thing the radiation kernels need from it is, per (layer,
// column), WHICH sub-columns (= g-points) are cloudy: cloudy sub-columns carry the layer's own water paths / band
// optics, clear ones carry zeros.  So instead of the reference's 4-7 arrays of (ngpt, ncol, nlay) doubles we store
// one bit per (g-point, layer, column): mask[(layer * nwords + (g >> 5)) * ncc + column], bit (g & 31).
//
//  * kissvec generator (irng = 0): seeded per column from the bottom four layer pressures (:324-337) and consumed in
//    (sub-column, layer) order by that column only -> column-parallel: one thread per column (device).
//  * Mersenne twister (irng = 1, climt's default): ONE stream for the whole call consumed in (sub-column, column,
//    layer) order (:360-368) -> inherently serial; generated on the host for bit parity (mcica_host.h).
#pragma once
#include "cb_common.h"

namespace cb {
namespace mcica {

constexpr int kMaxLayers = 203;  // parrrtm.f90:31 / parrrsw.f90:27 (mxlay)

struct Kiss {  // kissvec, mcica_subcol_gen_lw.f90:530-562 (wrap-around 32-bit integer arithmetic)
  unsigned s1, s2, s3, s4;
  CB_HD double next() {
    s1 = 69069u * s1 + 1327217885u;
    s2 = s2 ^ (s2 << 13);
    s2 = s2 ^ (s2 >> 17);
    s2 = s2 ^ (s2 << 5);
    s3 = 18000u * (s3 & 65535u) + (s3 >> 16);
    s4 = 30903u * (s4 & 65535u) + (s4 >> 16);
    const int kiss = (int)(s1 + s2 + (s3 << 16) + s4);
    return kiss * 2.328306e-10 + 0.5;
  }
};

// threshold / overlap logic shared by both generators: feeds random numbers in (sub-column, layer) order for ONE
// column and sets mask bits.  icld: 1 random, 2 maximum-random, 3 maximum (:349-466, :470-472).
struct ColumnMasker {
  int icld;
  double prev;  // transformed CDF of the layer below (same sub-column)
  double rmax;  // icld = 3: the single number of this sub-column
  CB_HD void begin_subcolumn() { prev = 0.; rmax = 0.; }
  // r: next random number; l: 0-based layer (bottom first); cf_l / cf_below: cloud fractions (already zeroed below cldmin)
  CB_HD bool step(double r, int l, double cf_l, double cf_below) {
    double cdf = r;
    if (icld == 2 && l > 0) {
      if (prev > 1. - cf_below) cdf = prev;
      else cdf = r * (1. - cf_below);
    }
    prev = cdf;
    return cdf >= 1. - cf_l;
  }
};

// One column with the kissvec generator.  play [hPa] (nlay, ncol) column-fastest.  Returns 1 if the seed generator's
// precondition fails ('KISSVEC SEED GENERATOR REQUIRES PMID FROM BOTTOM FOUR LAYERS').
CB_HD int mask_column_kiss(const double* __restrict__ play, const double* __restrict__ cldfr, int ncol, int nlay,
                           int nsub, int nwords, int icld, int changeSeed, unsigned* __restrict__ mask, int ncc,
                           int c0, int c) {
  const size_t gc = (size_t)(c0 + c);
  if (icld == 0 || nlay < 4 || nlay > kMaxLayers) {
    for (int l = 0; l < nlay; ++l)
      for (int w = 0; w < nwords; ++w) mask[((size_t)l * nwords + w) * ncc + c] = 0u;
    return icld == 0 ? 0 : 1;
  }
  Kiss k;
  {
    const double p0 = play[gc] * 1.e2, p1 = play[(size_t)ncol + gc] * 1.e2, p2 = play[2 * (size_t)ncol + gc] * 1.e2,
                 p3 = play[3 * (size_t)ncol + gc] * 1.e2;
    if (p0 < p1) {
      for (int l = 0; l < nlay; ++l)
        for (int w = 0; w < nwords; ++w) mask[((size_t)l * nwords + w) * ncc + c] = 0u;
      return 1;
    }
    k.s1 = (unsigned)(int)((p0 - (double)(int)p0) * 1000000000.0);
    k.s2 = (unsigned)(int)((p1 - (double)(int)p1) * 1000000000.0);
    k.s3 = (unsigned)(int)((p2 - (double)(int)p2) * 1000000000.0);
    k.s4 = (unsigned)(int)((p3 - (double)(int)p3) * 1000000000.0);
  }
  for (int i = 0; i < changeSeed; ++i) k.next();
  ColumnMasker m;
  m.icld = icld;
  const double cldmin = 1.0e-20;
  // The generator is consumed in (sub-column, layer) order, the mask is stored (layer, word): the column's cloud fractions
  // and the 32 sub-columns of the current word live in thread-local arrays (L1), each mask word is written once.
  unsigned wd[kMaxLayers];
  for (int w = 0; w < nwords; ++w) {
    for (int l = 0; l < nlay; ++l) wd[l] = 0u;
    const int s1 = (w + 1) * 32 < nsub ? (w + 1) * 32 : nsub;
    for (int s = w * 32; s < s1; ++s) {
      m.begin_subcolumn();
      double r3 = 0.;
      if (icld == 3) r3 = k.next();
      double cf_below = 0.;
      for (int l = 0; l < nlay; ++l) {
        double cf = cldfr[(size_t)l * ncol + gc];
        if (cf < cldmin) cf = 0.;
        const double r = icld == 3 ? r3 : k.next();
        if (m.step(r, l, cf, cf_below)) wd[l] |= 1u << (s & 31);
        cf_below = cf;
      }
    }
    for (int l = 0; l < nlay; ++l) mask[((size_t)l * nwords + w) * ncc + c] = wd[l];
  }
  return 0;
}

}  // namespace mcica
}  // namespace cb
