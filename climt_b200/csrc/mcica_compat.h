// climt_b200 -- host-side glue for the reference's two-step McICA C ABI (host only).
//
// The reference splits a McICA call in two bind(c) symbols that talk through (ngpt, ncol, nlay) arrays of doubles
// (climt/_lib/rrtmg_lw/rrtmg_lw_c_binder.f90:50-92 + :94-174; rrtmg_sw/rrtmg_sw_c_binder.f90:59-107 + :109-201;
// called back to back by climt/_components/rrtmg/lw/_rrtmg_lw.pyx:261-320 and sw/_rrtmg_sw.pyx:341-417):
//   mcica_subcol_*_wrapper   sub-column generator: cloud fraction (0/1), water paths and band optics fanned out per g-point
//   rrtmg_*_mcica_wrapper    radiative transfer on those arrays
// The engines keep the same information as one bit per (g-point, layer, column) plus the layer's own water paths / band
// optics (mcica_core.cuh), because a cloudy sub-column always carries its layer's values and a clear one zeros
// (mcica_subcol_gen_lw.f90:470-497).  These helpers convert between the two forms:
//   expand()    mask + layer values -> the reference's per-g-point arrays (what mcica_subcol_*_wrapper must fill)
//   collapse()  per-g-point arrays  -> mask + layer values (what rrtmg_*_mcica_wrapper feeds the engine), refusing input the
//               bit form cannot hold: a sub-column cloud fraction that is not 0 or 1, or cloudy sub-columns of one layer
//               that carry different water paths / optics.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "engine_common.h"

namespace cb {
namespace mcica {

// index of element (g, i, l) of a Fortran (ngpt, ncol, nlay) array
inline size_t gidx(int ngpt, int ncol, int g, int i, int l) { return ((size_t)l * ncol + i) * ngpt + g; }

// mask[(l * nwords + (g >> 5)) * ncol + i], bit (g & 31): sub-column g of (layer l, column i) is cloudy.
// `layer` arrays are (nlay, ncol) column-fastest; `band` arrays Fortran (nbnd, ncol, nlay); ngb[g] = 0-based band of g-point g.
struct SubcolOut {      // what mcica_subcol_*_wrapper fills; band-resolved outputs beyond LW's taucmcl are optional
  double *cldfmcl, *ciwpmcl, *clwpmcl, *reicmcl, *relqmcl;
  double* bandmcl[4];   // taucmcl [, ssacmcl, asmcmcl, fsfcmcl]
};
struct SubcolIn {
  const double *ciwp, *clwp, *rei, *rel;
  const double* band[4];  // tauc [, ssac, asmc, fsfc] (nbnd, ncol, nlay)
  double clear_value[4];  // value of a clear sub-column: 0 for tau; SW: ssa 1, asm 0, fsf 0 (mcica_subcol_gen_sw.f90:523-548)
};

inline void expand(const unsigned* mask, int ncol, int nlay, int ngpt, int nwords, int nbnd, const int* ngb, int nband_arrays,
                   const SubcolIn& in, const SubcolOut& out) {
  WorkerPool::get().parallel_for(nlay, [&](int l) {
    for (int i = 0; i < ncol; ++i) {
      const size_t o = (size_t)l * ncol + i;
      out.reicmcl[o] = in.rei[o];
      out.relqmcl[o] = in.rel[o];
      const double ci = in.ciwp[o], cl = in.clwp[o];
      for (int g = 0; g < ngpt; ++g) {
        const bool on = (mask[((size_t)l * nwords + (g >> 5)) * ncol + i] >> (g & 31)) & 1u;
        const size_t k = o * ngpt + g;
        out.cldfmcl[k] = on ? 1.0 : 0.0;
        out.ciwpmcl[k] = on ? ci : 0.0;
        out.clwpmcl[k] = on ? cl : 0.0;
        for (int a = 0; a < nband_arrays; ++a)
          out.bandmcl[a][k] = on ? in.band[a][o * nbnd + ngb[g]] : in.clear_value[a];
      }
    }
  });
}

struct CollapseOut {
  std::vector<unsigned> mask;                // [nlay][nwords][ncol]
  std::vector<double> cldfr, ciwp, clwp;     // (nlay, ncol)
  std::vector<double> band[4];               // (nbnd, ncol, nlay) Fortran order
};

// Returns "" on success, else a message.
inline std::string collapse(int ncol, int nlay, int ngpt, int nwords, int nbnd, const int* ngb, int nband_arrays,
                            const double* cldfmcl, const double* ciwpmcl, const double* clwpmcl,
                            const double* const bandmcl[4], CollapseOut& out) {
  const size_t n2 = (size_t)nlay * ncol;
  out.mask.assign((size_t)nlay * nwords * ncol, 0u);
  out.cldfr.assign(n2, 0.0);
  out.ciwp.assign(n2, 0.0);
  out.clwp.assign(n2, 0.0);
  for (int a = 0; a < nband_arrays; ++a) out.band[a].assign(n2 * nbnd, 0.0);
  std::atomic<int> bad{0};
  WorkerPool::get().parallel_for(nlay, [&](int l) {
    std::vector<char> seen(nbnd);
    for (int i = 0; i < ncol; ++i) {
      const size_t o = (size_t)l * ncol + i;
      int ncloudy = 0;
      double ci = 0., cl = 0.;
      std::fill(seen.begin(), seen.end(), 0);
      for (int g = 0; g < ngpt; ++g) {
        const size_t k = o * ngpt + g;
        const double f = cldfmcl[k];
        if (f == 0.0) continue;
        if (f != 1.0) { bad.store(1); continue; }
        out.mask[((size_t)l * nwords + (g >> 5)) * ncol + i] |= 1u << (g & 31);
        if (ncloudy == 0) { ci = ciwpmcl[k]; cl = clwpmcl[k]; }
        else if (ciwpmcl[k] != ci || clwpmcl[k] != cl) bad.store(2);
        const int b = ngb[g];
        for (int a = 0; a < nband_arrays; ++a) {
          const double v = bandmcl[a][k];
          double& dst = out.band[a][o * nbnd + b];
          if (!seen[b]) dst = v;
          else if (dst != v) bad.store(2);
        }
        seen[b] = 1;
        ++ncloudy;
      }
      if (ncloudy) {
        // any value >= cldmin marks the layer as cloudy for cldprmc; the fraction itself is not used under McICA
        out.cldfr[o] = (double)ncloudy / (double)ngpt;
        out.ciwp[o] = ci;
        out.clwp[o] = cl;
      }
    }
  });
  if (bad.load() == 1)
    return "rrtmg_*_mcica_wrapper: sub-column cloud fractions must be 0 or 1 (what mcica_subcol_*_wrapper produces)";
  if (bad.load() == 2)
    return "rrtmg_*_mcica_wrapper: the cloudy sub-columns of a layer must carry one set of water paths / band optics "
           "(what mcica_subcol_*_wrapper produces)";
  return "";
}

inline void nan_fill(double* p, size_t n) {
  if (!p) return;
  const double q = std::nan("");
  for (size_t k = 0; k < n; ++k) p[k] = q;
}

}  // namespace mcica
}  // namespace cb
