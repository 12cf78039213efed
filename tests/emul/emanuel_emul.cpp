// TEST-ONLY host emulation of the CUDA Emanuel engine: the warp-cooperative column code of the kernel compiled for the host,
// where a "warp" is one lane that walks every lane-parallel loop serially (emanuel_core.cuh: CB_LANES_FOR).
#include <cstring>
#include <vector>

#include "../../include/climt_b200.h"
#include "../../climt_b200/csrc/emanuel_core.cuh"

using namespace cb::emanuel;

// layout 1: arrays in the component's (ncol, nlev[+1]) layout; layout 0: (nlev[+1], ncol).  Read in place through strides, as the kernel does.
extern "C" int emul_emanuel_run(const cb200_emanuel_params* p, int ncol, int nlev, int NL, double dt, int qs_mode, int layout,
                                const cb200_emanuel_inputs* in, const cb200_emanuel_outputs* out) {
  Par par;
  std::memcpy(&par, p, sizeof(Par));
  const size_t n = (size_t)ncol, L = (size_t)nlev;
  In ni{};
  ni.nlev = nlev;
  if (layout == 0) { ni.ls = n; ni.cs = 1; ni.ls_i = n; ni.cs_i = 1; }
  else { ni.ls = 1; ni.cs = L; ni.ls_i = 1; ni.cs_i = L + 1; }
  ni.t = in->t; ni.q = in->q; ni.u = in->u; ni.v = in->v; ni.p = in->p; ni.ph = in->ph; ni.qs = qs_mode == QS_GIVEN ? in->qs : nullptr;
  ni.cbmf = in->cbmf; ni.qs_mode = qs_mode;
  Out no{};
  if (layout == 0) { no.ls = n; no.cs = 1; }
  else { no.ls = 1; no.cs = L; }
  no.ft = out->ft; no.fq = out->fq; no.fu = out->fu; no.fv = out->fv; no.precip = out->precip; no.wd = out->wd; no.tprime = out->tprime;
  no.qprime = out->qprime; no.cbmf = out->cbmf; no.cape = out->cape; no.iflag = out->iflag;
  Work W;
  W.n1 = nlev + 4; W.nm = NL + 2;
  const double nan = 0.0 / 0.0;
  std::vector<double> sv((size_t)V_COUNT * W.n1), wm((size_t)M_COUNT * W.nm * W.nm);
  W.m = wm.data();
  for (int c = 0; c < ncol; ++c) {
    // poison the shared-memory slice and the matrix slot: whatever the code reads must have been written by it for this column
    std::fill(sv.begin(), sv.end(), nan);
    std::fill(wm.begin(), wm.end(), nan);
    convect_warp(par, ni, W, no, (size_t)c, 0, sv.data(), wm.data(), NL, dt);
  }
  return 0;
}
