// TEST-ONLY host emulation of the CUDA Emanuel engine (same per-thread code as the kernel, stepped serially).
#include <cstring>
#include <vector>

#include "../../include/climt_b200.h"
#include "../../climt_b200/csrc/emanuel_core.cuh"

using namespace cb::emanuel;

// arrays in the component's (ncol, nlev[+1]) layout; transposed here the way the engine's k_transpose does
extern "C" int emul_emanuel_run(const cb200_emanuel_params* p, int ncol, int nlev, int NL, double dt, int qs_mode,
                                const cb200_emanuel_inputs* in, const cb200_emanuel_outputs* out) {
  Par par;
  std::memcpy(&par, p, sizeof(Par));
  const size_t n = (size_t)ncol, L = (size_t)nlev;
  auto tr = [&](const double* src, size_t per) {
    std::vector<double> d(per * n);
    for (size_t c = 0; c < n; ++c)
      for (size_t k = 0; k < per; ++k) d[k * n + c] = src[c * per + k];
    return d;
  };
  const std::vector<double> t = tr(in->t, L), q = tr(in->q, L), u = tr(in->u, L), v = tr(in->v, L), pp = tr(in->p, L), ph = tr(in->ph, L + 1);
  std::vector<double> qs;
  if (qs_mode == QS_GIVEN) qs = tr(in->qs, L);
  std::vector<double> ft(L * n), fq(L * n), fu(L * n), fv(L * n);
  In ni{nlev, n, t.data(), q.data(), u.data(), v.data(), pp.data(), ph.data(), qs_mode == QS_GIVEN ? qs.data() : nullptr, in->cbmf, qs_mode};
  Out no{n, ft.data(), fq.data(), fu.data(), fv.data(), out->precip, out->wd, out->tprime, out->qprime, out->cbmf, out->cape, out->iflag};
  Work W;
  W.ncc = ncol; W.n1 = nlev + 3; W.nm = NL + 2;
  // poison the workspace: whatever the kernel reads must have been written by it
  const double nan = 0.0 / 0.0;
  std::vector<double> wv((size_t)V_COUNT * W.n1 * n, nan), wm((size_t)M_COUNT * W.nm * W.nm * n, nan);
  W.v = wv.data(); W.m = wm.data();
  for (int c = 0; c < ncol; ++c) convect_column(par, ni, W, no, 0, c, NL, dt);
  auto back = [&](double* dst, const std::vector<double>& s) {
    for (size_t c = 0; c < n; ++c)
      for (size_t k = 0; k < L; ++k) dst[c * L + k] = s[k * n + c];
  };
  back(out->ft, ft); back(out->fq, fq); back(out->fu, fu); back(out->fv, fv);
  return 0;
}
