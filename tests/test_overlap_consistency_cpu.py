"""An independent pin for `rtrnmr`, which the reference holds no golden for.

The reference has two routes to maximum-random cloud overlap: the deterministic `rtrnmr` (rrtmg_lw_rtrnmr.f90, icld = 2 without
McICA) and McICA -- the sub-column generator drawing maximum-random masks (mcica_subcol_gen_lw.f90) followed by `rtrnmc`.  The
second route IS pinned by the reference's MCICA goldens (tests/test_lw_emul.py::test_oracle_mcica_golden).  They must agree in the
mean: the ensemble average of McICA fluxes over many seeds converges to the overlap assumption's flux.  The same comparison for
random overlap (rtrn, golden-pinned, vs McICA icld = 1) calibrates what "agree" means at this sample size; and the two overlap
assumptions themselves differ by several times that, so the check can tell them apart."""
import numpy as np

import helpers as H
from climt_b200 import synthetic as SY
from oracle.rrtmg import lw_mcica

NSAMPLES = 240


def _ensemble(icld, st):
    orc = H.lw_oracle(cloud_overlap=icld)
    acc = {"uflx": 0.0, "dflx": 0.0}
    for seed in range(NSAMPLES):
        o = lw_mcica(orc, st, 5000 + seed, irng=1)
        for k in acc:
            acc[k] = acc[k] + o[k]
    return {k: v / NSAMPLES for k, v in acc.items()}


def test_rtrnmr_agrees_with_the_golden_pinned_mcica_route_in_the_mean():
    st = SY.make_lw_state(10, 40, seed=41, clouds=True)
    assert ((st["cldfr"] > 0.05) & (st["cldfr"] < 0.95)).sum() > 30        # partial cloud in stacked layers: overlap matters
    det = {icld: H.run_lw_oracle(H.lw_oracle(cloud_overlap=icld), st) for icld in (1, 2)}
    scale = float(det[2]["uflx"].max())
    err = {}
    for icld in (1, 2):
        m = _ensemble(icld, st)
        err[icld] = max(float(np.abs(m[k] - det[icld][k]).max()) for k in ("uflx", "dflx")) / scale
    apart = max(float(np.abs(det[2][k] - det[1][k]).max()) for k in ("uflx", "dflx")) / scale
    assert err[1] < 0.006, err                      # control: rtrn vs McICA(random), both golden-pinned -- sampling noise only
    assert err[2] < 0.008, err                      # rtrnmr vs McICA(maximum-random)
    assert err[2] < 2.5 * err[1] + 0.002, err       # ... no worse than the control allows
    assert apart > 3 * err[2], (apart, err)         # and the two overlap assumptions are much further apart than that
