"""Reference-held checks that reach a HUMID column.

Every RRTMG golden of the reference is a dry column (`specific_humidity` defaults to 0, climt/_core/initialization.py:755), so
the water-vapour branches of taumol (self / foreign continuum, the H2O key-species interpolation) are reached by no golden
vector.  What the reference does hold for a humid column is restated here:

* tests/test_rrtmg_comparison.py:27-108 -- CORK (correlated-k, 14 LW / 3 SW bands; pinned BIT-EXACTLY to the reference's own
  numba kernels, tests/golden/cork_reference.npz) against RRTMG: upward LW flux / downward SW flux within 25 % at every level,
  and the OLR falls by >= 5 W m-2 from q = 1e-6 to q = 1e-2.  Run on the default (dry) state as the reference does, on the
  reference's two uniform humidities, and on a tropical-like profile.  The two schemes share no spectroscopic table, so their
  agreement (observed: <1 % in the upward LW flux, ~3 % in the SW) bounds any gross error of the humid branches.
* tests/test_conservation.py:55-70,135-176,225-283 -- the column enthalpy tendency of the heating rates equals the net flux
  through the column's boundaries (atol 1e-3 W m-2 there) -- on humid, cloudy columns, deterministic and McICA.

CPU half: the oracles (what the CUDA path is compared with everywhere else).  GPU half (-m gpu): the CUDA engines themselves.
"""
import numpy as np
import pytest

import helpers as H
from climt_b200 import constants as C, state as S, synthetic as SY

NZ = 30
Q_CASES = {"default_dry": 0.0, "dry_1e-6": 1e-6, "tropical_profile": None, "moist_1e-2": 1e-2}


def _states():
    n = len(Q_CASES)
    lw = H.default_lw_abi_state(NZ, n)
    sw = H.default_sw_abi_state(NZ, n)
    p, ps = lw["play"], lw["plev"][0]
    q = np.zeros_like(p)
    for i, v in enumerate(Q_CASES.values()):
        q[:, i] = (0.018 * (p[:, i] / ps[i]) ** 3 + 1e-6) if v is None else v
    lw["h2o"] = S.mass_to_volume_mixing_ratio(q, 18.02)
    sw["h2o"] = lw["h2o"].copy()
    sw["coszen"][:] = np.cos(np.pi / 4)          # tests/test_rrtmg_comparison.py:58
    cork_lw = {"T": lw["tlay"], "p": lw["play"] * 100.0, "p_int": lw["plev"] * 100.0, "T_surf": lw["tsfc"], "q": q,
               "co2": lw["co2"], "emissivity": np.ones((14, n)), "tau_cloud_lw": np.zeros((NZ, n, 14))}
    cork_sw = {"T": sw["tlay"], "p": sw["play"] * 100.0, "p_int": sw["plev"] * 100.0, "T_surf": sw["tsfc"], "q": q,
               "co2": sw["co2"], "zenith": np.full(n, np.pi / 4), "albedo": sw["asdir"].copy(), "earth_sun_factor": np.ones(n),
               "tau_cloud_sw": np.zeros((NZ, n, 3)), "ssa_cloud": np.zeros((NZ, n, 3)), "g_cloud": np.zeros((NZ, n, 3))}
    return lw, sw, cork_lw, cork_sw


def _check(up_rr, up_ck, dn_rr, dn_ck):
    names = list(Q_CASES)
    rel = np.abs(up_ck - up_rr) / np.maximum(np.abs(up_rr), 1e-3)
    assert rel.max() < 0.25, rel.max(axis=0)                     # the reference's bar (test_rrtmg_comparison.py:44-47)
    assert rel.max() < 0.02, rel.max(axis=0)                     # what the two schemes actually achieve, humid columns included
    rel_sw = np.abs(dn_ck - dn_rr) / np.maximum(np.abs(dn_rr), 1e-3)
    assert rel_sw.max() < 0.25 and rel_sw.max() < 0.06, rel_sw.max(axis=0)
    dry, moist = names.index("dry_1e-6"), names.index("moist_1e-2")
    drop_rr, drop_ck = up_rr[-1, dry] - up_rr[-1, moist], up_ck[-1, dry] - up_ck[-1, moist]
    assert drop_ck > 5.0 and drop_rr > 5.0                       # test_rrtmg_comparison.py:103-108
    assert abs(drop_rr - drop_ck) < 0.2 * drop_ck                # and RRTMG's humidity response is CORK's (34 vs 38 W m-2)
    # humidity must matter for the shortwave too: near-infrared water-vapour absorption
    assert dn_rr[0, dry] - dn_rr[0, moist] > 100.0 and dn_ck[0, dry] - dn_ck[0, moist] > 100.0


def test_oracle_rrtmg_agrees_with_reference_cork_on_humid_columns():
    from climt_b200 import cork
    from oracle import cork as OC
    lw, sw, ck_lw, ck_sw = _states()
    rr = H.run_lw_oracle(H.lw_oracle(cloud_overlap=0), lw)
    rrs = H.sw_oracle(cloud_overlap=0)(sw, adjes=1.0, dyofyr=0)
    o_lw = OC.lw_call(cork.load_k_table("earth_low_res_lw"), ck_lw, H.CORK_G, H.CORK_CPD, H.CORK_SIGMA)
    o_sw = OC.sw_call(cork.load_k_table("earth_low_res_sw"), ck_sw, H.CORK_G, H.CORK_CPD)
    _check(rr["uflx"], o_lw["up_broad"], rrs["swdflx"], o_sw["down_broad"])


@pytest.mark.gpu
def test_cuda_rrtmg_agrees_with_cuda_cork_on_humid_columns():
    from climt_b200 import cork
    from climt_b200.engine import LWEngine, SWEngine
    lw, sw, ck_lw, ck_sw = _states()
    n = len(Q_CASES)
    e = LWEngine(icld=0)
    rr = e.run_host(n, NZ, H.to_abi(lw))
    e.close()
    es = SWEngine(icld=0)
    rrs = es.run_host(n, NZ, H.to_abi_sw(sw), dyofyr=0)
    es.close()
    ce = cork.CorkEngine("earth_low_res_lw")
    o_lw = ce.lw_host(n, NZ, H.cork_arrays(ck_lw, "lw"))
    ce.close()
    ce = cork.CorkEngine("earth_low_res_sw")
    o_sw = ce.sw_host(n, NZ, H.cork_arrays(ck_sw, "sw"), earth_sun_factor=1.0)
    ce.close()
    _check(rr["uflx"], o_lw["up_broad"], rrs["dflx"], o_sw["down_broad"])


# ---- column energy budget (tests/test_conservation.py): sum over layers of cp * dT/dt * dp / g == net flux in - net flux out ------
def _budget_residual(flux_up, flux_dn, hr_per_day, plev_hpa):
    k = C.rrtmg_constants()
    dp = (plev_hpa[:-1] - plev_hpa[1:]) * 100.0
    heating = np.sum(hr_per_day / 86400.0 * k["cpdair"] * dp / k["grav"], axis=0)         # W m-2
    net = flux_dn - flux_up                                                                # downward positive
    return heating - (net[-1] - net[0])


def test_oracle_heating_rates_close_the_column_energy_budget_on_humid_cloudy_columns():
    st = SY.make_lw_state(24, 40, seed=5, clouds=True, aerosol=True)
    o = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), st)
    assert np.abs(_budget_residual(o["uflx"], o["dflx"], o["hr"], st["plev"])).max() < 1e-3
    assert np.abs(_budget_residual(o["uflxc"], o["dflxc"], o["hrc"], st["plev"])).max() < 1e-3
    sts = SY.make_sw_state(24, 40, seed=5, clouds=True)
    os_ = H.sw_oracle()(sts, dyofyr=100)
    assert np.abs(_budget_residual(os_["swuflx"], os_["swdflx"], os_["swhr"], sts["plev"])).max() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("mcica", [False, True])
def test_cuda_heating_rates_close_the_column_energy_budget(mcica):
    from climt_b200.engine import LWEngine, SWEngine
    ncol, nlay = 3000, 60
    st = SY.make_lw_state(ncol, nlay, seed=15, clouds=True, aerosol=True)
    assert st["h2o"].max() > 1e-2                                       # humid columns
    e = LWEngine(icld=2 if mcica else 1, mcica=mcica, irng=0, permuteseed=7)
    o = e.run_host(ncol, nlay, H.to_abi(st))
    e.close()
    assert np.abs(_budget_residual(o["uflx"], o["dflx"], o["hr"], st["plev"])).max() < 1e-3
    assert np.abs(_budget_residual(o["uflxc"], o["dflxc"], o["hrc"], st["plev"])).max() < 1e-3
    sts = SY.make_sw_state(ncol, nlay, seed=15, clouds=True, overcast_only=not mcica)
    es = SWEngine(icld=2 if mcica else 1, mcica=mcica, irng=0, permuteseed=7)
    os_ = es.run_host(ncol, nlay, H.to_abi_sw(sts), dyofyr=100)
    es.close()
    assert np.abs(_budget_residual(os_["uflx"], os_["dflx"], os_["hr"], sts["plev"])).max() < 1e-3
    assert np.abs(_budget_residual(os_["uflxc"], os_["dflxc"], os_["hrc"], sts["plev"])).max() < 1e-3
