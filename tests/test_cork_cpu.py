"""CORK correlated-k path on the CPU: (1) the oracle (oracle/cork_oracle.cpp + oracle/cork.py glue) against golden vectors
produced by the reference's own numba kernels (tests/golden/make_cork_golden.py); (2) the CUDA engine's per-thread code,
compiled for the host, against the same goldens and the oracle for every unit width and table flavour."""
import numpy as np
import pytest

import helpers as H

GOLD = np.load(H.os.path.join(H.HERE, "golden", "cork_reference.npz"))
CASES = ("clear", "cloudy", "d2")
LW_KEYS = {"up_broad": "up_broad", "down_broad": "down_broad", "heating_rate": "heating_rate", "up_band": "up_band",
           "down_band": "down_band", "tau_band": "tau_band", "trans_band": "trans_band", "hr_band": "hr_band"}


def _state(case):
    return {k.split("/")[-1]: GOLD[k] for k in GOLD.files if k.startswith(case + "/in/")}


def _close(a, b, rtol, atol=0.0):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_numba_kernels(case):
    """bit-for-bit in practice (both sides are IEEE double with libm exp/log); asserted at 1e-13"""
    from oracle import cork as OC
    s = _state(case)
    lw = OC.lw_call(H.cork_table("earth_low_res_lw"), s, H.CORK_G, H.CORK_CPD, H.CORK_SIGMA, D=float(s["diffusivity"]))
    for k, v in lw.items():
        _close(v, GOLD[f"{case}/lw/{k}"], 1e-13)
    sw = OC.sw_call(H.cork_table("earth_low_res_sw"), s, H.CORK_G, H.CORK_CPD)
    for k, v in sw.items():
        _close(v, GOLD[f"{case}/sw/{k}"], 1e-13, atol=1e-300)


@pytest.mark.parametrize("umax", [1, 2, 4, 8])
@pytest.mark.parametrize("case", CASES)
def test_kernel_code_lw_matches_reference(case, umax):
    s = _state(case)
    out = H.run_cork_emul(H.cork_table("earth_low_res_lw"), "lw", H.cork_arrays(s, "lw"), float(s["diffusivity"]), umax=umax)
    for k in LW_KEYS:
        _close(out[k], GOLD[f"{case}/lw/{k}"], 1e-12, atol=1e-13)


@pytest.mark.parametrize("umax", [1, 2])
@pytest.mark.parametrize("case", CASES)
def test_kernel_code_sw_matches_reference(case, umax):
    s = _state(case)
    out = H.run_cork_emul(H.cork_table("earth_low_res_sw"), "sw", H.cork_arrays(s, "sw"), float(s["earth_sun_factor"][0]), umax=umax)
    for k in ("up_broad", "down_broad", "up_band", "down_band", "tau_band"):
        _close(out[k], GOLD[f"{case}/sw/{k}"], 1e-12, atol=1e-12)
    _close(out["heating_rate"], GOLD[f"{case}/sw/heating_rate"], 1e-9, atol=1e-16)
    _close(out["hr_band"], GOLD[f"{case}/sw/hr_band"], 1e-9, atol=1e-11)
    assert np.all(out["up_broad"][:, 0] == 0) and np.all(out["down_broad"][:, 0] == 0)  # night column (mu0 <= 1e-4)


def test_float64_5d_tables_against_oracle():
    """the reference's small test tables: float64 k, (T, P) axes only, non-premixed single gas"""
    from oracle import cork as OC
    rng = np.random.default_rng(2)
    nlev, ncol = 9, 7
    p_int = np.linspace(1.0e5, 50.0, nlev + 1)[:, None] * rng.uniform(0.95, 1.0, (1, ncol))
    p = 0.5 * (p_int[1:] + p_int[:-1])
    T = 220.0 + 70.0 * (p / 1e5) ** 0.3 + rng.normal(0, 1, (nlev, ncol))
    q = 0.01 * (p / 1e5) ** 2
    for name, which in (("test_2band_lw", "lw"), ("test_2band_sw", "sw")):
        tbl = H.cork_table(name)
        nb = tbl["k_coefficients"].shape[1]
        gas = OC.column_amount(q, p_int, H.CORK_G)[None]
        tau = OC.optical_depth(tbl, T, p, gas)
        w = np.asarray(tbl["gpoint_weights"], dtype=np.float64)
        arrays = {"T": T, "p": p, "p_int": p_int, "gas_q": q[None]}
        if which == "lw":
            Ts = T[0] + 2.0
            em = rng.uniform(0.9, 1.0, (nb, ncol))
            ps, ss = OC.planck_sources(tbl, T, Ts, H.CORK_SIGMA, nb, w.shape[1])
            ub, db, u, d = OC.lw_transport(tau, ps, ss, em, w, 1.66)
            arrays.update(T_surf=Ts, emissivity=em)
            out = H.run_cork_emul(tbl, "lw", arrays, 1.66)
        else:
            zen, alb = np.deg2rad(rng.uniform(0, 80, ncol)), rng.uniform(0.1, 0.3, ncol)
            ray = tbl["rayleigh_coefficient"]
            tau_ray = ray[:, None, None, None] * (np.abs(np.diff(p_int, axis=0)) / H.CORK_G)[None, None]
            tot = tau + tau_ray
            ssa = np.where(tot > 0, tau_ray / tot, 0.0)
            ub, db, u, d = OC.sw_two_stream(tot, ssa, np.zeros_like(tot), zen, alb, np.asarray(tbl["solar_source_per_gpoint"]) * 1.01, w)
            arrays.update(zenith=zen, albedo=alb)
            out = H.run_cork_emul(tbl, "sw", arrays, 1.01)
        _close(out["up_broad"], u, 1e-12)
        _close(out["down_broad"], d, 1e-12)
        _close(out["up_band"], ub, 1e-12)
        _close(out["down_band"], db, 1e-12)


def test_table_loader_and_flags():
    from climt_b200 import cork
    lw = cork.load_k_table("earth_low_res_lw")
    assert lw["k_coefficients"].shape == (1, 14, 8, 12, 8, 7, 10) and lw["k_coefficients"].dtype == np.float32
    names, has_h2o, has_co2, fully, bg = cork.table_flags(lw)
    assert names == ["effective"] and has_h2o and has_co2 and not fully and bg
    sw = cork.load_k_table("earth_low_res_sw")
    assert cork.table_flags(sw)[4] and not cork.table_flags(sw)[2]
    with pytest.raises(FileNotFoundError):
        cork.load_k_table("no_such_table")
    ct, _keep = cork.make_ctable(lw)
    assert (ct.ngas, ct.nband, ct.ngpt, ct.nT, ct.nP, ct.nX, ct.nC, ct.premixed, ct.co2_logk) == (1, 14, 8, 12, 8, 7, 10, 1, 1)
