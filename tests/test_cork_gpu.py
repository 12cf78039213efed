"""CORK correlated-k parity on the GPU: the compiled sm_100a kernels through the C ABI against golden vectors produced by
the reference's own numba kernels, against the oracle at other sizes, and the grey-limit identity of the reference's
tests/test_grey_limit.py.  Tolerance: 1e-6 relative on fluxes (BASELINE.json north_star); observed ~1e-13."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
RTOL = 1e-6
GOLD = np.load(H.os.path.join(H.HERE, "golden", "cork_reference.npz"))
CASES = ("clear", "cloudy", "d2")


def _state(case):
    return {k.split("/")[-1]: GOLD[k] for k in GOLD.files if k.startswith(case + "/in/")}


@pytest.fixture(scope="module")
def engines():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from climt_b200 import cork
    e = {"lw": cork.CorkEngine("earth_low_res_lw"), "sw": cork.CorkEngine("earth_low_res_sw")}
    yield e
    for x in e.values():
        x.close()


@pytest.mark.parametrize("case", CASES)
def test_cuda_cork_lw_matches_reference_kernels(engines, case):
    s = _state(case)
    nlev, ncol = s["T"].shape
    out = engines["lw"].lw_host(ncol, nlev, H.cork_arrays(s, "lw"), diffusivity_factor=float(s["diffusivity"]))
    for k in ("up_broad", "down_broad", "up_band", "down_band", "tau_band", "trans_band"):
        np.testing.assert_allclose(out[k], GOLD[f"{case}/lw/{k}"], rtol=RTOL, atol=1e-12, err_msg=k)
    np.testing.assert_allclose(out["heating_rate"], GOLD[f"{case}/lw/heating_rate"], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(out["hr_band"], GOLD[f"{case}/lw/hr_band"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("case", CASES)
def test_cuda_cork_sw_matches_reference_kernels(engines, case):
    s = _state(case)
    nlev, ncol = s["T"].shape
    out = engines["sw"].sw_host(ncol, nlev, H.cork_arrays(s, "sw"), earth_sun_factor=float(s["earth_sun_factor"][0]))
    for k in ("up_broad", "down_broad", "up_band", "down_band", "tau_band"):
        np.testing.assert_allclose(out[k], GOLD[f"{case}/sw/{k}"], rtol=RTOL, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(out["heating_rate"], GOLD[f"{case}/sw/heating_rate"], rtol=1e-5, atol=1e-12)
    assert np.all(out["up_broad"][:, 0] == 0.0)  # night column


def _big_state(ncol, nlev, seed=8):
    from climt_b200 import synthetic as SY
    rng = np.random.default_rng(seed)
    lw = SY.make_lw_state(ncol, nlev, seed=seed)
    return {"T": lw["tlay"], "p": lw["play"] * 100.0, "p_int": lw["plev"] * 100.0, "T_surf": lw["tsfc"], "q": lw["h2o"] * 0.622,
            "co2": np.full((nlev, ncol), 4e-4), "emissivity": rng.uniform(0.9, 1.0, (14, ncol)),
            "tau_cloud_lw": (rng.uniform(size=(nlev, ncol, 14)) < 0.05) * 1.5,
            "zenith": np.deg2rad(rng.uniform(0, 100, ncol)), "albedo": rng.uniform(0.05, 0.3, ncol),
            "earth_sun_factor": np.full(ncol, 0.98),
            "tau_cloud_sw": np.zeros((nlev, ncol, 3)), "ssa_cloud": np.zeros((nlev, ncol, 3)), "g_cloud": np.zeros((nlev, ncol, 3))}


def test_full_size_grid_subset_and_chunking(engines, monkeypatch):
    """8192 x 60 (BASELINE configs[1] grid): multi-chunk host pipeline == one device call bit-for-bit; a column subset
    equals the oracle; energy conservation of the broadband fluxes."""
    import torch
    from climt_b200 import cork
    from oracle import cork as OC
    ncol, nlev = 8192, 60
    s = _big_state(ncol, nlev)
    eng = engines["lw"]
    arrays = H.cork_arrays(s, "lw")
    host = eng.lw_host(ncol, nlev, arrays)          # default host chunk 4096 -> 2 chunks
    ins, outs = eng.shapes(ncol, nlev)
    dev_in = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in arrays.items()}
    dev_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in host}
    eng.lw_device(ncol, nlev, dev_in, dev_out)
    torch.cuda.synchronize()
    for k in host:
        np.testing.assert_array_equal(dev_out[k].cpu().numpy(), host[k], err_msg=k)
    idx = np.arange(0, ncol, 257)
    sub = {k: (v[..., idx] if v.ndim < 3 or k in ("emissivity",) else v[:, idx]) for k, v in s.items()}
    sub["emissivity"] = s["emissivity"][:, idx]
    ref = OC.lw_call(eng.table, sub, H.CORK_G, H.CORK_CPD, H.CORK_SIGMA)
    assert H.rel_err(host["up_broad"][:, idx], ref["up_broad"]) < RTOL
    assert H.rel_err(host["down_broad"][:, idx], ref["down_broad"]) < RTOL
    # per-band fluxes add up to the broadband ones
    np.testing.assert_allclose(host["up_band"].sum(axis=0), host["up_broad"], rtol=1e-12)
    # shortwave on the same grid
    engs = engines["sw"]
    hs = engs.sw_host(ncol, nlev, H.cork_arrays(s, "sw"), earth_sun_factor=0.98)
    refs = OC.sw_call(engs.table, sub, H.CORK_G, H.CORK_CPD)
    np.testing.assert_allclose(hs["down_broad"][:, idx], refs["down_broad"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(hs["up_broad"][:, idx], refs["up_broad"], rtol=RTOL, atol=1e-9)
    night = np.cos(s["zenith"]) <= 1e-4
    assert night.any() and np.all(hs["down_broad"][:, night] == 0.0)


@pytest.mark.parametrize("u", ["1", "2", "8"])
def test_unit_widths_agree(monkeypatch, u):
    from climt_b200 import cork
    monkeypatch.setenv("CLIMT_B200_CORK_U", u)
    monkeypatch.setenv("CLIMT_B200_HOST_CHUNK", "128")
    s = _state("cloudy")
    nlev, ncol = s["T"].shape
    eng = cork.CorkEngine("earth_low_res_lw")
    out = eng.lw_host(ncol, nlev, H.cork_arrays(s, "lw"))
    eng.close()
    np.testing.assert_allclose(out["up_broad"], GOLD["cloudy/lw/up_broad"], rtol=RTOL)
    np.testing.assert_allclose(out["down_band"], GOLD["cloudy/lw/down_band"], rtol=RTOL, atol=1e-12)


def test_components_through_state_dicts():
    """CorkLongwaveRadiation / CorkShortwaveRadiation called like the reference components (3-d state, band-last outputs)."""
    from climt_b200 import cork
    from climt_b200.sympl_shim import DataArray
    s = _state("cloudy")
    nlev, ncol = s["T"].shape
    ny, nx = 2, ncol // 2

    def da(a, dims, units):
        return DataArray(a, dims, {"units": units})
    m3, i3 = ("mid_levels", "lat", "lon"), ("interface_levels", "lat", "lon")
    state = {
        "air_temperature": da(s["T"].reshape(nlev, ny, nx), m3, "degK"), "air_pressure": da(s["p"].reshape(nlev, ny, nx), m3, "Pa"),
        "air_pressure_on_interface_levels": da(s["p_int"].reshape(nlev + 1, ny, nx), i3, "Pa"),
        "surface_temperature": da(s["T_surf"].reshape(ny, nx), ("lat", "lon"), "degK"),
        "surface_longwave_emissivity": da(s["emissivity"].reshape(14, ny, nx), ("num_longwave_bands", "lat", "lon"), "dimensionless"),
        "specific_humidity": da(s["q"].reshape(nlev, ny, nx), m3, "kg/kg"),
        "mole_fraction_of_carbon_dioxide_in_air": da(s["co2"].reshape(nlev, ny, nx), m3, "mole/mole"),
        "longwave_optical_thickness_due_to_cloud": da(s["tau_cloud_lw"].reshape(nlev, ny, nx, 14), m3 + ("num_longwave_bands",), "dimensionless"),
        "zenith_angle": da(s["zenith"].reshape(ny, nx), ("lat", "lon"), "radians"),
        "surface_albedo_for_direct_shortwave": da(s["albedo"].reshape(ny, nx), ("lat", "lon"), "dimensionless"),
        "flux_adjustment_for_earth_sun_distance": da(s["earth_sun_factor"].reshape(ny, nx), ("lat", "lon"), "dimensionless"),
        "shortwave_optical_thickness_due_to_cloud": da(s["tau_cloud_sw"].reshape(nlev, ny, nx, 3), m3 + ("num_shortwave_bands",), "dimensionless"),
        "single_scattering_albedo_due_to_cloud": da(s["ssa_cloud"].reshape(nlev, ny, nx, 3), m3 + ("num_shortwave_bands",), "dimensionless"),
        "cloud_asymmetry_parameter": da(s["g_cloud"].reshape(nlev, ny, nx, 3), m3 + ("num_shortwave_bands",), "dimensionless"),
    }
    lw = cork.CorkLongwaveRadiation(optics="correlated_k", table="earth_low_res_lw")
    tend, diag = lw(state)
    np.testing.assert_allclose(diag["upwelling_longwave_flux_in_air"].values.reshape(nlev + 1, ncol), GOLD["cloudy/lw/up_broad"], rtol=RTOL)
    per_band = diag["downwelling_longwave_flux_in_air_per_band"].values            # (lev, lat, lon, band)
    np.testing.assert_allclose(np.moveaxis(per_band.reshape(nlev + 1, ncol, 14), -1, 0), GOLD["cloudy/lw/down_band"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(tend["air_temperature"].values.reshape(nlev, ncol), GOLD["cloudy/lw/heating_rate"], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(diag["air_temperature_tendency_from_longwave"].values, tend["air_temperature"].values * 86400.0)
    sw = cork.CorkShortwaveRadiation(optics="correlated_k", table="earth_low_res_sw")
    tend, diag = sw(state)
    np.testing.assert_allclose(diag["downwelling_shortwave_flux_in_air"].values.reshape(nlev + 1, ncol), GOLD["cloudy/sw/down_broad"], rtol=RTOL, atol=1e-9)
    with pytest.raises(ValueError):
        cork.CorkLongwaveRadiation(optics="line_by_line")   # cork/lw/component.py:60-61
    # diagnostics_level >= 1 adds the reference's extra diagnostics (cork/lw/component.py:189-202); values: tests/test_cork_esft_diag.py
    lw1 = cork.CorkLongwaveRadiation(optics="correlated_k", table="earth_low_res_lw", diagnostics_level=1)
    assert {"lw_layer_transmittance", "lw_up_per_gpoint", "lw_down_per_gpoint"} <= set(lw1.diagnostic_properties)


def test_grey_limit_matches_gray_engine():
    """tests/test_grey_limit.py of the reference: one band, one g-point, constant k, planck_fraction = 1 -> the CORK sweep
    equals the grey sweep fed with the cumulative D*tau profile (rtol 1e-10 there)."""
    from climt_b200 import cork
    from climt_b200.gray import gray_lw_host
    s = _state("clear")
    nlev, ncol = s["T"].shape
    tbl = cork.load_k_table("single_band_unit_lw")
    eng = cork.CorkEngine(tbl)
    names = cork.table_flags(tbl)[0]
    arrays = {"T": s["T"], "p": s["p"], "p_int": s["p_int"], "T_surf": s["T_surf"], "emissivity": np.ones((1, ncol)),
              "gas_q": (np.full((1, nlev, ncol), 4e-4) * (cork.MOLAR_MASS.get(names[0], cork.MOLAR_MASS_DRY_AIR) / cork.MOLAR_MASS_DRY_AIR))}
    out = eng.lw_host(ncol, nlev, arrays)
    eng.close()
    tau_int = np.zeros((nlev + 1, ncol))
    tau_int[1:] = np.cumsum(1.66 * out["tau_band"][0], axis=0)
    down, up, tend = gray_lw_host(s["T"], s["p_int"], s["T_surf"], tau_int, H.CORK_SIGMA, H.CORK_G, H.CORK_CPD)
    np.testing.assert_allclose(out["up_broad"], up, rtol=1e-10)
    np.testing.assert_allclose(out["down_broad"], down, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["heating_rate"], tend, rtol=1e-8, atol=1e-14)


def test_missing_inputs_are_reported():
    from climt_b200 import cork
    s = _state("clear")
    nlev, ncol = s["T"].shape
    eng = cork.CorkEngine("earth_low_res_lw")
    a = H.cork_arrays(s, "lw")
    a.pop("co2_vmr")
    with pytest.raises(ValueError, match="co2_vmr_grid axis but co2_vmr was not provided"):
        eng.lw_host(ncol, nlev, a)
    sw_in = {k: v for k, v in H.cork_arrays(s, "sw").items() if k not in ("tau_cloud", "ssa_cloud", "g_cloud")}
    with pytest.raises(ValueError, match="not a shortwave table"):
        eng.sw_host(ncol, nlev, sw_in)
    eng.close()
