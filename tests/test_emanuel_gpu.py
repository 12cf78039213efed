"""Emanuel convection on the GPU, through the C ABI and the drop-in components: against golden vectors
produced by the reference's numba port (tests/golden/make_emanuel_golden.py), against the oracle (a restatement of the Fortran)
with the Fortran component's settings, at BASELINE.json configs[4]'s grid size, and the device-pointer path in both layouts.
Tolerance 1e-6 relative to each field's scale (BASELINE.json), convective_state exact; observed ~1e-13."""
import datetime

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
RTOL = 1e-6
CASES = ("tropics60", "tropics30_longstep", "levels72")
SYMPL = dict(cpd=1004.64, cpv=1846.0, cl=2500.0, rv=461.5, rd=287.0, lv0=2.5e6, g=9.80665, rowl=1e3)


@pytest.fixture(scope="module")
def gold():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return np.load(H.EMANUEL_GOLDEN)


def _flat(tend, diag):
    return {"ft": tend["air_temperature"], "fq": tend["specific_humidity"], "fu": tend["eastward_wind"], "fv": tend["northward_wind"],
            "iflag": diag["convective_state"], "precip": diag["convective_precipitation_rate"],
            "wd": diag["convective_downdraft_velocity_scale"], "tprime": diag["convective_downdraft_temperature_scale"],
            "qprime": diag["convective_downdraft_specific_humidity_scale"], "cbmf": diag["cloud_base_mass_flux"],
            "cape": diag["atmosphere_convective_available_potential_energy"]}


@pytest.mark.parametrize("case", CASES)
def test_python_drop_in_matches_the_reference_numba_port(gold, case):
    from climt_b200 import emanuel
    st = H.emanuel_case(gold, case)
    comp = emanuel.EmanuelConvectionPython()
    tend, diag = comp.array_call({k: v for k, v in st.items() if k in comp.input_properties}, datetime.timedelta(seconds=float(st["timestep"])))
    ref = {k: gold[f"{case}/out/{n}"] for k, n in H.EMANUEL_OUT.items()}
    H.emanuel_compare(_flat(tend, diag), ref, RTOL, case)
    assert diag["convective_state"].dtype == np.int32
    np.testing.assert_array_equal(diag["air_temperature_tendency_from_convection"], tend["air_temperature"] * 86400.0)
    assert set(diag) == set(comp.diagnostic_properties) and set(tend) == set(comp.tendency_properties)


@pytest.mark.parametrize("case", CASES[:2])
def test_fortran_drop_in_matches_oracle(gold, case):
    """climt.EmanuelConvection: constructor keywords -> CONVECT parameters, sympl constants, bolton_q_sat, 273.0 K switch."""
    from climt_b200 import emanuel
    from oracle import emanuel as OE
    st = H.emanuel_case(gold, case)
    dt = float(st["timestep"])
    comp = emanuel.EmanuelConvection(minimum_convecting_layer=2, entrainment_mixing_coefficient=1.2, downdraft_area_fraction=0.08,
                                     precipitation_fraction_outside_cloud=0.2, convective_momentum_transfer_coefficient=0.5,
                                     convection_bouyancy_threshold=0.6, mass_flux_relaxation_rate=0.2, mass_flux_damping_rate=0.15,
                                     autoconversion_water_content_threshold=0.0009)
    tend, diag = comp.array_call({k: v for k, v in st.items() if k in comp.input_properties}, datetime.timedelta(seconds=dt))
    tweak = dict(minorig=2, entp=1.2, sigd=0.08, sigs=0.2, cu=0.5, dtmax=0.6, alpha=0.2, damp=0.15, elcrit=0.0009)
    rt, rd = OE.fortran_component_call(st, dt, SYMPL, par=tweak)
    H.emanuel_compare(_flat(tend, diag), _flat(rt, rd), RTOL, case)
    with pytest.raises(ValueError):
        emanuel.EmanuelConvection(downdraft_area_fraction=1.5)


def test_three_dimensional_state_through_call(gold):
    """component(state, timestep) on a (lat, lon, lev) DataArray state: '*' flattening and re-wrapping."""
    from climt_b200 import emanuel
    from climt_b200.sympl_shim import DataArray
    st = H.emanuel_case(gold, "tropics60")
    ncol, nlev = st["air_temperature"].shape
    ny, nx = 8, ncol // 8
    units = {k: v["units"] for k, v in emanuel.EmanuelConvection.input_properties.items()}

    def da(name):
        a = st[name]
        if a.ndim == 2:
            lev = "interface_levels" if a.shape[1] == nlev + 1 else "mid_levels"
            return DataArray(a.reshape(ny, nx, a.shape[1]), ("lat", "lon", lev), {"units": units[name]})
        return DataArray(a.reshape(ny, nx), ("lat", "lon"), {"units": units[name]})

    state = {k: da(k) for k in units}
    comp = emanuel.EmanuelConvectionPython()
    tend, diag = comp(state, datetime.timedelta(seconds=float(st["timestep"])))
    assert tend["air_temperature"].values.shape == (ny, nx, nlev)
    ref = gold["tropics60/out/tendency_air_temperature"]
    np.testing.assert_allclose(tend["air_temperature"].values.reshape(ncol, nlev), ref, rtol=0, atol=RTOL * np.abs(ref).max())
    np.testing.assert_array_equal(diag["convective_state"].values.reshape(ncol), gold["tropics60/out/convective_state"])


def test_gmd_grid_size_and_device_layouts(gold):
    """64 800 columns x 60 levels (BASELINE.json configs[4]: 360 x 180 x 60): host pipeline (several chunks) == device-pointer call
    in the component's layout == device-pointer call in the radiation engines' (level, column) layout, bit for bit; a column
    subset equals the oracle; the scheme's enthalpy closure holds on every convecting column."""
    import torch
    from climt_b200 import emanuel, synthetic as SY
    from oracle import emanuel as OE
    ncol, nlev, dt = 64800, 60, 1200.0
    st = SY.make_emanuel_state(ncol, nlev, seed=11)
    par = dict(OE.FORTRAN_DEFAULTS, **SYMPL)
    eng = emanuel.EmanuelEngine(par)
    arrays = H.emanuel_arrays(st)
    host = eng.run_host(arrays, dt, qs_mode=emanuel.QS_BOLTON)
    assert eng.last_launches >= 2      # several pipeline chunks, one kernel launch each
    sub = slice(0, 2048)
    sst = {k: v[sub] for k, v in st.items()}
    rt, rd = OE.fortran_component_call(sst, dt, SYMPL)
    ref = {"ft": rt["air_temperature"], "fq": rt["specific_humidity"], "fu": rt["eastward_wind"], "fv": rt["northward_wind"],
           "iflag": rd["convective_state"], "precip": rd["convective_precipitation_rate"], "wd": rd["convective_downdraft_velocity_scale"],
           "tprime": rd["convective_downdraft_temperature_scale"], "qprime": rd["convective_downdraft_specific_humidity_scale"],
           "cbmf": rd["cloud_base_mass_flux"], "cape": rd["atmosphere_convective_available_potential_energy"]}
    H.emanuel_compare({k: v[sub] for k, v in host.items()}, ref, RTOL, "subset")
    frac = {int(f): float((host["iflag"] == f).mean()) for f in np.unique(host["iflag"])}
    assert frac.get(1, 0) > 0.3, frac
    for layout in (1, 0):
        ins, outs = eng.shapes(ncol, nlev, layout)
        tens = {k: torch.from_numpy(np.ascontiguousarray(v if layout == 1 or v.ndim == 1 else v.T)).cuda() for k, v in arrays.items()}
        out = {k: torch.empty(outs[k], dtype=torch.int32 if k == "iflag" else torch.float64, device="cuda") for k in outs}
        eng.run_device(ncol, nlev, tens, out, dt, qs_mode=emanuel.QS_BOLTON, layout=layout)
        torch.cuda.synchronize()
        for k in host:
            got = out[k].cpu().numpy()
            if layout == 0 and got.ndim == 2:
                got = got.T
            assert np.array_equal(got, host[k]), (layout, k)
    conv = host["iflag"] >= 1
    dp = st["air_pressure_on_interface_levels"][:, :-1] - st["air_pressure_on_interface_levels"][:, 1:]
    q, T = st["specific_humidity"], st["air_temperature"]
    cpn = par["cpd"] * (1 - q) + par["cpv"] * q
    lv = par["lv0"] - (par["cl"] - par["cpv"]) * (T - 273.15)
    ents = ((cpn * host["ft"] + lv * host["fq"]) * dp).sum(axis=1)
    scale = (np.abs(cpn * host["ft"]) * dp).sum(axis=1) + 1e-30
    assert np.all(np.abs(ents[conv]) / scale[conv] < 1e-8)
    eng.close()


def test_reference_named_symbols_one_column(gold):
    """init_emanuel_convection_fortran + emanuel_convection (bind(c) names of convect43c.f90) on single columns."""
    import ctypes
    from climt_b200 import _native
    from oracle import emanuel as OE
    L = _native.lib()
    st = H.emanuel_case(gold, "tropics30_longstep")
    dt = float(st["timestep"])
    par = dict(OE.FORTRAN_DEFAULTS, **SYMPL)
    ci, cd = ctypes.c_int, ctypes.c_double
    order = ("elcrit", "tlcrit", "entp", "sigd", "sigs", "omtrain", "omtsnow", "coeffr", "coeffs", "cu", "dtmax", "beta", "alpha", "damp",
             "cpd", "cpv", "cl", "rv", "rd", "lv0", "g", "rowl", "delt0")
    args = [ctypes.byref(ci(0)), ctypes.byref(ci(1))] + [ctypes.byref(cd(par[k])) for k in order]
    L.init_emanuel_convection_fortran.restype = None
    L.init_emanuel_convection_fortran(*args)
    qs = OE.bolton_q_sat(st["air_temperature"], st["air_pressure"] * 100, par["rd"], par["rv"])
    ref = OE.convect(par, st["air_temperature"], st["specific_humidity"], qs, st["eastward_wind"], st["northward_wind"],
                     st["air_pressure"], st["air_pressure_on_interface_levels"], st["cloud_base_mass_flux"], dt)
    nlev = st["air_temperature"].shape[1]
    dpp = ctypes.POINTER(ctypes.c_double)
    L.emanuel_convection.restype = None
    cols = [int(c) for c in np.flatnonzero(ref["iflag"] >= 1)[:3]] + [int(np.flatnonzero(ref["iflag"] == 0)[0])]
    for c in cols:
        a = {k: np.ascontiguousarray(v[c]) for k, v in (("t", st["air_temperature"]), ("q", st["specific_humidity"]), ("qs", qs),
                                                         ("u", st["eastward_wind"]), ("v", st["northward_wind"]), ("p", st["air_pressure"]),
                                                         ("ph", st["air_pressure_on_interface_levels"]))}
        o = {k: np.zeros(nlev) for k in ("ft", "fq", "fu", "fv")}
        s = {k: cd(0.0) for k in ("precip", "wd", "tprime", "qprime", "cape")}
        cb = cd(float(st["cloud_base_mass_flux"][c]))
        flag = ci(-1)
        P = lambda x: x.ctypes.data_as(dpp)
        L.emanuel_convection(P(a["t"]), P(a["q"]), P(a["qs"]), P(a["u"]), P(a["v"]), P(a["p"]), P(a["ph"]), ctypes.byref(ci(nlev)),
                             ctypes.byref(ci(nlev - 3)), ctypes.byref(ci(0)), ctypes.byref(cd(dt)), ctypes.byref(flag), P(o["ft"]), P(o["fq"]),
                             P(o["fu"]), P(o["fv"]), ctypes.byref(s["precip"]), ctypes.byref(s["wd"]), ctypes.byref(s["tprime"]),
                             ctypes.byref(s["qprime"]), ctypes.byref(cb), ctypes.byref(s["cape"]), None, None)
        assert flag.value == ref["iflag"][c]
        np.testing.assert_allclose(o["ft"], ref["ft"][c], rtol=0, atol=RTOL * max(np.abs(ref["ft"][c]).max(), 1e-30))
        np.testing.assert_allclose(cb.value, ref["cbmf"][c], rtol=RTOL)
        np.testing.assert_allclose(s["precip"].value, ref["precip"][c], rtol=RTOL, atol=1e-300)
