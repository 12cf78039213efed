"""SimplePhysics (Reed & Jablonowski 2012 as climt drives it).

PARITY ONLY PARTLY PINNED (see oracle/simple_physics.py): the reference's Fortran cannot run here and has no numba port.  Pins:
 * the reference's cached component outputs (TestSimplePhysics-{column,3d}: calm dry default state -- last-bit pattern of the
   boundary layer's potential-temperature round trip included);
 * the reference's own property tests, restated (tests/test_conservation.py:295-313): the column's moist enthalpy changes by
   exactly the surface fluxes times the time step;
 * closed-form checks of each process (saturation adjustment, bulk formulae, a diffusion step conserves column integrals).
GPU: the CUDA kernel through the C ABI and the drop-in component against the oracle on random moist, windy states
(1e-11 relative; fp64, libm differences in exp / pow / log only)."""
import datetime

import numpy as np
import pytest

import helpers as H
from climt_b200 import constants as CN, state as S
from oracle import simple_physics as OS

C = {k: v[0] for k, v in CN.DEFAULTS.items()}
DT = 1200.0


def default_state(nz, ncol=1):
    g = S.default_grid(nz, ncol)
    return dict(g, air_temperature=np.full((nz, ncol), 290.0), specific_humidity=np.zeros((nz, ncol)),
                eastward_wind=np.zeros((nz, ncol)), northward_wind=np.zeros((nz, ncol)), surface_temperature=np.full(ncol, 300.0),
                surface_specific_humidity=np.zeros(ncol), latitude=np.zeros(ncol))


def random_state(nz, ncol, seed, saturated=True):
    rng = np.random.default_rng(seed)
    g = S.default_grid(nz, ncol, p_surf=rng.uniform(9.6e4, 1.03e5, ncol))
    p = g["air_pressure"]
    ts = rng.uniform(271.5, 305.0, ncol)
    ts[:4] = rng.uniform(255.0, 270.5, 4)                       # the cold-surface branch of the saturation formula
    T = np.maximum((ts - rng.uniform(0.0, 3.0, ncol))[None, :] * (p / g["surface_air_pressure"][None, :]) ** 0.19, 200.0)
    es = 610.78 * np.exp(-2.5e6 / 461.5 * (1.0 / T - 1.0 / 273.16))
    qsat = 287.0 / 461.5 * es / p
    rh = rng.uniform(0.2, 1.15 if saturated else 0.95, (nz, ncol))  # > 1: large-scale condensation fires
    wind = rng.uniform(0.0, 35.0, ncol)                         # beyond 20 m/s: the capped drag coefficient
    wind[4] = 0.0
    ang = rng.uniform(0, 2 * np.pi, ncol)
    shear = (p / g["surface_air_pressure"][None, :]) ** 0.3
    return dict(g, air_temperature=T, specific_humidity=qsat * rh, eastward_wind=(wind * np.cos(ang))[None, :] * shear,
                northward_wind=(wind * np.sin(ang))[None, :] * shear, surface_temperature=ts,
                surface_specific_humidity=rng.uniform(0.0, 0.02, ncol), latitude=rng.uniform(-80.0, 80.0, ncol))


def moist_enthalpy(st):
    dp = st["air_pressure_on_interface_levels"][:-1] - st["air_pressure_on_interface_levels"][1:]
    return np.sum((C["heat_capacity_of_dry_air_at_constant_pressure"] * st["air_temperature"]
                   + C["latent_heat_of_condensation"] * st["specific_humidity"]) * dp / C["gravitational_acceleration"], axis=0)


@pytest.mark.parametrize("kind,nz", [("column", 30), ("3d", 28)])
def test_oracle_matches_the_reference_caches(kind, nz):
    g = H.golden()
    st = default_state(nz)
    diag, new = OS.component_call(st, DT, C)
    cls = f"TestSimplePhysics-{kind}"
    for name in ("stratiform_precipitation_rate", "surface_upward_latent_heat_flux", "surface_upward_sensible_heat_flux"):
        np.testing.assert_array_equal(diag[name][0], g[f"{cls}/tend/{name}"].reshape(-1)[0])
    for name in ("specific_humidity", "northward_wind", "eastward_wind"):
        np.testing.assert_array_equal(new[name][:, 0], g[f"{cls}/diag/{name}"].reshape(nz, -1)[:, 0])
    ref_T = g[f"{cls}/diag/air_temperature"].reshape(nz, -1)[:, 0]
    np.testing.assert_allclose(new["air_temperature"][:, 0], ref_T, rtol=3e-16, atol=0)
    assert (ref_T != 290.0).any()      # the cache does carry the theta round trip's last-bit pattern ...
    assert np.mean(new["air_temperature"][:, 0] == ref_T) > 0.8   # ... and the restatement reproduces most of it bit for bit


@pytest.mark.parametrize("options", [dict(boundary_layer=False, use_external_surface_specific_humidity=True), dict(boundary_layer=False)])
def test_moist_enthalpy_changes_by_the_surface_fluxes(options):
    """tests/test_conservation.py:295-313 of the reference: u = 3 m/s, one 1 s step, |d(enthalpy) - fluxes * dt| <= 1e-3"""
    st = default_state(30)
    st["eastward_wind"][:] = 3.0
    diag, new = OS.component_call(st, 1.0, C, **options)
    forcing = (diag["surface_upward_sensible_heat_flux"] + diag["surface_upward_latent_heat_flux"]) * 1.0
    assert abs(forcing[0]) > 1.0
    assert abs((moist_enthalpy(dict(st, **new)) - moist_enthalpy(st) - forcing)[0]) <= 1e-3


def test_each_process_in_closed_form():
    st = random_state(40, 64, 5)
    K = OS.constants(OS.DEFAULTS, C)
    # condensation alone: supersaturated layers relax towards saturation, conserve cp T + L q layer by layer, rain = column loss
    diag, new = OS.component_call(st, DT, C, boundary_layer=False, surface_fluxes=False)
    dq, dT = new["specific_humidity"] - st["specific_humidity"], new["air_temperature"] - st["air_temperature"]
    assert (dq < 0).sum() > 50 and np.all(dq <= 0)
    np.testing.assert_allclose(K["cpair"] * dT, -K["latvap"] * dq, rtol=1e-9, atol=1e-9)
    dp = st["air_pressure_on_interface_levels"][:-1] - st["air_pressure_on_interface_levels"][1:]
    np.testing.assert_allclose(diag["stratiform_precipitation_rate"], -(dq * dp).sum(0) / (K["gravit"] * K["rhow"] * DT), rtol=1e-10, atol=1e-18)
    # surface fluxes alone: bulk formulae on the lowest layer
    diag, new = OS.component_call(st, DT, C, large_scale_condensation=False, boundary_layer=False, use_external_surface_specific_humidity=True)
    wind = np.hypot(st["eastward_wind"][0], st["northward_wind"][0])
    rho = st["air_pressure"][0] / (K["rair"] * st["air_temperature"][0])
    np.testing.assert_allclose(diag["surface_upward_sensible_heat_flux"],
                               rho * K["cpair"] * K["C"] * wind * (st["surface_temperature"] - st["air_temperature"][0]), rtol=1e-13)
    assert diag["surface_upward_sensible_heat_flux"][4] == 0.0 and np.all(diag["surface_upward_latent_heat_flux"] >= 0)
    np.testing.assert_array_equal(new["air_temperature"][1:], st["air_temperature"][1:])
    # the boundary layer conserves the column integrals of u, v, q (implicit diffusion with closed ends)
    calm = dict(st)
    base_d, base = OS.component_call(calm, DT, C, large_scale_condensation=False, boundary_layer=False)
    full_d, full = OS.component_call(calm, DT, C, large_scale_condensation=False)
    for name in ("eastward_wind", "northward_wind", "specific_humidity"):
        np.testing.assert_allclose((full[name] * dp).sum(0), (base[name] * dp).sum(0), rtol=1e-11, atol=1e-9)
    assert np.abs(full["eastward_wind"] - base["eastward_wind"]).max() > 0.1


def test_unprovided_combination_is_rejected():
    from climt_b200.simple_physics import SimplePhysics
    with pytest.raises(ValueError, match="uninitialised"):
        SimplePhysics(boundary_layer=True, surface_fluxes=False)


OPTION_SETS = [dict(), dict(boundary_layer=False), dict(large_scale_condensation=False, use_external_surface_specific_humidity=True),
               dict(use_external_surface_temperature=False), dict(use_external_surface_temperature=False, simulate_cyclone=True),
               dict(boundary_layer=False, surface_fluxes=False), dict(top_of_boundary_layer=70000.0, boundary_layer_influence_height=1e4,
                                                                      drag_coefficient_heat_fluxes=0.002)]


def _params(options):
    """cb200_simple_physics_params as the drop-in fills it (no engine is created: the constructor only checks the library exists)"""
    from climt_b200.simple_physics import SimplePhysics
    return SimplePhysics(**options).params()


@pytest.mark.parametrize("options", OPTION_SETS)
def test_kernel_code_on_the_host_matches_the_oracle(options):
    """the per-column code of k_simple_physics (csrc/simple_physics_core.cuh) compiled for the host, NaN-poisoned workspace"""
    st = random_state(45, 97, 21)
    diag_ref, new_ref = OS.component_call(st, DT, C, **options)
    out = H.run_simple_physics_emul(st, DT, _params(options))
    for k, name in (("t", "air_temperature"), ("q", "specific_humidity"), ("u", "eastward_wind"), ("v", "northward_wind")):
        _assert_close(out[k], new_ref[name], name)
    for k, name in (("precl", "stratiform_precipitation_rate"), ("sens_ht_flux", "surface_upward_sensible_heat_flux"),
                    ("lat_ht_flux", "surface_upward_latent_heat_flux")):
        _assert_close(out[k], diag_ref[name], name)
    # the Fortran's own level order (model top first), as the reference-named symbol passes it
    flipped = {k: (v[::-1].copy() if v.ndim == 2 else v) for k, v in st.items()}
    out1 = H.run_simple_physics_emul(flipped, DT, _params(options), order=1)
    for k in ("t", "q", "u", "v"):
        np.testing.assert_array_equal(out1[k][::-1], out[k])


def test_kernel_code_reproduces_the_reference_cache():
    g = H.golden()
    out = H.run_simple_physics_emul(default_state(30), DT, _params({}))
    np.testing.assert_allclose(out["t"][:, 0], g["TestSimplePhysics-column/diag/air_temperature"][:, 0, 0], rtol=3e-16, atol=0)
    assert out["precl"][0] == 0.0 and out["sens_ht_flux"][0] == 0.0 and out["lat_ht_flux"][0] == 0.0


# ------------------------------------------------------------------------------------------------ GPU
def _assert_close(got, ref, name):
    scale = float(np.abs(ref).max()) or 1.0
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13 * scale, err_msg=name)


@pytest.mark.gpu
@pytest.mark.parametrize("options", OPTION_SETS)
def test_drop_in_matches_the_oracle(options):
    from climt_b200.simple_physics import SimplePhysics
    st = random_state(60, 777, 11)
    diag_ref, new_ref = OS.component_call(st, DT, C, **options)
    diag, new = SimplePhysics(**options).array_call(dict(st), datetime.timedelta(seconds=DT))
    for k in diag_ref:
        _assert_close(diag[k], diag_ref[k], k)
    for k in new_ref:
        _assert_close(new[k], new_ref[k], k)
    if not options:
        assert (diag_ref["stratiform_precipitation_rate"] > 0).sum() > 100 and (diag_ref["surface_upward_latent_heat_flux"] == 0).sum() > 0


@pytest.mark.gpu
def test_reference_caches_three_dimensional_state_and_device_path():
    import torch
    from climt_b200.simple_physics import SimplePhysics
    g = H.golden()
    comp = SimplePhysics()
    st = default_state(28, 16 * 32)
    st3 = {k: (v.reshape(v.shape[0], 16, 32) if v.ndim == 2 else v.reshape(16, 32)) for k, v in st.items()}
    diag, new = comp.array_call(st3, datetime.timedelta(seconds=DT))
    np.testing.assert_allclose(new["air_temperature"], g["TestSimplePhysics-3d/diag/air_temperature"], rtol=1e-15)
    assert new["air_temperature"].shape == (28, 16, 32) and diag["surface_upward_latent_heat_flux"].shape == (16, 32)
    for k in ("stratiform_precipitation_rate", "surface_upward_latent_heat_flux", "surface_upward_sensible_heat_flux"):
        np.testing.assert_array_equal(diag[k], g[f"TestSimplePhysics-3d/tend/{k}"])
    # device-resident state: same bits as the host path
    rs = random_state(60, 4097, 12)
    d_h, n_h = comp.array_call(dict(rs), datetime.timedelta(seconds=DT))
    d_d, n_d = comp.array_call({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in rs.items()}, datetime.timedelta(seconds=DT))
    torch.cuda.synchronize()
    for k in n_h:
        assert n_d[k].is_cuda
        np.testing.assert_array_equal(n_d[k].cpu().numpy(), n_h[k])
    for k in d_h:
        np.testing.assert_array_equal(d_d[k].cpu().numpy(), d_h[k])
    # the reference's conservation property on the GPU result
    st1 = default_state(30, 256)
    st1["eastward_wind"][:] = 3.0
    d1, n1 = SimplePhysics(boundary_layer=False).array_call(dict(st1), datetime.timedelta(seconds=1))
    forcing = d1["surface_upward_sensible_heat_flux"] + d1["surface_upward_latent_heat_flux"]
    assert np.abs(moist_enthalpy(dict(st1, **n1)) - moist_enthalpy(st1) - forcing).max() <= 1e-3


@pytest.mark.gpu
def test_reference_named_symbols_top_down_in_place():
    """set_fortran_constants + simple_physics exactly as _simple_physics.pyx:36-180 calls them"""
    import ctypes
    from climt_b200 import _native
    L = _native.lib()
    st = random_state(30, 300, 13)
    ref_d, ref_n = OS.component_call(st, DT, C)
    K = OS.constants(OS.DEFAULTS, C)
    dbl = lambda x: ctypes.byref(ctypes.c_double(x))  # noqa: E731
    L.set_fortran_constants(*[dbl(K[k]) for k in ("gravit", "cpair", "rair", "latvap", "rh2o", "radius", "omega", "rhow", "pbltop",
                                                  "pblconst", "C", "Cd0", "Cd1", "Cm")])
    flip = lambda a: np.ascontiguousarray(a[::-1])  # noqa: E731
    t, q, u, v = (flip(st[k]) for k in ("air_temperature", "specific_humidity", "eastward_wind", "northward_wind"))
    pmid, pint = flip(st["air_pressure"]), flip(st["air_pressure_on_interface_levels"])
    pdel = pint[1:] - pint[:-1]
    rpdel = 1.0 / pdel
    n = 300
    precl, sens, lath = np.zeros(n), np.zeros(n), np.zeros(n)
    dp = ctypes.POINTER(ctypes.c_double)
    ptr = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    ints = [ctypes.c_int(x) for x in (n, 30, 0, 1, 1, 1, 1, 0)]  # pcols pver test lsc pbl surf ext_ts ext_qsurf
    L.simple_physics(ctypes.byref(ints[0]), ctypes.byref(ints[1]), dbl(DT), ptr(st["latitude"]), ptr(t), ptr(q), ptr(u), ptr(v), ptr(pmid),
                     ptr(pint), ptr(pdel), ptr(rpdel), ptr(st["surface_air_pressure"]), ptr(precl), ctypes.byref(ints[2]),
                     ctypes.byref(ints[3]), ctypes.byref(ints[4]), ctypes.byref(ints[5]), ctypes.byref(ints[6]),
                     ptr(st["surface_temperature"]), ctypes.byref(ints[7]), ptr(st["surface_specific_humidity"]), ptr(sens), ptr(lath))
    _assert_close(t[::-1], ref_n["air_temperature"], "t")
    _assert_close(q[::-1], ref_n["specific_humidity"], "q")
    _assert_close(u[::-1], ref_n["eastward_wind"], "u")
    _assert_close(precl, ref_d["stratiform_precipitation_rate"], "precl")
    _assert_close(np.where(lath < 0, 0, lath), ref_d["surface_upward_latent_heat_flux"], "lath")


def test_boundary_layer_sweeps_solve_the_tridiagonal_system_they_claim():
    """The Fortran's forward / backward sweeps (simple_physics_custom.f90:470-520) are a Thomas solve of
    -CA(k) x(k+1) + (1 + CA(k) + CC(k)) x(k) - CC(k) x(k-1) = x_old(k).  Here the same system is assembled as a dense matrix from
    the state after the surface-flux step and handed to numpy.linalg.solve: an independent check of the sweep algebra (indices,
    elimination factors, the theta <-> T conversion) in both the oracle and the kernel code."""
    st = random_state(30, 16, 33, saturated=False)
    opts = dict(large_scale_condensation=False)
    K = OS.constants(OS.DEFAULTS, C)
    _, base = OS.component_call(st, DT, C, boundary_layer=False, **opts)       # state at the entry of the boundary-layer step
    _, full = OS.component_call(st, DT, C, **opts)
    emul = H.run_simple_physics_emul(st, DT, _params(opts))
    flip = lambda a: np.asarray(a)[::-1]                                        # noqa: E731  top-down, as the Fortran indexes
    pmid, pint = flip(st["air_pressure"]), flip(st["air_pressure_on_interface_levels"])
    nlev, ncol = pmid.shape
    t0 = flip(st["air_temperature"])
    q0 = flip(st["specific_humidity"])
    zvir = K["rh2o"] / K["rair"] - 1.0
    za = K["rair"] / K["gravit"] * t0[-1] * (1.0 + zvir * q0[-1]) * 0.5 * (np.log(st["surface_air_pressure"]) - np.log(pint[nlev - 1]))
    wind = np.hypot(st["eastward_wind"][0], st["northward_wind"][0])
    Ke_s = K["C"] * wind * za
    Km_s = np.where(wind < 20.0, (K["Cd0"] + K["Cd1"] * wind) * wind * za, K["Cm"] * wind * za)
    taper = np.where(pint >= K["pbltop"], 1.0, np.exp(-(K["pbltop"] - pint) ** 2 / K["pblconst"] ** 2))
    tb = flip(base["air_temperature"])
    kap = K["rair"] / K["cpair"]
    for name, Ksfc, to_x, from_x in (("specific_humidity", Ke_s, None, None), ("eastward_wind", Km_s, None, None),
                                     ("air_temperature", Ke_s, (100000.0 / pmid) ** kap, (pmid / 100000.0) ** kap)):
        x_old = flip(base[name]) * (to_x if to_x is not None else 1.0)
        want = np.zeros_like(x_old)
        for c in range(ncol):
            A = np.zeros((nlev, nlev))
            for k in range(nlev):
                ca = cc = 0.0
                if k < nlev - 1:
                    rho = pint[k + 1, c] / (K["rair"] * (tb[k + 1, c] + tb[k, c]) / 2.0)
                    ca = DT * K["gravit"] ** 2 * Ksfc[c] * taper[k + 1, c] * rho ** 2 / ((pmid[k + 1, c] - pmid[k, c]) * (pint[k + 1, c] - pint[k, c]))
                    A[k, k + 1] = -ca
                if k > 0:
                    rho = pint[k, c] / (K["rair"] * (tb[k, c] + tb[k - 1, c]) / 2.0)
                    cc = DT * K["gravit"] ** 2 * Ksfc[c] * taper[k, c] * rho ** 2 / ((pmid[k, c] - pmid[k - 1, c]) * (pint[k + 1, c] - pint[k, c]))
                    A[k, k - 1] = -cc
                A[k, k] = 1.0 + ca + cc
            want[:, c] = np.linalg.solve(A, x_old[:, c])
        if from_x is not None:
            want = want * from_x
        scale = float(np.abs(want).max())
        np.testing.assert_allclose(flip(full[name]), want, rtol=1e-10, atol=1e-12 * scale, err_msg=name)
        key = {"specific_humidity": "q", "eastward_wind": "u", "air_temperature": "t"}[name]
        np.testing.assert_allclose(emul[key][::-1], want, rtol=1e-10, atol=1e-12 * scale, err_msg=name + " (kernel code)")
    assert np.abs(full["eastward_wind"] - base["eastward_wind"]).max() > 0.05     # the diffusion did something
