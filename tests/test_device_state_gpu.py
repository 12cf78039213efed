"""Device-resident component path: a state of torch CUDA tensors goes through array_call without touching the
host -- marshal kernel + engines on device pointers -- and must agree with the host-buffer path of the same component."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# the marshal kernel evaluates log() / cos() with the CUDA math library (<= 2 ulp from numpy's), everything downstream is the
# same kernel chain as the host path: fluxes agree to ~1e-13 relative; the tolerance below is the stated bound.
RTOL, ATOL = 1e-11, 1e-11


def _raw_lw(nz, ncol):
    from climt_b200 import state as S
    st = dict(S.default_rrtmg_lw_state(nz, ncol))
    st["air_pressure"] = st["air_pressure"] / 100.0
    st["air_pressure_on_interface_levels"] = st["air_pressure_on_interface_levels"] / 100.0
    st["mass_content_of_cloud_ice_in_atmosphere_layer"] = st["mass_content_of_cloud_ice_in_atmosphere_layer"] * 1e3
    st["mass_content_of_cloud_liquid_water_in_atmosphere_layer"] = st["mass_content_of_cloud_liquid_water_in_atmosphere_layer"] * 1e3
    return st


def _raw_sw(nz, ncol):
    from climt_b200 import state as S
    st = dict(S.default_rrtmg_sw_state(nz, ncol))
    st["air_pressure"] = st["air_pressure"] / 100.0
    st["air_pressure_on_interface_levels"] = st["air_pressure_on_interface_levels"] / 100.0
    st["mass_content_of_cloud_ice_in_atmosphere_layer"] = st["mass_content_of_cloud_ice_in_atmosphere_layer"] * 1e3
    st["mass_content_of_cloud_liquid_water_in_atmosphere_layer"] = st["mass_content_of_cloud_liquid_water_in_atmosphere_layer"] * 1e3
    return st


def _perturb(st, seed):
    """make the columns differ (the default state repeats one column)"""
    rng = np.random.default_rng(seed)
    ncol = st["air_temperature"].shape[1]
    st["air_temperature"] = st["air_temperature"] + rng.uniform(-5, 5, st["air_temperature"].shape)
    st["surface_temperature"] = st["surface_temperature"] + rng.uniform(-3, 3, ncol)
    st["specific_humidity"] = st["specific_humidity"] * rng.uniform(0.5, 1.5, st["specific_humidity"].shape)
    if "zenith_angle" in st:
        st["zenith_angle"] = np.deg2rad(rng.uniform(0.0, 100.0, ncol))   # some columns below the horizon
    return st


def _to_device(st):
    import torch
    return {k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if isinstance(v, np.ndarray) and v.ndim > 0 else v) for k, v in st.items()}


def _compare(dev, host):
    import torch
    for name, ref in host.items():
        got = dev[name]
        assert isinstance(got, torch.Tensor) and got.is_cuda and got.dtype == torch.float64, name
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=ATOL, err_msg=name)


@pytest.mark.parametrize("ncol", [1, 700])
def test_longwave_component_on_device_state(ncol):
    from climt_b200.rrtmg_lw import RRTMGLongwave
    st = _perturb(_raw_lw(30, ncol), 5)
    comp = RRTMGLongwave()
    tend_h, diag_h = comp.array_call(st)
    tend_d, diag_d = comp.array_call(_to_device(st))
    _compare(tend_d, tend_h)
    _compare(diag_d, diag_h)


def test_shortwave_component_on_device_state():
    from climt_b200.rrtmg_sw import RRTMGShortwave
    st = _perturb(_raw_sw(30, 700), 6)
    comp = RRTMGShortwave(ignore_day_of_year=True)
    tend_h, diag_h = comp.array_call(st)
    tend_d, diag_d = comp.array_call(_to_device(st))
    _compare(tend_d, tend_h)
    _compare(diag_d, diag_h)
    assert float(diag_d["downwelling_shortwave_flux_in_air"].max()) > 100.0


def test_marshal_kernel_matches_numpy():
    """cb200_marshal_device against the reference's host arithmetic (climt/_core/util.py: mass_to_volume_mixing_ratio,
    get_interface_values; sw/component.py np.cos(zenith))."""
    import torch
    from climt_b200 import device_state, state as S
    st = _perturb(_raw_sw(40, 333), 9)
    d = _to_device(st)
    q, tlev, cz = device_state.marshal(d["specific_humidity"], d["air_temperature"], d["surface_temperature"], d["air_pressure"],
                                       d["air_pressure_on_interface_levels"], zenith=d["zenith_angle"])
    torch.cuda.synchronize()
    np.testing.assert_allclose(q.cpu().numpy(), S.mass_to_volume_mixing_ratio(st["specific_humidity"], 18.02), rtol=1e-15)
    ref = S.get_interface_values(st["air_temperature"], st["surface_temperature"], st["air_pressure"], st["air_pressure_on_interface_levels"])
    np.testing.assert_allclose(tlev.cpu().numpy(), ref, rtol=1e-13)
    np.testing.assert_allclose(cz.cpu().numpy(), np.cos(st["zenith_angle"]), rtol=0, atol=2e-16)


def _bad_cloud_state():
    st = _raw_lw(30, 64)
    for k in ("cloud_area_fraction_in_atmosphere_layer", "mass_content_of_cloud_ice_in_atmosphere_layer", "cloud_ice_particle_size"):
        st[k] = st[k].copy()
    st["cloud_area_fraction_in_atmosphere_layer"][10, :] = 0.5
    st["mass_content_of_cloud_ice_in_atmosphere_layer"][10, :] = 10.0
    st["cloud_ice_particle_size"][10, :] = 500.0
    return st


def test_synchronous_mode_raises_like_the_host_path():
    from climt_b200.rrtmg_lw import RRTMGLongwave
    st = _bad_cloud_state()
    comp = RRTMGLongwave()
    with pytest.raises(ValueError, match="ICE RADIUS OUT OF BOUNDS"):
        comp.array_call(st)
    with pytest.raises(ValueError, match="ICE RADIUS OUT OF BOUNDS"):
        comp.array_call(_to_device(st))


def test_asynchronous_mode_defers_validation():
    """asynchronous=True returns without synchronising; the Fortran `stop` conditions are reported by engine.check() once the
    caller has synchronised."""
    import torch
    from climt_b200.rrtmg_lw import RRTMGLongwave
    comp = RRTMGLongwave(asynchronous=True)
    comp.array_call(_to_device(_bad_cloud_state()))      # no exception here
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="ICE RADIUS OUT OF BOUNDS"):
        comp._engine.check()
    good = _perturb(_raw_lw(30, 64), 2)
    tend, diag = comp.array_call(_to_device(good))
    torch.cuda.synchronize()
    comp._engine.check()
    ref_t, _ = RRTMGLongwave().array_call(good)
    np.testing.assert_allclose(tend["air_temperature"].cpu().numpy(), ref_t["air_temperature"], rtol=RTOL, atol=ATOL)
