"""SimplePhysics -- drop-in for climt.SimplePhysics (climt/_components/simple_physics/component.py:12-271): the Reed-Jablonowski
(2012) package -- large-scale condensation, bulk surface fluxes, implicit boundary-layer diffusion -- as a sympl Stepper.
The Fortran (climt/_lib/simple_physics/simple_physics_custom.f90) and the level flip of its Cython shim run in `k_simple_physics`
(csrc/simple_physics.cu) on the component's own (level, column) arrays, surface first.

Host path: numpy in / numpy out through cb200_simple_physics_run_host.  A state of torch CUDA tensors stays on the device
(SURVEY.md 8f-1/8f-4): outputs are CUDA tensors, asynchronous on the current stream.

`boundary_layer=True` with `surface_fluxes=False` makes the reference read uninitialised diffusivities; rejected here.
"""
import ctypes

import numpy as np

from . import _native, device_state
from .constants import get_constant
from .sympl_shim import Stepper

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_IN = ("t", "q", "u", "v", "pmid", "pint", "ps", "ts", "qsurf", "lat")
_OUT = ("t", "q", "u", "v", "precl", "sens_ht_flux", "lat_ht_flux")
_STATE = {"t": "air_temperature", "q": "specific_humidity", "u": "eastward_wind", "v": "northward_wind", "pmid": "air_pressure",
          "pint": "air_pressure_on_interface_levels", "ps": "surface_air_pressure", "ts": "surface_temperature",
          "qsurf": "surface_specific_humidity", "lat": "latitude"}


class Params(ctypes.Structure):
    """cb200_simple_physics_params (include/climt_b200.h)"""
    _fields_ = ([(n, ctypes.c_double) for n in ("gravit", "cpair", "rair", "latvap", "rh2o", "radius", "omega", "rhow", "pbltop",
                                                "pblconst", "C", "Cd0", "Cd1", "Cm")] +
                [(n, ctypes.c_int) for n in ("test", "do_lsc", "do_pbl", "do_surf_flux", "use_ts_ext", "use_qsurf_ext",
                                             "clamp_latent_heat_flux")])


class _InHost(ctypes.Structure):
    _fields_ = [(n, _dp) for n in _IN]


class _OutHost(ctypes.Structure):
    _fields_ = [(n, _dp) for n in _OUT]


class _InDev(ctypes.Structure):
    _fields_ = [(n, _vp) for n in _IN]


class _OutDev(ctypes.Structure):
    _fields_ = [(n, _vp) for n in _OUT]


def _check(L, rc):
    if rc == -3:
        raise ValueError(L.cb200_global_error().decode())
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())


def simple_physics_host(params, arrays, dtime, order=0, device=0):
    """arrays: t q u v pmid (nlev, ncol), pint (nlev+1, ncol), ps ts qsurf lat (ncol) -> dict of the seven outputs"""
    L = _native.lib()
    nlev, ncol = np.shape(arrays["t"])
    keep, s = [], _InHost()
    for k in _IN:
        a = np.ascontiguousarray(arrays[k], dtype=np.float64)
        want = (nlev + 1, ncol) if k == "pint" else ((nlev, ncol) if k in ("t", "q", "u", "v", "pmid") else (ncol,))
        if a.shape != want:
            raise ValueError(f"{k}: expected shape {want}, got {a.shape}")
        keep.append(a)
        setattr(s, k, a.ctypes.data_as(_dp))
    out = {k: np.empty((nlev, ncol) if k in ("t", "q", "u", "v") else (ncol,)) for k in _OUT}
    o = _OutHost()
    for k in _OUT:
        setattr(o, k, out[k].ctypes.data_as(_dp))
    L.cb200_simple_physics_run_host.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.POINTER(Params), ctypes.POINTER(_InHost),
                                                                     ctypes.POINTER(_OutHost)]
    _check(L, L.cb200_simple_physics_run_host(device, ncol, nlev, order, float(dtime), ctypes.byref(params), ctypes.byref(s),
                                              ctypes.byref(o)))
    return out


def simple_physics_device(params, tensors, dtime, order=0, stream=None):
    """torch CUDA tensors (same names and shapes as simple_physics_host); asynchronous on the current stream"""
    import torch
    L = _native.lib()
    t = tensors["t"]
    nlev, ncol = t.shape
    keep, s = [], _InDev()
    for k in _IN:
        a = tensors[k].to(dtype=torch.float64).contiguous()
        keep.append(a)
        setattr(s, k, a.data_ptr())
    out = {k: torch.empty((nlev, ncol) if k in ("t", "q", "u", "v") else (ncol,), dtype=torch.float64, device=t.device) for k in _OUT}
    o = _OutDev()
    for k in _OUT:
        setattr(o, k, out[k].data_ptr())
    work = torch.empty((2, nlev, ncol), dtype=torch.float64, device=t.device)
    L.cb200_simple_physics_run_device.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.POINTER(Params), ctypes.POINTER(_InDev),
                                                                       ctypes.POINTER(_OutDev), _vp, _vp]
    sp = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    _check(L, L.cb200_simple_physics_run_device(t.device.index or 0, ncol, nlev, order, float(dtime), ctypes.byref(params),
                                                ctypes.byref(s), ctypes.byref(o), work.data_ptr(), sp))
    device_state.keep_alive_on(stream, keep + [work] + list(out.values()))
    return out


class SimplePhysics(Stepper):
    """Interface to the simple physics package (Reed and Jablonowski 2012), as climt.SimplePhysics."""

    input_properties = {
        "air_temperature": {"dims": ["mid_levels", "*"], "units": "degK"},
        "air_pressure": {"dims": ["mid_levels", "*"], "units": "Pa"},
        "air_pressure_on_interface_levels": {"dims": ["interface_levels", "*"], "units": "Pa"},
        "surface_air_pressure": {"dims": ["*"], "units": "Pa"},
        "surface_temperature": {"dims": ["*"], "units": "degK"},
        "specific_humidity": {"dims": ["mid_levels", "*"], "units": "kg/kg"},
        "northward_wind": {"dims": ["mid_levels", "*"], "units": "m s^-1"},
        "eastward_wind": {"dims": ["mid_levels", "*"], "units": "m s^-1"},
        "surface_specific_humidity": {"dims": ["*"], "units": "kg/kg"},
        "latitude": {"dims": ["*"], "units": "degrees_north"},
    }
    diagnostic_properties = {
        "stratiform_precipitation_rate": {"dims": ["*"], "units": "m s^-1"},
        "surface_upward_latent_heat_flux": {"dims": ["*"], "units": "W m^-2"},
        "surface_upward_sensible_heat_flux": {"dims": ["*"], "units": "W m^-2"},
    }
    output_properties = {
        "air_temperature": {"units": "degK"},
        "specific_humidity": {"units": "kg/kg"},
        "northward_wind": {"units": "m s^-1"},
        "eastward_wind": {"units": "m s^-1"},
    }

    def __init__(self, simulate_cyclone=False, large_scale_condensation=True, boundary_layer=True, surface_fluxes=True,
                 use_external_surface_temperature=True, use_external_surface_specific_humidity=False,
                 top_of_boundary_layer=85000.0, boundary_layer_influence_height=20000.0, drag_coefficient_heat_fluxes=0.0011,
                 base_momentum_drag_coefficient=0.0007, wind_dependent_momentum_drag_coefficient=0.000065,
                 maximum_momentum_drag_coefficient=0.002, device=0, **kwargs):
        if boundary_layer and not surface_fluxes:
            raise ValueError("SimplePhysics(boundary_layer=True, surface_fluxes=False): the reference's Fortran reads uninitialised "
                             "diffusivities in this combination (simple_physics_custom.f90:368-381, 441-452); not provided")
        self._cyclone, self._lsc, self._pbl, self._surface_flux = simulate_cyclone, large_scale_condensation, boundary_layer, surface_fluxes
        self._use_ext_ts, self._use_ext_qsurf = use_external_surface_temperature, use_external_surface_specific_humidity
        self._Ct, self._pbl_top, self._delta_pbl = drag_coefficient_heat_fluxes, top_of_boundary_layer, boundary_layer_influence_height
        self._Cd0, self._Cd1, self._Cm = base_momentum_drag_coefficient, wind_dependent_momentum_drag_coefficient, maximum_momentum_drag_coefficient
        self._device = device
        _native.lib()
        super().__init__(**kwargs)

    def params(self):
        """the constants are re-read on every call, as the reference does (component.py:240)"""
        p = Params()
        p.gravit = get_constant("gravitational_acceleration", "m/s^2")
        p.cpair = get_constant("heat_capacity_of_dry_air_at_constant_pressure", "J/kg/degK")
        p.rair = get_constant("gas_constant_of_dry_air", "J/kg/degK")
        p.latvap = get_constant("latent_heat_of_condensation", "J/kg")
        p.rh2o = get_constant("gas_constant_of_vapor_phase", "J/kg/degK")
        p.radius = get_constant("planetary_radius", "m")
        p.omega = get_constant("planetary_rotation_rate", "s^-1")
        p.rhow = get_constant("density_of_liquid_water", "kg/m^3")
        p.pbltop, p.pblconst, p.C, p.Cd0, p.Cd1, p.Cm = self._pbl_top, self._delta_pbl, self._Ct, self._Cd0, self._Cd1, self._Cm
        p.test, p.do_lsc, p.do_pbl, p.do_surf_flux = int(self._cyclone), int(self._lsc), int(self._pbl), int(self._surface_flux)
        p.use_ts_ext, p.use_qsurf_ext, p.clamp_latent_heat_flux = int(self._use_ext_ts), int(self._use_ext_qsurf), 1
        return p

    def array_call(self, state, timestep):
        t = state["air_temperature"]
        on_device = type(t).__module__.startswith("torch") and getattr(t, "is_cuda", False)
        nlev = t.shape[0]
        shape_mid, shape_sfc = tuple(t.shape), tuple(state["surface_air_pressure"].shape)
        arrays = {}
        for k, name in _STATE.items():
            a = state[name]
            arrays[k] = a.reshape(nlev + 1, -1) if k == "pint" else (a.reshape(nlev, -1) if k in ("t", "q", "u", "v", "pmid") else a.reshape(-1))
        dt = timestep.total_seconds()
        o = simple_physics_device(self.params(), arrays, dt) if on_device else simple_physics_host(self.params(), arrays, dt, device=self._device)
        new_state = {"eastward_wind": o["u"].reshape(shape_mid), "northward_wind": o["v"].reshape(shape_mid),
                     "air_temperature": o["t"].reshape(shape_mid), "specific_humidity": o["q"].reshape(shape_mid)}
        diagnostics = {"stratiform_precipitation_rate": o["precl"].reshape(shape_sfc),
                       "surface_upward_sensible_heat_flux": o["sens_ht_flux"].reshape(shape_sfc),
                       "surface_upward_latent_heat_flux": o["lat_ht_flux"].reshape(shape_sfc)}
        return diagnostics, new_state
