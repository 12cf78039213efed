"""GrayLongwaveRadiation -- drop-in for climt.GrayLongwaveRadiation (climt/_components/radiation.py:18-109)."""
import ctypes

import numpy as np

from . import _native
from .constants import get_constant
from .sympl_shim import TendencyComponent

_dp = ctypes.POINTER(ctypes.c_double)


def gray_lw_host(t, p_int, t_surf, tau, sigma, g, cpd, device=0):
    """numpy in / numpy out through cb200_gray_lw_run_host."""
    L = _native.lib()
    nlay, ncol = t.shape
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (t, p_int, t_surf, tau)]
    assert a[1].shape == (nlay + 1, ncol) and a[3].shape == (nlay + 1, ncol) and a[2].shape == (ncol,)
    down, up, tend = np.empty((nlay + 1, ncol)), np.empty((nlay + 1, ncol)), np.empty((nlay, ncol))
    L.cb200_gray_lw_run_host.argtypes = [ctypes.c_int] * 3 + [_dp] * 4 + [ctypes.c_double] * 3 + [_dp] * 3
    rc = L.cb200_gray_lw_run_host(device, ncol, nlay, *[x.ctypes.data_as(_dp) for x in a], sigma, g, cpd,
                                  down.ctypes.data_as(_dp), up.ctypes.data_as(_dp), tend.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())
    return down, up, tend


def gray_lw_device(t, p_int, t_surf, tau, sigma, g, cpd, down, up, tend, stream=None):
    """torch CUDA tensors, asynchronous on the current stream."""
    import torch
    L = _native.lib()
    nlay, ncol = t.shape
    vp = ctypes.c_void_p
    L.cb200_gray_lw_run_device.argtypes = [ctypes.c_int] * 3 + [vp] * 4 + [ctypes.c_double] * 3 + [vp] * 4
    s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    rc = L.cb200_gray_lw_run_device(t.device.index or 0, ncol, nlay, t.data_ptr(), p_int.data_ptr(), t_surf.data_ptr(),
                                    tau.data_ptr(), sigma, g, cpd, down.data_ptr(), up.data_ptr(), tend.data_ptr(), s)
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())


class GrayLongwaveRadiation(TendencyComponent):
    input_properties = {
        "longwave_optical_depth_on_interface_levels": {"dims": ["interface_levels", "*"], "units": "dimensionless", "alias": "tau"},
        "air_temperature": {"dims": ["mid_levels", "*"], "units": "degK", "alias": "sl"},
        "surface_temperature": {"dims": ["*"], "units": "degK", "alias": "T_surface"},
        "air_pressure": {"dims": ["mid_levels", "*"], "units": "Pa", "alias": "p"},
        "air_pressure_on_interface_levels": {"dims": ["interface_levels", "*"], "units": "Pa", "alias": "p_interface"},
    }
    diagnostic_properties = {
        "downwelling_longwave_flux_in_air": {"dims": ["interface_levels", "*"], "units": "W m^-2", "alias": "lw_down"},
        "upwelling_longwave_flux_in_air": {"dims": ["interface_levels", "*"], "units": "W m^-2", "alias": "lw_up"},
        "air_temperature_tendency_from_longwave": {"dims": ["mid_levels", "*"], "units": "degK day^-1"},
    }
    tendency_properties = {"air_temperature": {"units": "degK s^-1"}}

    def __init__(self, device=0, **kwargs):
        self._device = device
        _native.lib()
        super().__init__(**kwargs)

    def array_call(self, state):
        def pick(alias, name):
            return state[alias] if alias in state else state[name]
        t = np.asarray(pick("sl", "air_temperature"))
        tau = np.asarray(pick("tau", "longwave_optical_depth_on_interface_levels"))
        t_surf = np.asarray(pick("T_surface", "surface_temperature"))
        p_int = np.asarray(pick("p_interface", "air_pressure_on_interface_levels"))
        orig_t, orig_p = t.shape, p_int.shape
        t2, tau2 = t.reshape(t.shape[0], -1), tau.reshape(tau.shape[0], -1)
        ts2, p2 = t_surf.reshape(-1), p_int.reshape(p_int.shape[0], -1)
        down, up, tend = gray_lw_host(t2, p2, ts2, tau2, get_constant("stefan_boltzmann_constant"),
                                      get_constant("gravitational_acceleration"),
                                      get_constant("heat_capacity_of_dry_air_at_constant_pressure"), self._device)
        tend = tend.reshape(orig_t)
        return {"sl": tend}, {"lw_down": down.reshape(orig_p), "lw_up": up.reshape(orig_p),
                              "air_temperature_tendency_from_longwave": tend * 86400.0}
