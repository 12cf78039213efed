"""Instellation -- drop-in for climt.Instellation (climt/_components/instellation/component.py:9-81): zenith angle of the
sun from latitude, longitude and the model time.  The per-column arithmetic of `_instellation_kernel_np` (:84-132) runs in
`k_instellation` (csrc/adjacent_engine.cu); the per-call orbital scalars (:135-191) in `cb200_instellation_orbit`.

A state of torch CUDA tensors stays on the device: `zenith_angle` comes back as a CUDA tensor, and
`instellation_device(..., want_coszen=True)` additionally returns the cosine the shortwave engine reads.
"""
import ctypes
import datetime

import numpy as np

from . import _native, device_state
from .sympl_shim import DiagnosticComponent

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p


def days_from_2000(model_time):
    """days since 2000-01-01 12:00 (component.py:64-76)"""
    d = model_time - datetime.datetime(2000, 1, 1, 12, 0)
    return d.days + (d.seconds + d.microseconds / 1000000.0) / (24 * 3600.0)


def julian_centuries(model_time):
    return days_from_2000(model_time) / 36525.0


def orbit(jc):
    """(sin_dec, cos_dec, right_ascension, gmst) of the call -- host arithmetic of the C ABI, no GPU involved."""
    L = _native.lib()
    L.cb200_instellation_orbit.argtypes = [ctypes.c_double] + [_dp] * 4
    L.cb200_instellation_orbit.restype = None
    out = [ctypes.c_double() for _ in range(4)]
    L.cb200_instellation_orbit(float(jc), *[ctypes.byref(x) for x in out])
    return tuple(x.value for x in out)


def instellation_host(lat_deg, lon_deg, jc, device=0):
    """numpy (ncol,) in / numpy (ncol,) out through cb200_instellation_run_host"""
    L = _native.lib()
    lat = np.ascontiguousarray(lat_deg, dtype=np.float64).reshape(-1)
    lon = np.ascontiguousarray(lon_deg, dtype=np.float64).reshape(-1)
    if lat.size != lon.size:
        raise ValueError("latitude and longitude differ in size")
    zen = np.empty(lat.size)
    L.cb200_instellation_run_host.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_double, _dp]
    rc = L.cb200_instellation_run_host(device, lat.size, lat.ctypes.data_as(_dp), lon.ctypes.data_as(_dp), float(jc),
                                       zen.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())
    return zen


def instellation_device(lat_deg, lon_deg, jc, want_coszen=False, stream=None):
    """torch CUDA tensors in / out, asynchronous on the current stream -> zenith [, coszen]"""
    import torch
    L = _native.lib()
    lat = lat_deg.to(dtype=torch.float64).contiguous().reshape(-1)
    lon = lon_deg.to(dtype=torch.float64).contiguous().reshape(-1)
    zen = torch.empty_like(lat)
    cz = torch.empty_like(lat) if want_coszen else None
    L.cb200_instellation_run_device.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_double, _vp, _vp, _vp]
    s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    rc = L.cb200_instellation_run_device(lat.device.index or 0, lat.numel(), lat.data_ptr(), lon.data_ptr(), float(jc),
                                         zen.data_ptr(), cz.data_ptr() if cz is not None else None, s)
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())
    device_state.keep_alive_on(stream, [lat, lon, zen, cz])
    return (zen, cz) if want_coszen else zen


class Instellation(DiagnosticComponent):
    """Zenith angle given latitude, longitude and time (Earth-sun system), as climt.Instellation."""

    input_properties = {
        "latitude": {"dims": ["*"], "units": "degrees_north"},
        "longitude": {"dims": ["*"], "units": "degrees_east"},
    }
    diagnostic_properties = {"zenith_angle": {"dims": ["*"], "units": "radians"}}

    def __init__(self, device=0, **kwargs):
        self._device = device
        _native.lib()
        super().__init__(**kwargs)

    def array_call(self, state):
        lat, lon = state["latitude"], state["longitude"]
        jc = julian_centuries(state["time"])
        if type(lat).__module__.startswith("torch") and getattr(lat, "is_cuda", False):
            return {"zenith_angle": instellation_device(lat, lon, jc).reshape(lat.shape)}
        lat = np.asarray(lat)
        return {"zenith_angle": instellation_host(lat, lon, jc, self._device).reshape(lat.shape)}
