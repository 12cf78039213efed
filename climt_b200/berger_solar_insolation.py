"""BergerSolarInsolation -- drop-in for climt.BergerSolarInsolation (climt/_components/berger_solar_insolation.py:495-693): solar
insolation and zenith angle from Berger's (1978) spectral solutions for the orbital parameters (the CAM 3 approach).

As in the reference, the four orbital parameters of a year are evaluated on the host in numpy, once per year
(`orbital_parameters`, reference :579-632; coefficients in climt_b200/data/berger1978.npz, extracted by
tools/extract_berger_tables.py); the per-column part -- the reference's numba kernel `_get_solar_parameters_np` (:635-680) -- runs
in `k_berger` (csrc/adjacent_engine.cu).  A state of torch CUDA tensors stays on the device.
"""
import ctypes
import os

import numpy as np

from . import _native
from .constants import get_constant
from .sympl_shim import DiagnosticComponent

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_TABLES = None


def _tables():
    global _TABLES
    if _TABLES is None:
        with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "berger1978.npz")) as z:
            _TABLES = {k: z[k] for k in z.files}
    return _TABLES


def orbital_parameters(years_since_jan_1_1950):
    """-> (lambda_m0, eccentricity, omega_tilde, obliquity): Berger (1978) equations 1-6 as the reference sums them (:579-625)"""
    T = _tables()
    t, a2d = years_since_jan_1_1950, float(T["arcsec_to_degree"])
    obliquity = 23.320556
    obliquity += np.sum(T["A"] * a2d * np.cos((T["f"] * a2d * t + T["delta"]) * np.pi / 180.0))
    obliquity = obliquity * np.pi / 180.0
    cos_sum = np.sum(T["P"] * np.cos(T["alpha"] * a2d * t + T["zeta"]))    # zeta is in radians (CAM 3 shr_orb_mod)
    sin_sum = np.sum(T["P"] * np.sin(T["alpha"] * a2d * t + T["zeta"]))
    e2 = cos_sum * cos_sum + sin_sum * sin_sum
    e = np.sqrt(e2)
    e3 = e * e2
    pi_val = np.arctan2(sin_sum, cos_sum)
    if pi_val < 0:
        pi_val += 2.0 * np.pi
    omega_tilde = pi_val * 180.0 / np.pi + 50.439273 * a2d * t + 3.392506
    omega_tilde += np.sum(T["F"] * np.sin((T["f_prime"] * a2d * t + T["delta_prime"]) * np.pi / 180.0))
    omega_tilde = omega_tilde % 360.0
    omega_tilde = omega_tilde * np.pi / 180.0
    beta = np.sqrt(1.0 - e2)
    lambda_m0 = 2.0 * ((0.5 * e + 0.125 * e3) * (1.0 + beta) * np.sin(omega_tilde + np.pi)
                       - 0.25 * e2 * (0.5 + beta) * np.sin(2 * (omega_tilde + np.pi))
                       + 0.125 * e3 * (1.0 / 3.0 + beta) * np.sin(3 * (omega_tilde + np.pi)))
    return lambda_m0, e, omega_tilde, obliquity


def years_since_vernal_equinox(dt):
    """fraction of the year since March 20, noon UTC (:683-688)"""
    year_start, year_end = type(dt)(dt.year, 3, 20, 12), type(dt)(dt.year + 1, 3, 20, 12)
    return (dt - year_start).total_seconds() / (year_end - year_start).total_seconds()


def fractional_day(dt):
    return (dt - type(dt)(dt.year, dt.month, dt.day)).total_seconds() / (24.0 * 60.0 * 60.0)


def solar_parameters(orbit, ysve, fday, lat, lon, solar_constant, device=0):
    """lat / lon: numpy arrays or torch CUDA tensors (ncol,) -> insolation, zenith (same kind), rho"""
    L = _native.lib()
    lambda_m0, ecc, omega_tilde, obliquity = (float(x) for x in orbit)
    rho = ctypes.c_double()
    scal = [lambda_m0, ecc, omega_tilde, obliquity, float(ysve), float(fday), float(solar_constant)]
    if type(lat).__module__.startswith("torch"):
        import torch
        la, lo = lat.to(dtype=torch.float64).contiguous().reshape(-1), lon.to(dtype=torch.float64).contiguous().reshape(-1)
        ins, zen = torch.empty_like(la), torch.empty_like(la)
        L.cb200_berger_run_device.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp] + [ctypes.c_double] * 7 + [_vp, _vp, _dp, _vp]
        rc = L.cb200_berger_run_device(la.device.index or 0, la.numel(), la.data_ptr(), lo.data_ptr(), *scal, ins.data_ptr(), zen.data_ptr(),
                                       ctypes.byref(rho), torch.cuda.current_stream().cuda_stream)
    else:
        la, lo = np.ascontiguousarray(lat, dtype=np.float64).reshape(-1), np.ascontiguousarray(lon, dtype=np.float64).reshape(-1)
        ins, zen = np.empty(la.size), np.empty(la.size)
        L.cb200_berger_run_host.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp] + [ctypes.c_double] * 7 + [_dp, _dp, _dp]
        rc = L.cb200_berger_run_host(device, la.size, la.ctypes.data_as(_dp), lo.ctypes.data_as(_dp), *scal, ins.ctypes.data_as(_dp),
                                     zen.ctypes.data_as(_dp), ctypes.byref(rho))
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())
    return ins, zen, rho.value


class BergerSolarInsolation(DiagnosticComponent):
    """Determines solar insolation using spectral solutions for orbital constants from Berger 1978, as climt's."""

    input_properties = {
        "longitude": {"dims": ["*"], "units": "degrees_east"},
        "latitude": {"dims": ["*"], "units": "degrees_north"},
    }
    diagnostic_properties = {
        "solar_insolation": {"dims": ["*"], "units": "W m^-2"},
        "solar_zenith_angle": {"dims": ["*"], "units": "radians"},
        "obliquity": {"dims": [], "units": "radians"},
        "eccentricity": {"dims": [], "units": "radians"},
        "normalized_earth_sun_distance": {"dims": [], "units": "dimensionless"},
    }

    def __init__(self, device=0, **kwargs):
        self._orbital_parameters = {}
        self._device = device
        _native.lib()
        super().__init__(**kwargs)

    def array_call(self, state):
        solar_constant = get_constant("stellar_irradiance", "W/m^2")
        lat, lon, time = state["latitude"], state["longitude"], state["time"]
        if time.year not in self._orbital_parameters:
            self._orbital_parameters[time.year] = orbital_parameters(float(time.year - 1950))
        orbit = self._orbital_parameters[time.year]
        ins, zen, rho = solar_parameters(orbit, years_since_vernal_equinox(time), fractional_day(time), lat, lon, solar_constant, self._device)
        shape = tuple(lat.shape)
        return {"solar_insolation": ins.reshape(shape), "solar_zenith_angle": zen.reshape(shape), "obliquity": orbit[3],
                "eccentricity": orbit[1], "normalized_earth_sun_distance": rho}
