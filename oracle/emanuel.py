"""ctypes front-end of the C++ Emanuel-convection restatement (oracle/emanuel_oracle.cpp) plus the component glue.
TEST INFRASTRUCTURE ONLY.

`convect` is the column loop of _emanuel_convection.pyx:96-201 over CONVECT (convect43c.f90:146-1148);
`fortran_component_call` follows EmanuelConvection.array_call (climt/_components/emanuel/component.py:279-340) and
`python_component_call` EmanuelConvectionPython.array_call (climt/_components/emanuel/pure_python_v3.py:143-208).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)

PAR_FIELDS = ("minorig", "elcrit", "tlcrit", "entp", "sigd", "sigs", "omtrain", "omtsnow", "coeffr", "coeffs", "cu", "beta",
              "dtmax", "alpha", "damp", "cpd", "cpv", "cl", "rv", "rd", "lv0", "g", "rowl", "delt0", "t_rain")

# EmanuelConvection.__init__ defaults (component.py:100-118) with sympl's default constants (component.py:236-246)
FORTRAN_DEFAULTS = dict(minorig=1, elcrit=0.0011, tlcrit=-55.0, entp=1.5, sigd=0.05, sigs=0.12, omtrain=50.0, omtsnow=5.5,
                        coeffr=1.0, coeffs=0.8, cu=0.7, beta=10.0, dtmax=0.9, alpha=0.1, damp=0.1, delt0=300.0, t_rain=273.0)
# EmanuelConvectionPython.__init__ (pure_python_v3.py:95-117) + condensibles._H2O_DEFAULTS (condensibles.py:22-29)
PYTHON_DEFAULTS = dict(minorig=1, elcrit=0.0011, tlcrit=-55.0, entp=1.5, sigd=0.05, sigs=0.12, omtrain=50.0, omtsnow=5.5,
                       coeffr=1.0, coeffs=0.8, cu=0.7, beta=10.0, dtmax=0.9, alpha=0.1, damp=0.1, cpd=1005.7, cpv=1870.0,
                       cl=2500.0, rv=461.5, rd=287.04, lv0=2.501e6, g=9.8, rowl=1000.0, delt0=300.0, t_rain=273.15)


def build(force=False):
    so = os.path.join(_HERE, "liborc_emanuel.so")
    srcs = [os.path.join(_HERE, f) for f in ("emanuel_oracle.cpp", "ftn.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liborc_emanuel.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def convect(par, t, q, qs, u, v, p, ph, cbmf, dt, max_conv_lev=None):
    """par: dict with PAR_FIELDS.  Arrays (ncol, nlev) [ph: (ncol, nlev+1)], mbar.  Returns a dict; cbmf is not modified."""
    ncol, nlev = t.shape
    nl = nlev - 3 if max_conv_lev is None else max_conv_lev   # component.py:297
    pv = np.array([float(par[k]) for k in PAR_FIELDS])
    o = {k: np.zeros((ncol, nlev)) for k in ("ft", "fq", "fu", "fv")}
    o.update({k: np.zeros(ncol) for k in ("precip", "wd", "tprime", "qprime", "cape")})
    o["cbmf"] = _c(cbmf).copy()
    o["iflag"] = np.zeros(ncol, dtype=np.int32)
    lib().orc_emanuel_convect(_p(pv), ncol, nlev, nl, ctypes.c_double(dt), _p(_c(t)), _p(_c(q)), _p(_c(qs)), _p(_c(u)), _p(_c(v)),
                              _p(_c(p)), _p(_c(ph)), _p(o["cbmf"]), o["iflag"].ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                              _p(o["ft"]), _p(o["fq"]), _p(o["fu"]), _p(o["fv"]), _p(o["precip"]), _p(o["wd"]), _p(o["tprime"]),
                              _p(o["qprime"]), _p(o["cape"]))
    return o


def bolton_q_sat(T, p, Rd, Rh2O):
    """climt/_core/util.py:177-180 (p in Pa)"""
    es = 611.2 * np.exp(17.67 * (T - 273.15) / (T - 29.65))
    epsilon = Rd / Rh2O
    return epsilon * es / (p - (1 - epsilon) * es)


def python_qs(T, P, RD, RV):
    """compute_qs for water (climt/_core/condensibles.py:66-76, 104-123); P in hPa"""
    TC = T - 273.15
    with np.errstate(over="ignore", invalid="ignore"):
        es = np.where(TC >= 0.0, 6.112 * np.exp(17.67 * TC / (243.5 + TC)), np.exp(23.33086 - 6111.72784 / T + 0.15215 * np.log(T)))
    EPS = RD / RV
    return EPS * es / (P - (1.0 - EPS) * es)


def _diag(o):
    return {"convective_state": o["iflag"], "convective_precipitation_rate": o["precip"], "convective_downdraft_velocity_scale": o["wd"],
            "convective_downdraft_temperature_scale": o["tprime"], "convective_downdraft_specific_humidity_scale": o["qprime"],
            "cloud_base_mass_flux": o["cbmf"], "atmosphere_convective_available_potential_energy": o["cape"],
            "air_temperature_tendency_from_convection": o["ft"] * 86400.0}


def _tend(o):
    return {"air_temperature": o["ft"], "specific_humidity": o["fq"], "eastward_wind": o["fu"], "northward_wind": o["fv"]}


def python_component_call(state, dt, par=None):
    par = dict(PYTHON_DEFAULTS, **(par or {}))
    T, P = state["air_temperature"], state["air_pressure"]
    qs = python_qs(T, P, par["rd"], par["rv"])
    o = convect(par, T, state["specific_humidity"], qs, state["eastward_wind"], state["northward_wind"], P,
                state["air_pressure_on_interface_levels"], state["cloud_base_mass_flux"], dt)
    return _tend(o), _diag(o)


def fortran_component_call(state, dt, constants, par=None):
    """constants: dict with cpd, cpv, cl, rv, rd, lv0, g, rowl (sympl's values; component.py:236-246)"""
    par = dict(FORTRAN_DEFAULTS, **constants, **(par or {}))
    T, P = state["air_temperature"], state["air_pressure"]
    qs = bolton_q_sat(T, P * 100, par["rd"], par["rv"])
    o = convect(par, T, state["specific_humidity"], qs, state["eastward_wind"], state["northward_wind"], P,
                state["air_pressure_on_interface_levels"], state["cloud_base_mass_flux"], dt)
    return _tend(o), _diag(o)
