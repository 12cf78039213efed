"""Emanuel moist convection -- drop-ins for climt.EmanuelConvection and climt.EmanuelConvectionPython.

Mirrors climt/_components/emanuel/component.py:18-340 (the Fortran-backed component) and
climt/_components/emanuel/pure_python_v3.py:48-208 (the numba port): same constructor keywords, property dictionaries
and ``array_call(state, timestep)`` return values.  The column routine (CONVECT 4.3c + TLIFT), the column loop and the
saturation-humidity pre-step run in one CUDA kernel (csrc/emanuel_engine.cu, per-thread code csrc/emanuel_core.cuh)
behind the C ABI of include/climt_b200.h; there is no CPU implementation here.
"""
import ctypes

import numpy as np

from . import _native
from .constants import get_constant
from .sympl_shim import ImplicitTendencyComponent

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)

PARAM_FIELDS = ("minorig", "elcrit", "tlcrit", "entp", "sigd", "sigs", "omtrain", "omtsnow", "coeffr", "coeffs", "cu", "beta",
                "dtmax", "alpha", "damp", "cpd", "cpv", "cl", "rv", "rd", "lv0", "g", "rowl", "delt0", "t_rain")
QS_GIVEN, QS_BOLTON, QS_PYTHON = 0, 1, 2
EM_IN = ("t", "q", "u", "v", "p", "ph", "qs", "cbmf")
EM_OUT = ("ft", "fq", "fu", "fv", "precip", "wd", "tprime", "qprime", "cbmf", "cape")


class EmanuelParams(ctypes.Structure):
    """cb200_emanuel_params (include/climt_b200.h)."""
    _fields_ = [(n, ctypes.c_double) for n in PARAM_FIELDS]


class EmanuelInputs(ctypes.Structure):
    _fields_ = [(n, _dp) for n in EM_IN]


class EmanuelOutputs(ctypes.Structure):
    _fields_ = [(n, _dp) for n in EM_OUT] + [("iflag", _ip)]


def make_params(**values):
    p = EmanuelParams()
    for n in PARAM_FIELDS:
        setattr(p, n, float(values[n]))
    return p


def _bind(L):
    vp = ctypes.c_void_p
    pi, po = ctypes.POINTER(EmanuelInputs), ctypes.POINTER(EmanuelOutputs)
    L.cb200_emanuel_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(EmanuelParams), ctypes.c_int]
    L.cb200_emanuel_destroy.argtypes = [vp]
    L.cb200_emanuel_destroy.restype = None
    L.cb200_emanuel_last_error.argtypes = [vp]
    L.cb200_emanuel_last_error.restype = ctypes.c_char_p
    L.cb200_emanuel_last_launches.argtypes = [vp]
    L.cb200_emanuel_enable_timing.argtypes = [vp, ctypes.c_int]
    L.cb200_emanuel_last_kernel_ms.argtypes = [vp]
    L.cb200_emanuel_last_kernel_ms.restype = ctypes.c_double
    L.cb200_emanuel_run_device.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                           pi, po, vp]
    L.cb200_emanuel_run_host.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, pi, po]


class EmanuelEngine:
    """Handle of one cb200_emanuel_engine (parameters are per instance)."""

    def __init__(self, params, device=0):
        self._L = _native.lib()
        _bind(self._L)
        self.params = params if isinstance(params, EmanuelParams) else make_params(**params)
        self._h = ctypes.c_void_p()
        rc = self._L.cb200_emanuel_create(ctypes.byref(self._h), ctypes.byref(self.params), device)
        if rc:
            raise (ValueError if rc == -3 else RuntimeError)(self._L.cb200_global_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.cb200_emanuel_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return self._L.cb200_emanuel_last_error(self._h).decode()

    @staticmethod
    def shapes(ncol, nlev, layout=1):
        """layout 1: the component's (ncol, nlev) arrays; layout 0: (nlev, ncol), the radiation engines' layout"""
        m, i = ((ncol, nlev), (ncol, nlev + 1)) if layout == 1 else ((nlev, ncol), (nlev + 1, ncol))
        ins = {"t": m, "q": m, "u": m, "v": m, "p": m, "ph": i, "qs": m, "cbmf": (ncol,)}
        outs = {"ft": m, "fq": m, "fu": m, "fv": m, "precip": (ncol,), "wd": (ncol,), "tprime": (ncol,), "qprime": (ncol,),
                "cbmf": (ncol,), "cape": (ncol,), "iflag": (ncol,)}
        return ins, outs

    def run_host(self, arrays, dt, qs_mode=QS_BOLTON, max_conv_lev=None, out=None):
        """arrays: t, q, u, v, p, ph [mbar], cbmf (+ qs when qs_mode = 0) in the component's (ncol, nlev) layout."""
        ncol, nlev = arrays["t"].shape
        nl = nlev - 3 if max_conv_lev is None else int(max_conv_lev)   # component.py:297
        ins, outs = self.shapes(ncol, nlev)
        pin, keep = EmanuelInputs(), []
        for k in EM_IN:
            a = arrays.get(k)
            if a is None:
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != ins[k]:
                raise ValueError(f"{k}: expected shape {ins[k]}, got {a.shape}")
            keep.append(a)
            setattr(pin, k, a.ctypes.data_as(_dp))
        if out is None:
            out = {k: np.empty(outs[k]) for k in EM_OUT}
            out["iflag"] = np.empty(ncol, dtype=np.int32)
        pout = EmanuelOutputs()
        for k in EM_OUT:
            a = out[k]
            if a.shape != outs[k] or a.dtype != np.float64 or not a.flags.c_contiguous:
                raise ValueError(f"output {k}: need C-contiguous float64 {outs[k]}")
            setattr(pout, k, a.ctypes.data_as(_dp))
        if out["iflag"].shape != (ncol,) or out["iflag"].dtype != np.int32:
            raise ValueError("output iflag: need int32 (ncol,)")
        pout.iflag = out["iflag"].ctypes.data_as(_ip)
        rc = self._L.cb200_emanuel_run_host(self._h, ncol, nlev, nl, float(dt), int(qs_mode), ctypes.byref(pin), ctypes.byref(pout))
        if rc:
            raise (ValueError if rc == -3 else RuntimeError)(self._err())
        return out

    def run_device(self, ncol, nlev, tensors, out, dt, qs_mode=QS_BOLTON, layout=0, max_conv_lev=None, stream=None):
        """tensors / out: float64 CUDA tensors (iflag: int32) with the shapes of `shapes(ncol, nlev, layout)`; asynchronous."""
        import torch
        nl = nlev - 3 if max_conv_lev is None else int(max_conv_lev)
        ins, outs = self.shapes(ncol, nlev, layout)
        pin, pout = EmanuelInputs(), EmanuelOutputs()
        for k, t in tensors.items():
            if t is None:
                continue
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == ins[k]):
                raise ValueError(f"{k}: need contiguous float64 CUDA tensor of shape {ins[k]}")
            setattr(pin, k, ctypes.cast(t.data_ptr(), _dp))
        for k in EM_OUT + ("iflag",):
            t = out[k]
            want = torch.int32 if k == "iflag" else torch.float64
            if not (t.is_cuda and t.dtype == want and t.is_contiguous() and tuple(t.shape) == outs[k]):
                raise ValueError(f"output {k}: need contiguous {want} CUDA tensor of shape {outs[k]}")
            setattr(pout, k, ctypes.cast(t.data_ptr(), _ip if k == "iflag" else _dp))
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        rc = self._L.cb200_emanuel_run_device(self._h, ncol, nlev, nl, float(dt), int(qs_mode), int(layout), ctypes.byref(pin),
                                              ctypes.byref(pout), ctypes.c_void_p(s))
        if rc:
            raise (ValueError if rc == -3 else RuntimeError)(self._err())

    def enable_timing(self, on=True):
        self._L.cb200_emanuel_enable_timing(self._h, 1 if on else 0)

    @property
    def last_kernel_ms(self):
        return self._L.cb200_emanuel_last_kernel_ms(self._h)

    @property
    def last_launches(self):
        return self._L.cb200_emanuel_last_launches(self._h)


# ---------------------------------------------------------------------------------------------------------------------
def _p(dims, units, **kw):
    return dict({"dims": dims, "units": units}, **kw)


_INPUTS = {
    "air_temperature": _p(["*", "mid_levels"], "degK"),
    "specific_humidity": _p(["*", "mid_levels"], "kg/kg"),
    "eastward_wind": _p(["*", "mid_levels"], "m s^-1"),
    "northward_wind": _p(["*", "mid_levels"], "m s^-1"),
    "air_pressure": _p(["*", "mid_levels"], "mbar"),
    "air_pressure_on_interface_levels": _p(["*", "interface_levels"], "mbar"),
    "cloud_base_mass_flux": _p(["*"], "kg m^-2 s^-1"),
}
_DIAGNOSTICS = {
    "convective_state": _p(["*"], "dimensionless", dtype=np.int32),
    "convective_precipitation_rate": _p(["*"], "mm day^-1"),
    "convective_downdraft_velocity_scale": _p(["*"], "m s^-1"),
    "convective_downdraft_temperature_scale": _p(["*"], "degK"),
    "convective_downdraft_specific_humidity_scale": _p(["*"], "kg/kg"),
    "cloud_base_mass_flux": _p(["*"], "kg m^-2 s^-1"),
    "atmosphere_convective_available_potential_energy": _p(["*"], "J kg^-1"),
    "air_temperature_tendency_from_convection": _p(["*", "mid_levels"], "degK day^-1"),
}
_TENDENCIES = {
    "air_temperature": {"units": "degK s^-1"},
    "specific_humidity": {"units": "kg/kg s^-1"},
    "eastward_wind": {"units": "m s^-2"},
    "northward_wind": {"units": "m s^-2"},
}


class _EmanuelBase(ImplicitTendencyComponent):
    input_properties = _INPUTS
    diagnostic_properties = _DIAGNOSTICS
    tendency_properties = _TENDENCIES
    _qs_mode = QS_BOLTON

    def _call_engine(self, state, timestep):
        T = state["air_temperature"]
        num_cols, num_levs = T.shape
        arrays = {"t": T, "q": state["specific_humidity"], "u": state["eastward_wind"], "v": state["northward_wind"],
                  "p": state["air_pressure"], "ph": state["air_pressure_on_interface_levels"]}
        cbmf = state.get("cloud_base_mass_flux")
        arrays["cbmf"] = np.zeros(num_cols) if cbmf is None else cbmf
        o = self._engine.run_host(arrays, timestep.total_seconds(), qs_mode=self._qs_mode, max_conv_lev=num_levs - 3)
        tendencies = {"air_temperature": o["ft"], "specific_humidity": o["fq"], "eastward_wind": o["fu"], "northward_wind": o["fv"]}
        diagnostics = {
            "convective_state": o["iflag"],
            "convective_precipitation_rate": o["precip"],
            "convective_downdraft_velocity_scale": o["wd"],
            "convective_downdraft_temperature_scale": o["tprime"],
            "convective_downdraft_specific_humidity_scale": o["qprime"],
            "cloud_base_mass_flux": o["cbmf"],
            "atmosphere_convective_available_potential_energy": o["cape"],
            "air_temperature_tendency_from_convection": o["ft"] * 86400.0,
        }
        return tendencies, diagnostics


class EmanuelConvection(_EmanuelBase):
    """Drop-in for climt.EmanuelConvection (climt/_components/emanuel/component.py:18-340): the Fortran CONVECT 4.3c."""
    _qs_mode = QS_BOLTON   # bolton_q_sat, component.py:306-311

    def __init__(self, minimum_convecting_layer=1, autoconversion_water_content_threshold=0.0011,
                 autoconversion_temperature_threshold=-55, entrainment_mixing_coefficient=1.5, downdraft_area_fraction=0.05,
                 precipitation_fraction_outside_cloud=0.12, speed_water_droplets=50.0, speed_snow=5.5,
                 rain_evaporation_coefficient=1.0, snow_evaporation_coefficient=0.8,
                 convective_momentum_transfer_coefficient=0.7, downdraft_surface_velocity_coefficient=10.0,
                 convection_bouyancy_threshold=0.9, mass_flux_relaxation_rate=0.1, mass_flux_damping_rate=0.1,
                 reference_mass_flux_timescale=300.0, device=0, **kwargs):
        if convective_momentum_transfer_coefficient < 0 or convective_momentum_transfer_coefficient > 1:
            raise ValueError("Momentum transfer coefficient must be between 0 and 1.")
        if downdraft_area_fraction < 0 or downdraft_area_fraction > 1:
            raise ValueError("Downdraft fraction must be between 0 and 1.")
        if precipitation_fraction_outside_cloud < 0 or precipitation_fraction_outside_cloud > 1:
            raise ValueError("Outside cloud precipitation fraction must be between 0 and 1.")
        # the constants _set_fortran_constants reads from sympl (component.py:236-246) -- including its
        # specific_enthalpy_of_vapor_phase for the liquid heat capacity CL
        self._params = dict(
            minorig=minimum_convecting_layer, elcrit=autoconversion_water_content_threshold,
            tlcrit=autoconversion_temperature_threshold, entp=entrainment_mixing_coefficient, sigd=downdraft_area_fraction,
            sigs=precipitation_fraction_outside_cloud, omtrain=speed_water_droplets, omtsnow=speed_snow,
            coeffr=rain_evaporation_coefficient, coeffs=snow_evaporation_coefficient,
            cu=convective_momentum_transfer_coefficient, beta=downdraft_surface_velocity_coefficient,
            dtmax=convection_bouyancy_threshold, alpha=mass_flux_relaxation_rate, damp=mass_flux_damping_rate,
            cpd=get_constant("heat_capacity_of_dry_air_at_constant_pressure", "J/kg/degK"),
            cpv=get_constant("heat_capacity_of_vapor_phase", "J/kg/degK"),
            cl=get_constant("specific_enthalpy_of_vapor_phase", "J/kg"),
            rv=get_constant("gas_constant_of_vapor_phase", "J/kg/degK"),
            rd=get_constant("gas_constant_of_dry_air", "J/kg/degK"),
            lv0=get_constant("latent_heat_of_condensation", "J/kg"),
            g=get_constant("gravitational_acceleration", "m/s^2"),
            rowl=get_constant("density_of_liquid_phase", "kg/m^3"),
            delt0=reference_mass_flux_timescale, t_rain=273.0)
        self._ntracers = 0
        self._engine = EmanuelEngine(self._params, device=device)
        super().__init__(**kwargs)

    def array_call(self, raw_state, timestep):
        return self._call_engine(raw_state, timestep)


class EmanuelConvectionPython(_EmanuelBase):
    """Drop-in for climt.EmanuelConvectionPython (climt/_components/emanuel/pure_python_v3.py:48-208) with water as the
    condensible (condensibles.py:22-29); other condensible species are not provided."""
    _qs_mode = QS_PYTHON   # compute_qs, pure_python_v3.py:165

    _DEFAULTS = dict(IPBL=0, MINORIG=1, ELCRIT=0.0011, TLCRIT=-55.0, ENTP=1.5, SIGD=0.05, SIGS=0.12, OMTRAIN=50.0, OMTSNOW=5.5,
                     COEFFR=1.0, COEFFS=0.8, CU=0.7, BETA=10.0, DTMAX=0.9, ALPHA=0.1, DAMP=0.1, CPD=1005.7, RD=287.04, G=9.8,
                     DELT0=300.0)

    def __init__(self, device=0, **kwargs):
        vals = dict(self._DEFAULTS)
        for key in list(kwargs):
            if key in vals:            # pure_python_v3.py:115-117: upper-case keywords override the defaults
                vals[key] = kwargs[key]
        for key, value in vals.items():
            setattr(self, key, value)
        if int(vals["IPBL"]) != 0:
            raise NotImplementedError("IPBL != 0 (dry adiabatic adjustment) is not provided by the CUDA engine")
        self._params = dict(
            minorig=int(vals["MINORIG"]), elcrit=vals["ELCRIT"], tlcrit=vals["TLCRIT"], entp=vals["ENTP"], sigd=vals["SIGD"],
            sigs=vals["SIGS"], omtrain=vals["OMTRAIN"], omtsnow=vals["OMTSNOW"], coeffr=vals["COEFFR"], coeffs=vals["COEFFS"],
            cu=vals["CU"], beta=vals["BETA"], dtmax=vals["DTMAX"], alpha=vals["ALPHA"], damp=vals["DAMP"], cpd=vals["CPD"],
            cpv=1870.0, cl=2500.0, rv=461.5, rd=vals["RD"], lv0=2.501e6, g=vals["G"], rowl=1000.0, delt0=vals["DELT0"], t_rain=273.15)
        self._engine = EmanuelEngine(self._params, device=device)
        super().__init__(**{k: v for k, v in kwargs.items() if k not in vals})

    def array_call(self, state, timestep):
        t, ph = state["air_temperature"], state["air_pressure_on_interface_levels"]
        if t.shape[1] != ph.shape[1] - 1:
            # pure_python_v3.py:151-159 also accepts (level, column) arrays; it returns the tendencies in that layout
            tr = {k: (v.T if getattr(v, "ndim", 0) == 2 else v) for k, v in state.items()}
            tend, diag = self._call_engine(tr, timestep)
            tend = {k: np.ascontiguousarray(v.T) for k, v in tend.items()}
            diag["air_temperature_tendency_from_convection"] = tend["air_temperature"] * 86400.0
            return tend, diag
        return self._call_engine(state, timestep)
