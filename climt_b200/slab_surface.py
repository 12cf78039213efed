"""SlabSurface -- drop-in for climt.SlabSurface (climt/_components/slab_surface.py:10-447): surface energy balance of a slab.
The per-column arithmetic of `_slab_surface_kernel_np` (:449-517) runs in `k_slab_surface` (csrc/adjacent_engine.cu).

Host path: numpy arrays in the component's dims (fluxes ("*", "interface_levels"), surface value = [..., 0]); only the surface
value of each flux column crosses PCIe.  Device path: torch CUDA tensors; the flux arguments may be the
radiation engines' (interface_levels, column) outputs (`flux_layout="level_major"`), read in place.

`include_ekman=True` (slab_surface.py:296-403) adds an Ekman heat-transport convergence from the wind-stress field.  That part
couples neighbouring columns through finite differences on the 2-D lat-lon grid and is host numpy in the reference
(`_core/horizontal_operators.py`); it is host numpy here too (`ekman_terms`), feeding the same column kernel.  Host path only:
a device-resident state with include_ekman=True is rejected.
"""
import ctypes

import numpy as np

from . import _native, device_state
from .sympl_shim import TendencyComponent

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p

AREA_MAP = {"land": 0, "land_ice": 1, "sea": 2, "sea_ice": 3}  # slab_surface.py:7

_FLUX = ("downwelling_shortwave_flux_in_air", "downwelling_longwave_flux_in_air", "upwelling_shortwave_flux_in_air",
         "upwelling_longwave_flux_in_air")
# field order of cb200_slab_inputs after the four fluxes (include/climt_b200.h)
_VEC_BEFORE = ("surface_upward_latent_heat_flux", "surface_upward_sensible_heat_flux")
_VEC_AFTER = ("upward_heat_flux_at_ground_level_in_soil", "heat_flux_into_sea_water_due_to_sea_ice", "sea_water_density",
              "surface_material_density", "heat_capacity_of_soil", "surface_thermal_capacity", "ocean_mixed_layer_thickness",
              "soil_layer_thickness", "ocean_heat_transport_convergence")


class SlabInputsHost(ctypes.Structure):
    _fields_ = ([(n, _dp) for n in ("sw_down", "lw_down", "sw_up", "lw_up", "lh", "sh")] + [("area_type", _ip)] +
                [(n, _dp) for n in ("up_heat_soil", "heat_flux_sea_ice", "sea_water_dens", "surf_dens", "heat_cap_soil",
                                    "surf_therm_cap", "ocean_mix_thick", "soil_layer_thick", "ocean_heat_transport")])


class SlabInputsDevice(ctypes.Structure):
    _fields_ = [(n, _vp) for n, _ in SlabInputsHost._fields_]


def _grads(field, lat, lon, radius):
    """d/dx, d/dy on the sphere by centred differences (climt/_core/horizontal_operators.py:19-33); zeros on degenerate grids"""
    latr, lonr = np.deg2rad(lat), np.deg2rad(lon)
    if field.shape[0] < 3 or field.shape[1] < 3:
        z = np.zeros_like(field, dtype=float)
        return z, z
    dfdlat = np.gradient(field, axis=0) / np.gradient(latr, axis=0)
    dfdlon = np.gradient(field, axis=1) / np.gradient(lonr, axis=1)
    return dfdlon / (radius * np.cos(latr)), dfdlat / radius


def ekman_terms(lat2d, lon2d, tau_x, tau_y, surface_temperature, sea_water_density, area_code, eq_cap_latitude, omega, c_sw, radius):
    """Ekman heat-transport convergence [W m-2] and Ekman pumping [m s-1] on the (lat, lon) grid, as SlabSurface.array_call
    forms them (slab_surface.py:296-403): wind stress zeroed outside open ocean before differentiating, Coriolis parameter
    capped below `eq_cap_latitude`, mass transport with the full 1/f variation, pumping with the local-f approximation."""
    open_ocean = area_code.reshape(lat2d.shape) == AREA_MAP["sea"]
    tau_x = np.where(open_ocean, tau_x, 0.0)
    tau_y = np.where(open_ocean, tau_y, 0.0)
    f = 2.0 * omega * np.sin(np.deg2rad(lat2d))
    f_floor = 2.0 * omega * np.sin(np.deg2rad(eq_cap_latitude))
    f_capped = np.where(f >= 0.0, 1.0, -1.0) * np.maximum(np.abs(f), f_floor)
    Mx, My = tau_y / f_capped, -tau_x / f_capped
    dtauy_dx, _ = _grads(tau_y, lat2d, lon2d, radius)
    _, dtaux_dy = _grads(tau_x, lat2d, lon2d, radius)
    w_ek = (dtauy_dx - dtaux_dy) / (f_capped * sea_water_density)
    dfx_dx, _ = _grads(surface_temperature * Mx, lat2d, lon2d, radius)
    _, dfy_dy = _grads(surface_temperature * My, lat2d, lon2d, radius)
    q_ekman = -c_sw * (dfx_dx + dfy_dy)
    return np.where(open_ocean, q_ekman, 0.0), np.where(open_ocean, w_ek, 0.0)


def area_type_codes(area_type):
    """strings of AREA_MAP (as the reference's state holds them) or integer codes -> int32 codes; unknown strings -> 0 like
    the reference's `np.zeros` + per-key assignment (slab_surface.py:176-178)"""
    a = np.asarray(area_type)
    if a.dtype.kind in "iu":
        return np.ascontiguousarray(a, dtype=np.int32)
    s = a.astype(str)
    code = np.zeros(s.shape, dtype=np.int32)
    for k, v in AREA_MAP.items():
        code[s == k] = v
    return code


def slab_surface_host(state, device=0):
    """state: numpy arrays keyed by the component's input names -> (tend_ts, depth), each (ncol,)"""
    L = _native.lib()
    code = area_type_codes(state["area_type"]).reshape(-1)
    n = code.size
    keep, s = [], SlabInputsHost()
    stride = None
    for fld, name in zip(("sw_down", "lw_down", "sw_up", "lw_up"), _FLUX):
        a = np.ascontiguousarray(state[name], dtype=np.float64)
        a = a.reshape(n, -1)  # ("*", "interface_levels"); a 1-D array is its own surface value (slab_surface.py:190-196)
        if stride is None:
            stride = a.shape[1]
        elif a.shape[1] != stride:
            raise ValueError("flux arrays differ in their number of interface levels")
        keep.append(a)
        setattr(s, fld, a.ctypes.data_as(_dp))
    names = dict(zip(("lh", "sh"), _VEC_BEFORE))
    names.update(zip(("up_heat_soil", "heat_flux_sea_ice", "sea_water_dens", "surf_dens", "heat_cap_soil", "surf_therm_cap",
                      "ocean_mix_thick", "soil_layer_thick", "ocean_heat_transport"), _VEC_AFTER))
    for fld, name in names.items():
        a = np.ascontiguousarray(state[name], dtype=np.float64).reshape(-1)
        if a.size != n:
            raise ValueError(f"{name}: {a.size} columns, expected {n}")
        keep.append(a)
        setattr(s, fld, a.ctypes.data_as(_dp))
    s.area_type = code.ctypes.data_as(_ip)
    tend, depth = np.empty(n), np.empty(n)
    L.cb200_slab_surface_run_host.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.POINTER(SlabInputsHost), _dp, _dp]
    rc = L.cb200_slab_surface_run_host(device, n, stride, ctypes.byref(s), tend.ctypes.data_as(_dp), depth.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())
    return tend, depth


def slab_surface_device(state, flux_layout="column_major", stream=None):
    """state: torch CUDA tensors; `area_type` an int32 tensor of AREA_MAP codes.  flux_layout "column_major": fluxes are
    (ncol, nlev+1) like the component's dims; "level_major": (nlev+1, ncol) as the radiation engines write them (row 0 = surface).
    Asynchronous on the current stream -> (tend_ts, depth) CUDA tensors."""
    import torch
    L = _native.lib()
    code = state["area_type"]
    if code.dtype != torch.int32:
        code = code.to(torch.int32)
    code = code.contiguous().reshape(-1)
    n = code.numel()
    keep, s = [code], SlabInputsDevice()
    stride = None
    for fld, name in zip(("sw_down", "lw_down", "sw_up", "lw_up"), _FLUX):
        a = state[name].to(dtype=torch.float64).contiguous()
        st = 1 if (flux_layout == "level_major" or a.dim() == 1) else a.reshape(n, -1).shape[1]
        if stride is None:
            stride = st
        elif st != stride:
            raise ValueError("flux arrays differ in layout")
        keep.append(a)
        setattr(s, fld, a.data_ptr())
    names = dict(zip(("lh", "sh"), _VEC_BEFORE))
    names.update(zip(("up_heat_soil", "heat_flux_sea_ice", "sea_water_dens", "surf_dens", "heat_cap_soil", "surf_therm_cap",
                      "ocean_mix_thick", "soil_layer_thick", "ocean_heat_transport"), _VEC_AFTER))
    for fld, name in names.items():
        a = state[name].to(dtype=torch.float64).contiguous().reshape(-1)
        if a.numel() != n:
            raise ValueError(f"{name}: {a.numel()} columns, expected {n}")
        keep.append(a)
        setattr(s, fld, a.data_ptr())
    s.area_type = code.data_ptr()
    tend = torch.empty(n, dtype=torch.float64, device=code.device)
    depth = torch.empty_like(tend)
    L.cb200_slab_surface_run_device.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.POINTER(SlabInputsDevice),
                                                _vp, _vp, _vp]
    sp = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    rc = L.cb200_slab_surface_run_device(code.device.index or 0, n, stride, ctypes.byref(s), tend.data_ptr(), depth.data_ptr(), sp)
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())
    device_state.keep_alive_on(stream, keep + [tend, depth])
    return tend, depth


class SlabSurface(TendencyComponent):
    """Surface energy balance of a slab of possibly varying heat capacity, as climt.SlabSurface."""

    input_properties = {
        "downwelling_longwave_flux_in_air": {"dims": ["*", "interface_levels"], "units": "W m^-2"},
        "downwelling_shortwave_flux_in_air": {"dims": ["*", "interface_levels"], "units": "W m^-2"},
        "upwelling_longwave_flux_in_air": {"dims": ["*", "interface_levels"], "units": "W m^-2"},
        "upwelling_shortwave_flux_in_air": {"dims": ["*", "interface_levels"], "units": "W m^-2"},
        "surface_upward_latent_heat_flux": {"dims": ["*"], "units": "W m^-2"},
        "surface_temperature": {"dims": ["*"], "units": "degK"},
        "surface_upward_sensible_heat_flux": {"dims": ["*"], "units": "W m^-2"},
        "surface_thermal_capacity": {"dims": ["*"], "units": "J kg^-1 degK^-1"},
        "surface_material_density": {"dims": ["*"], "units": "kg m^-3"},
        "upward_heat_flux_at_ground_level_in_soil": {"dims": ["*"], "units": "W m^-2"},
        "heat_flux_into_sea_water_due_to_sea_ice": {"dims": ["*"], "units": "W m^-2"},
        "area_type": {"dims": ["*"], "units": "dimensionless"},
        "soil_layer_thickness": {"dims": ["*"], "units": "m"},
        "ocean_mixed_layer_thickness": {"dims": ["*"], "units": "m"},
        "heat_capacity_of_soil": {"dims": ["*"], "units": "J kg^-1 degK^-1"},
        "sea_water_density": {"dims": ["*"], "units": "kg m^-3"},
        "ocean_heat_transport_convergence": {"dims": ["*"], "units": "W m^-2"},
    }
    tendency_properties = {"surface_temperature": {"dims": ["*"], "units": "degK s^-1"}}
    diagnostic_properties = {
        "depth_of_slab_surface": {"dims": ["*"], "units": "m"},
        "ocean_heat_transport_convergence": {"dims": ["*"], "units": "W m^-2"},
    }

    def __init__(self, include_ekman=False, equatorial_ekman_cap_latitude=5.0, device=0, flux_layout="column_major", **kwargs):
        self._include_ekman, self._eq_cap = include_ekman, equatorial_ekman_cap_latitude
        self._device, self._flux_layout = device, flux_layout
        if include_ekman:  # slab_surface.py:125-158: four 2-D inputs and two diagnostics more
            self.input_properties = dict(self.input_properties)
            self.input_properties.update({
                "surface_downward_eastward_stress": {"dims": ["lat", "lon"], "units": "N m^-2"},
                "surface_downward_northward_stress": {"dims": ["lat", "lon"], "units": "N m^-2"},
                "latitude": {"dims": ["lat", "lon"], "units": "degrees_north"},
                "longitude": {"dims": ["lat", "lon"], "units": "degrees_east"},
            })
            self.diagnostic_properties = dict(self.diagnostic_properties)
            self.diagnostic_properties.update({
                "ekman_heat_transport_convergence": {"dims": ["*"], "units": "W m^-2"},
                "ekman_pumping": {"dims": ["*"], "units": "m s^-1"},
            })
        _native.lib()
        super().__init__(**kwargs)

    def _ekman(self, state, code):
        """(q_ekman, w_ek) flattened to the column axis (slab_surface.py:296-403)"""
        from .constants import get_constant
        lat2d = np.asarray(state["latitude"], dtype=float)
        lon2d = np.asarray(state["longitude"], dtype=float)
        if lat2d.ndim == 1:  # a flattened single-column view: the operators return zeros below 3 points per dimension
            lat2d, lon2d = lat2d.reshape(-1, 1), lon2d.reshape(-1, 1)
        shp = lat2d.shape
        q, w = ekman_terms(lat2d, lon2d, np.asarray(state["surface_downward_eastward_stress"], dtype=float).reshape(shp),
                           np.asarray(state["surface_downward_northward_stress"], dtype=float).reshape(shp),
                           np.asarray(state["surface_temperature"], dtype=float).reshape(shp),
                           np.asarray(state["sea_water_density"], dtype=float).reshape(shp), code, self._eq_cap,
                           get_constant("planetary_rotation_rate", "s^-1"), get_constant("heat_capacity_of_sea_water", "J/kg/degK"),
                           get_constant("planetary_radius", "m"))
        return q.reshape(-1), w.reshape(-1)

    def array_call(self, state):
        at = state["area_type"]
        if type(at).__module__.startswith("torch") and getattr(at, "is_cuda", False):
            if self._include_ekman:
                raise NotImplementedError("SlabSurface(include_ekman=True) on a device-resident state: the Ekman terms are finite "
                                          "differences across columns, evaluated on the host as in the reference")
            tend, depth = slab_surface_device(state, self._flux_layout)
            oht = state["ocean_heat_transport_convergence"]
            return ({"surface_temperature": tend.reshape(at.shape)},
                    {"depth_of_slab_surface": depth.reshape(at.shape), "ocean_heat_transport_convergence": oht.reshape(at.shape)})
        shape = np.asarray(at).shape
        oht = np.asarray(state["ocean_heat_transport_convergence"], dtype=np.float64)
        extra = {}
        if self._include_ekman:
            q_ek, w_ek = self._ekman(state, area_type_codes(at))
            oht = oht.reshape(-1) + q_ek   # the total q-flux applied to sea cells, also what the diagnostic reports (:423-436)
            state = dict(state, ocean_heat_transport_convergence=oht)
            extra = {"ekman_heat_transport_convergence": q_ek.reshape(shape), "ekman_pumping": w_ek.reshape(shape)}
        tend, depth = slab_surface_host(state, self._device)
        return ({"surface_temperature": tend.reshape(shape)},
                dict({"depth_of_slab_surface": depth.reshape(shape), "ocean_heat_transport_convergence": oht.reshape(shape)}, **extra))
