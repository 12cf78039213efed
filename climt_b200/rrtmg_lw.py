"""RRTMGLongwave -- drop-in for climt.RRTMGLongwave (climt/_components/rrtmg/lw/component.py:30-522).

Same class attributes (`input_properties`, `tendency_properties`, `diagnostic_properties`), same
constructor keywords, same `array_call(state) -> (tendencies, diagnostics)` contract on raw numpy
arrays shaped per the property dims.  The Cython/Fortran call is replaced by the CUDA engine behind
include/climt_b200.h; there is no CPU fallback.
"""
import logging

import numpy as np

from . import device_state
from .constants import rrtmg_constants
from .engine import LWEngine
from .rrtmg_common import (rrtmg_cloud_ice_props_dict, rrtmg_cloud_liquid_props_dict,
                           rrtmg_cloud_overlap_method_dict, rrtmg_cloud_props_dict, rrtmg_random_number_dict)
from .state import get_interface_values, mass_to_volume_mixing_ratio
from .rrtmg_common import allocate_outputs
from .sympl_shim import TendencyComponent


def _p(dims, units):
    return {"dims": list(dims), "units": units}


class RRTMGLongwave(TendencyComponent):
    """The Rapid Radiative Transfer Model (RRTMG), longwave, on a B200."""

    num_longwave_bands = 16
    num_reduced_g_intervals = 140
    rrtm_iplon = 1

    # climt/_components/rrtmg/lw/component.py:42-131
    input_properties = {
        "air_pressure": _p(["mid_levels", "*"], "mbar"),
        "air_pressure_on_interface_levels": _p(["interface_levels", "*"], "mbar"),
        "air_temperature": _p(["mid_levels", "*"], "degK"),
        "surface_temperature": _p(["*"], "degK"),
        "specific_humidity": _p(["mid_levels", "*"], "g/g"),
        "mole_fraction_of_ozone_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_carbon_dioxide_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_methane_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_nitrous_oxide_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_oxygen_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_cfc11_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_cfc12_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_cfc22_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_carbon_tetrachloride_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "surface_longwave_emissivity": _p(["num_longwave_bands", "*"], "dimensionless"),
        "cloud_area_fraction_in_atmosphere_layer": _p(["mid_levels", "*"], "dimensionless"),
        "longwave_optical_thickness_due_to_cloud": _p(["mid_levels", "*", "num_longwave_bands"], "dimensionless"),
        "mass_content_of_cloud_ice_in_atmosphere_layer": _p(["mid_levels", "*"], "g m^-2"),
        "mass_content_of_cloud_liquid_water_in_atmosphere_layer": _p(["mid_levels", "*"], "g m^-2"),
        "cloud_ice_particle_size": _p(["mid_levels", "*"], "micrometer"),
        "cloud_water_droplet_radius": _p(["mid_levels", "*"], "micrometer"),
        "longwave_optical_thickness_due_to_aerosol": _p(["num_longwave_bands", "mid_levels", "*"], "dimensionless"),
    }
    tendency_properties = {"air_temperature": _p(["mid_levels", "*"], "degK day^-1")}
    diagnostic_properties = {
        "upwelling_longwave_flux_in_air": _p(["interface_levels", "*"], "W m^-2"),
        "downwelling_longwave_flux_in_air": _p(["interface_levels", "*"], "W m^-2"),
        "upwelling_longwave_flux_in_air_assuming_clear_sky": _p(["interface_levels", "*"], "W m^-2"),
        "downwelling_longwave_flux_in_air_assuming_clear_sky": _p(["interface_levels", "*"], "W m^-2"),
        "air_temperature_tendency_from_longwave_assuming_clear_sky": _p(["mid_levels", "*"], "degK day^-1"),
        "air_temperature_tendency_from_longwave": _p(["mid_levels", "*"], "degK day^-1"),
    }

    def __init__(self, calculate_change_up_flux=False, cloud_overlap_method=None,
                 cloud_optical_properties="liquid_and_ice_clouds", cloud_ice_properties="ebert_curry_two",
                 cloud_liquid_water_properties="radius_dependent_absorption", calculate_interface_temperature=True,
                 mcica=False, random_number_generator="mersenne_twister", device=0, asynchronous=False, **kwargs):
        self.input_properties = dict(RRTMGLongwave.input_properties)
        self._asynchronous = asynchronous  # device-resident calls only: do not synchronise / validate after each call
        self._calc_dflxdt = 1 if calculate_change_up_flux else 0
        self._mcica = mcica
        if mcica:
            self._permute_seed = None
            self._random_number_generator = rrtmg_random_number_dict[random_number_generator.lower()]
            if type(cloud_overlap_method) is str and cloud_overlap_method.lower() == "clear_only":
                logging.info("cloud_overlap_method == 'clear_only'. This overrides all other properties. "
                             "There are no clouds.")
            if cloud_optical_properties.lower() == "single_cloud_type":
                logging.warning("cloud_optical_properties must be 'direct_input' or 'liquid_and_ice_clouds' "
                                "for radiative calculations with clouds using McICA.")
        if cloud_overlap_method is None:
            cloud_overlap_method = "random"
        self._cloud_overlap = rrtmg_cloud_overlap_method_dict[cloud_overlap_method.lower()]
        self._cloud_optics = rrtmg_cloud_props_dict[cloud_optical_properties.lower()]
        self._ice_props = rrtmg_cloud_ice_props_dict[cloud_ice_properties.lower()]
        self._liq_props = rrtmg_cloud_liquid_props_dict[cloud_liquid_water_properties.lower()]
        self._calc_Tint = calculate_interface_temperature
        if not self._calc_Tint:
            self.input_properties["air_temperature_on_interface_levels"] = _p(["interface_levels", "*"], "degK")
        # constants are captured per instance at construction (climt reads sympl's registry here, :298-309)
        self._engine = LWEngine(rrtmg_constants(), device=device, icld=self._cloud_overlap, idrv=self._calc_dflxdt,
                                inflag=self._cloud_optics, iceflag=self._ice_props, liqflag=self._liq_props,
                                mcica=bool(mcica), irng=getattr(self, "_random_number_generator", 1))
        super().__init__(**kwargs)

    def array_call(self, state):
        if device_state.is_device_state(state):
            return self._array_call_device(state)
        state = {k: np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v for k, v in state.items()}
        n_layers, n_columns = state["air_temperature"].shape
        on_device = device_state.host_marshal_on_device()
        # on_device: the engine converts q and interpolates the interface temperatures on the device, chunk by chunk (the two numpy
        # expressions cost more host time than the whole call takes on the GPU); CLIMT_B200_HOST_MARSHAL=numpy: as the reference
        Q = state["specific_humidity"] if on_device else mass_to_volume_mixing_ratio(state["specific_humidity"], 18.02)
        if not self._calc_Tint:
            T_interface = state["air_temperature_on_interface_levels"]
        elif on_device:
            T_interface = None
        else:
            T_interface = get_interface_values(state["air_temperature"], state["surface_temperature"],
                                               state["air_pressure"], state["air_pressure_on_interface_levels"])
        self._engine.set_host_marshal(on_device, on_device and self._calc_Tint)
        diagnostics = allocate_outputs(self.diagnostic_properties, state, self.input_properties)
        tendencies = allocate_outputs(self.tendency_properties, state, self.input_properties)
        arrays = {
            "play": state["air_pressure"], "plev": state["air_pressure_on_interface_levels"],
            "tlay": state["air_temperature"], "tlev": T_interface, "tsfc": state["surface_temperature"],
            "h2ovmr": Q, "o3vmr": state["mole_fraction_of_ozone_in_air"],
            "co2vmr": state["mole_fraction_of_carbon_dioxide_in_air"],
            "ch4vmr": state["mole_fraction_of_methane_in_air"],
            "n2ovmr": state["mole_fraction_of_nitrous_oxide_in_air"],
            "o2vmr": state["mole_fraction_of_oxygen_in_air"],
            "cfc11vmr": state["mole_fraction_of_cfc11_in_air"], "cfc12vmr": state["mole_fraction_of_cfc12_in_air"],
            "cfc22vmr": state["mole_fraction_of_cfc22_in_air"],
            "ccl4vmr": state["mole_fraction_of_carbon_tetrachloride_in_air"],
            "emis": state["surface_longwave_emissivity"],
            "cldfr": state["cloud_area_fraction_in_atmosphere_layer"],
            "taucld": state["longwave_optical_thickness_due_to_cloud"],
            "cicewp": state["mass_content_of_cloud_ice_in_atmosphere_layer"],
            "cliqwp": state["mass_content_of_cloud_liquid_water_in_atmosphere_layer"],
            "reice": state["cloud_ice_particle_size"], "reliq": state["cloud_water_droplet_radius"],
            "tauaer": state["longwave_optical_thickness_due_to_aerosol"],
        }
        out = {
            "uflx": diagnostics["upwelling_longwave_flux_in_air"],
            "dflx": diagnostics["downwelling_longwave_flux_in_air"],
            "hr": tendencies["air_temperature"],
            "uflxc": diagnostics["upwelling_longwave_flux_in_air_assuming_clear_sky"],
            "dflxc": diagnostics["downwelling_longwave_flux_in_air_assuming_clear_sky"],
            "hrc": diagnostics["air_temperature_tendency_from_longwave_assuming_clear_sky"],
        }
        if self._mcica:
            # same draw as the reference (lw/component.py:415-423): every call gets a new seed from numpy's global
            # legacy generator, so `np.random.seed(k)` before the call makes it reproducible.  Unlike the
            # reference, the k-tables are NOT re-initialised on every call (:425-434).
            if self._random_number_generator == 0:
                self._permute_seed = np.random.randint(0, 1024)
            else:
                self._permute_seed = np.random.randint(0, 2 ** 31 - 1)
            self._engine.set_mcica(True, self._random_number_generator, self._permute_seed)
        self._engine.run_host(n_columns, n_layers, arrays, out)
        if self._calc_dflxdt:
            # d(upward flux)/d(surface temperature) (idrv = 1).  The reference has no diagnostic for it (its component never hands
            # the two arrays to the Cython call, lw/component.py:482-516 vs _rrtmg_lw.pyx:164-165), so it is kept on the instance.
            self.change_up_flux = {"duflx_dt": out["duflx_dt"], "duflxc_dt": out["duflxc_dt"]}
        diagnostics["air_temperature_tendency_from_longwave"] = tendencies["air_temperature"]
        return tendencies, diagnostics

    _ABI_FROM_STATE = {
        "play": "air_pressure", "plev": "air_pressure_on_interface_levels", "tlay": "air_temperature",
        "tsfc": "surface_temperature", "o3vmr": "mole_fraction_of_ozone_in_air",
        "co2vmr": "mole_fraction_of_carbon_dioxide_in_air", "ch4vmr": "mole_fraction_of_methane_in_air",
        "n2ovmr": "mole_fraction_of_nitrous_oxide_in_air", "o2vmr": "mole_fraction_of_oxygen_in_air",
        "cfc11vmr": "mole_fraction_of_cfc11_in_air", "cfc12vmr": "mole_fraction_of_cfc12_in_air",
        "cfc22vmr": "mole_fraction_of_cfc22_in_air", "ccl4vmr": "mole_fraction_of_carbon_tetrachloride_in_air",
        "emis": "surface_longwave_emissivity", "cldfr": "cloud_area_fraction_in_atmosphere_layer",
        "taucld": "longwave_optical_thickness_due_to_cloud", "cicewp": "mass_content_of_cloud_ice_in_atmosphere_layer",
        "cliqwp": "mass_content_of_cloud_liquid_water_in_atmosphere_layer", "reice": "cloud_ice_particle_size",
        "reliq": "cloud_water_droplet_radius", "tauaer": "longwave_optical_thickness_due_to_aerosol",
    }

    def _array_call_device(self, state):
        """State of torch CUDA tensors (already in this component's units and (levels, columns) layout): everything stays in
        HBM; returns torch CUDA tensors.  See climt_b200/device_state.py."""
        import torch
        st = {k: (device_state.dense(v) if isinstance(v, torch.Tensor) else v) for k, v in state.items()}
        n_layers, n_columns = st["air_temperature"].shape
        Q, T_interface, _ = device_state.marshal(st["specific_humidity"], st["air_temperature"], st["surface_temperature"],
                                                 st["air_pressure"], st["air_pressure_on_interface_levels"],
                                                 want_tlev=self._calc_Tint)
        if not self._calc_Tint:
            T_interface = st["air_temperature_on_interface_levels"]
        tensors = {k: st[v] for k, v in self._ABI_FROM_STATE.items()}
        tensors.update(h2ovmr=Q, tlev=T_interface)
        dev = st["air_temperature"].device
        new = lambda nlev: torch.empty((nlev, n_columns), dtype=torch.float64, device=dev)  # noqa: E731
        out = {"uflx": new(n_layers + 1), "dflx": new(n_layers + 1), "hr": new(n_layers), "uflxc": new(n_layers + 1),
               "dflxc": new(n_layers + 1), "hrc": new(n_layers)}
        if self._calc_dflxdt:
            out["duflx_dt"], out["duflxc_dt"] = new(n_layers + 1), new(n_layers + 1)
            self.change_up_flux = {"duflx_dt": out["duflx_dt"], "duflxc_dt": out["duflxc_dt"]}
        if self._mcica:
            self._permute_seed = np.random.randint(0, 1024) if self._random_number_generator == 0 else np.random.randint(0, 2 ** 31 - 1)
            self._engine.set_mcica(True, self._random_number_generator, self._permute_seed)
        self._engine.run_device(n_columns, n_layers, tensors, out)
        device_state.finish(self._engine, self._asynchronous)
        tendencies = {"air_temperature": out["hr"]}
        diagnostics = {
            "upwelling_longwave_flux_in_air": out["uflx"], "downwelling_longwave_flux_in_air": out["dflx"],
            "upwelling_longwave_flux_in_air_assuming_clear_sky": out["uflxc"],
            "downwelling_longwave_flux_in_air_assuming_clear_sky": out["dflxc"],
            "air_temperature_tendency_from_longwave_assuming_clear_sky": out["hrc"],
            "air_temperature_tendency_from_longwave": out["hr"],
        }
        return tendencies, diagnostics
