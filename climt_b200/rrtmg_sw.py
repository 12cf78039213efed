"""RRTMGShortwave -- drop-in for climt.RRTMGShortwave (climt/_components/rrtmg/sw/component.py:32-668).

Same properties, constructor keywords, log messages and `array_call` contract; the Cython/Fortran call is
replaced by the CUDA engine behind include/climt_b200.h (no CPU fallback).
"""
import logging

import numpy as np

from . import device_state
from .constants import get_constant, rrtmg_constants
from .engine import SWEngine
from .rrtmg_common import (rrtmg_aerosol_input_dict, rrtmg_cloud_ice_props_dict, rrtmg_cloud_liquid_props_dict,
                           rrtmg_cloud_overlap_method_dict, rrtmg_cloud_props_dict, rrtmg_random_number_dict)
from .state import get_interface_values, mass_to_volume_mixing_ratio
from .rrtmg_common import allocate_outputs
from .sympl_shim import TendencyComponent


def _p(dims, units):
    return {"dims": list(dims), "units": units}


class RRTMGShortwave(TendencyComponent):
    """The Rapid Radiative Transfer Model (RRTMG), shortwave, on a B200."""

    num_shortwave_bands = 14
    num_ecmwf_aerosols = 6
    num_reduced_g_intervals = 112
    rrtm_iplon = 1

    # climt/_components/rrtmg/sw/component.py:46-152
    input_properties = {
        "air_pressure": _p(["mid_levels", "*"], "mbar"),
        "air_pressure_on_interface_levels": _p(["interface_levels", "*"], "mbar"),
        "air_temperature": _p(["mid_levels", "*"], "degK"),
        "specific_humidity": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_ozone_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_carbon_dioxide_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_methane_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_nitrous_oxide_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mole_fraction_of_oxygen_in_air": _p(["mid_levels", "*"], "dimensionless"),
        "mass_content_of_cloud_ice_in_atmosphere_layer": _p(["mid_levels", "*"], "g m^-2"),
        "mass_content_of_cloud_liquid_water_in_atmosphere_layer": _p(["mid_levels", "*"], "g m^-2"),
        "cloud_ice_particle_size": _p(["mid_levels", "*"], "micrometer"),
        "cloud_water_droplet_radius": _p(["mid_levels", "*"], "micrometer"),
        "cloud_area_fraction_in_atmosphere_layer": _p(["mid_levels", "*"], "dimensionless"),
        "surface_temperature": _p(["*"], "degK"),
        "zenith_angle": _p(["*"], "radians"),
        "surface_albedo_for_direct_shortwave": _p(["*"], "dimensionless"),
        "surface_albedo_for_direct_near_infrared": _p(["*"], "dimensionless"),
        "surface_albedo_for_diffuse_near_infrared": _p(["*"], "dimensionless"),
        "surface_albedo_for_diffuse_shortwave": _p(["*"], "dimensionless"),
        "shortwave_optical_thickness_due_to_cloud": _p(["mid_levels", "*", "num_shortwave_bands"], "dimensionless"),
        "shortwave_optical_thickness_due_to_aerosol": _p(["num_shortwave_bands", "mid_levels", "*"], "dimensionless"),
        "single_scattering_albedo_due_to_cloud": _p(["mid_levels", "*", "num_shortwave_bands"], "dimensionless"),
        "single_scattering_albedo_due_to_aerosol": _p(["num_shortwave_bands", "mid_levels", "*"], "dimensionless"),
        "cloud_asymmetry_parameter": _p(["mid_levels", "*", "num_shortwave_bands"], "dimensionless"),
        "aerosol_asymmetry_parameter": _p(["num_shortwave_bands", "mid_levels", "*"], "dimensionless"),
        "cloud_forward_scattering_fraction": _p(["mid_levels", "*", "num_shortwave_bands"], "dimensionless"),
        "aerosol_optical_depth_at_55_micron": _p(["num_ecmwf_aerosols", "mid_levels", "*"], "dimensionless"),
        "solar_cycle_fraction": _p([], "dimensionless"),
        "flux_adjustment_for_earth_sun_distance": _p([], "dimensionless"),
    }
    tendency_properties = {"air_temperature": {"units": "degK day^-1"}}  # dims follow the input of the same name, as in the reference
    diagnostic_properties = {
        "upwelling_shortwave_flux_in_air": _p(["interface_levels", "*"], "W m^-2"),
        "downwelling_shortwave_flux_in_air": _p(["interface_levels", "*"], "W m^-2"),
        "upwelling_shortwave_flux_in_air_assuming_clear_sky": _p(["interface_levels", "*"], "W m^-2"),
        "downwelling_shortwave_flux_in_air_assuming_clear_sky": _p(["interface_levels", "*"], "W m^-2"),
        "air_temperature_tendency_from_shortwave_assuming_clear_sky": _p(["mid_levels", "*"], "degK day^-1"),
        "air_temperature_tendency_from_shortwave": _p(["mid_levels", "*"], "degK day^-1"),
    }

    def __init__(self, cloud_overlap_method=None, cloud_optical_properties="liquid_and_ice_clouds",
                 cloud_ice_properties="ebert_curry_two", cloud_liquid_water_properties="radius_dependent_absorption",
                 solar_variability_method=0, use_solar_constant_from_fortran=False, ignore_day_of_year=False,
                 facular_sunspot_amplitude=None, solar_variability_by_band=None, aerosol_type="no_aerosol", mcica=False,
                 random_number_generator="mersenne_twister", device=0, asynchronous=False, **kwargs):
        self._asynchronous = asynchronous  # device-resident calls only: do not synchronise / validate after each call
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
            if cloud_optical_properties.lower() == "liquid_and_ice_clouds":
                if cloud_ice_properties.lower() == "ebert_curry_one":
                    logging.warning("cloud_ice_properties should not be set to 'ebert_curry_one' for shortwave "
                                    "calculations with McICA.")
                if cloud_liquid_water_properties.lower() == "radius_independent_absorption":
                    logging.warning("cloud_liquid_water_properties must be set to 'radius_dependent_absorption' "
                                    "for use with McICA in the shortwave.")
        if cloud_overlap_method is None:
            cloud_overlap_method = "random"
        self._cloud_overlap = rrtmg_cloud_overlap_method_dict[cloud_overlap_method.lower()]
        self._cloud_optics = rrtmg_cloud_props_dict[cloud_optical_properties.lower()]
        self._ice_props = rrtmg_cloud_ice_props_dict[cloud_ice_properties.lower()]
        self._liq_props = rrtmg_cloud_liquid_props_dict[cloud_liquid_water_properties.lower()]
        self._solar_var_flag = solar_variability_method
        self._ignore_day_of_year = ignore_day_of_year
        self._fac_sunspot_coeff = np.ones(2) if facular_sunspot_amplitude is None else np.asarray(facular_sunspot_amplitude, dtype=float)
        self._solar_var_by_band = np.ones(16) if solar_variability_by_band is None else np.asarray(solar_variability_by_band, dtype=float)
        self._aerosol_type = rrtmg_aerosol_input_dict[aerosol_type.lower()]
        self._solar_const = 0.0 if use_solar_constant_from_fortran else get_constant("stellar_irradiance")
        self._engine = SWEngine(rrtmg_constants(), device=device, icld=self._cloud_overlap, iaer=self._aerosol_type,
                                inflag=self._cloud_optics, iceflag=self._ice_props, liqflag=self._liq_props,
                                isolvar=self._solar_var_flag, scon=self._solar_const,
                                indsolvar=self._fac_sunspot_coeff, bndsolvar=self._solar_var_by_band,
                                mcica=bool(mcica), irng=getattr(self, "_random_number_generator", 1))
        super().__init__(**kwargs)

    def array_call(self, state):
        if device_state.is_device_state(state):
            return self._array_call_device(state)
        st = {k: np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v for k, v in state.items()}
        assert st["air_pressure"].shape[0] + 1 == st["air_pressure_on_interface_levels"].shape[0]
        if device_state.host_marshal_on_device():
            # the engine converts q and interpolates the interface temperatures on the device, chunk by chunk (the two numpy
            # expressions cost more host time than the whole call takes on the GPU); CLIMT_B200_HOST_MARSHAL=numpy: as the reference
            Q, Tint = st["specific_humidity"], None
            self._engine.set_host_marshal(True, True)
        else:
            Q = mass_to_volume_mixing_ratio(st["specific_humidity"], 18.02)
            Tint = get_interface_values(st["air_temperature"], st["surface_temperature"], st["air_pressure"],
                                        st["air_pressure_on_interface_levels"])
            self._engine.set_host_marshal(False, False)
        diagnostics = allocate_outputs(self.diagnostic_properties, st, self.input_properties)
        tendencies = allocate_outputs(self.tendency_properties, st, self.input_properties)
        if self._ignore_day_of_year:
            day_of_year = 0
        else:
            t = st.get("time")
            day_of_year = t.timetuple().tm_yday if t is not None else int(st.get("day_of_year", 1))
        n_layers, n_columns = st["air_temperature"].shape
        arrays = {
            "play": st["air_pressure"], "plev": st["air_pressure_on_interface_levels"], "tlay": st["air_temperature"],
            "tlev": Tint, "tsfc": st["surface_temperature"], "h2ovmr": Q,
            "o3vmr": st["mole_fraction_of_ozone_in_air"], "co2vmr": st["mole_fraction_of_carbon_dioxide_in_air"],
            "ch4vmr": st["mole_fraction_of_methane_in_air"], "n2ovmr": st["mole_fraction_of_nitrous_oxide_in_air"],
            "o2vmr": st["mole_fraction_of_oxygen_in_air"],
            # _rrtmg_sw.pyx:233-252: asdir <- direct_sw, asdif <- diffuse_sw, aldir <- direct_nir, aldif <- diffuse_nir
            "asdir": st["surface_albedo_for_direct_shortwave"], "asdif": st["surface_albedo_for_diffuse_shortwave"],
            "aldir": st["surface_albedo_for_direct_near_infrared"], "aldif": st["surface_albedo_for_diffuse_near_infrared"],
            "coszen": np.cos(st["zenith_angle"]),
            "cldfr": st["cloud_area_fraction_in_atmosphere_layer"],
            "taucld": st["shortwave_optical_thickness_due_to_cloud"],
            "ssacld": st["single_scattering_albedo_due_to_cloud"], "asmcld": st["cloud_asymmetry_parameter"],
            "fsfcld": st["cloud_forward_scattering_fraction"],
            "cicewp": st["mass_content_of_cloud_ice_in_atmosphere_layer"],
            "cliqwp": st["mass_content_of_cloud_liquid_water_in_atmosphere_layer"],
            "reice": st["cloud_ice_particle_size"], "reliq": st["cloud_water_droplet_radius"],
            "tauaer": st["shortwave_optical_thickness_due_to_aerosol"],
            "ssaaer": st["single_scattering_albedo_due_to_aerosol"], "asmaer": st["aerosol_asymmetry_parameter"],
            "ecaer": st["aerosol_optical_depth_at_55_micron"],
        }
        out = {
            "uflx": diagnostics["upwelling_shortwave_flux_in_air"],
            "dflx": diagnostics["downwelling_shortwave_flux_in_air"],
            "hr": tendencies["air_temperature"],
            "uflxc": diagnostics["upwelling_shortwave_flux_in_air_assuming_clear_sky"],
            "dflxc": diagnostics["downwelling_shortwave_flux_in_air_assuming_clear_sky"],
            "hrc": diagnostics["air_temperature_tendency_from_shortwave_assuming_clear_sky"],
        }
        if self._mcica:
            # same seed draw as the reference (sw/component.py:535-545); tables are NOT re-initialised per call
            if self._random_number_generator == 0:
                self._permute_seed = np.random.randint(0, 1024)
            else:
                self._permute_seed = np.random.randint(0, 2 ** 31 - 1)
            self._engine.set_mcica(True, self._random_number_generator, self._permute_seed)
        self._engine.run_host(n_columns, n_layers, arrays, out,
                              adjes=float(np.asarray(st["flux_adjustment_for_earth_sun_distance"]).item()),
                              dyofyr=day_of_year, solcycfrac=float(np.asarray(st["solar_cycle_fraction"]).item()))
        diagnostics["air_temperature_tendency_from_shortwave"][:] = tendencies["air_temperature"]
        return tendencies, diagnostics

    _ABI_FROM_STATE = {
        "play": "air_pressure", "plev": "air_pressure_on_interface_levels", "tlay": "air_temperature",
        "tsfc": "surface_temperature", "o3vmr": "mole_fraction_of_ozone_in_air",
        "co2vmr": "mole_fraction_of_carbon_dioxide_in_air", "ch4vmr": "mole_fraction_of_methane_in_air",
        "n2ovmr": "mole_fraction_of_nitrous_oxide_in_air", "o2vmr": "mole_fraction_of_oxygen_in_air",
        "asdir": "surface_albedo_for_direct_shortwave", "asdif": "surface_albedo_for_diffuse_shortwave",
        "aldir": "surface_albedo_for_direct_near_infrared", "aldif": "surface_albedo_for_diffuse_near_infrared",
        "cldfr": "cloud_area_fraction_in_atmosphere_layer", "taucld": "shortwave_optical_thickness_due_to_cloud",
        "ssacld": "single_scattering_albedo_due_to_cloud", "asmcld": "cloud_asymmetry_parameter",
        "fsfcld": "cloud_forward_scattering_fraction", "cicewp": "mass_content_of_cloud_ice_in_atmosphere_layer",
        "cliqwp": "mass_content_of_cloud_liquid_water_in_atmosphere_layer", "reice": "cloud_ice_particle_size",
        "reliq": "cloud_water_droplet_radius", "tauaer": "shortwave_optical_thickness_due_to_aerosol",
        "ssaaer": "single_scattering_albedo_due_to_aerosol", "asmaer": "aerosol_asymmetry_parameter",
        "ecaer": "aerosol_optical_depth_at_55_micron",
    }

    def _array_call_device(self, state):
        """State of torch CUDA tensors (this component's units, (levels, columns) layout): zero copy, torch CUDA outputs.
        Scalars (time / day_of_year, flux_adjustment_for_earth_sun_distance, solar_cycle_fraction) stay host values."""
        import torch
        st = {k: (device_state.dense(v) if isinstance(v, torch.Tensor) and v.is_cuda else v) for k, v in state.items()}
        n_layers, n_columns = st["air_temperature"].shape
        Q, Tint, coszen = device_state.marshal(st["specific_humidity"], st["air_temperature"], st["surface_temperature"],
                                               st["air_pressure"], st["air_pressure_on_interface_levels"],
                                               zenith=st["zenith_angle"])
        if self._ignore_day_of_year:
            day_of_year = 0
        else:
            t = st.get("time")
            day_of_year = t.timetuple().tm_yday if t is not None else int(st.get("day_of_year", 1))
        tensors = {k: st[v] for k, v in self._ABI_FROM_STATE.items()}
        tensors.update(h2ovmr=Q, tlev=Tint, coszen=coszen)
        dev = st["air_temperature"].device
        new = lambda nlev: torch.empty((nlev, n_columns), dtype=torch.float64, device=dev)  # noqa: E731
        out = {"uflx": new(n_layers + 1), "dflx": new(n_layers + 1), "hr": new(n_layers), "uflxc": new(n_layers + 1),
               "dflxc": new(n_layers + 1), "hrc": new(n_layers)}
        if self._mcica:
            self._permute_seed = np.random.randint(0, 1024) if self._random_number_generator == 0 else np.random.randint(0, 2 ** 31 - 1)
            self._engine.set_mcica(True, self._random_number_generator, self._permute_seed)

        def scalar(v):
            return float(v.reshape(-1)[0].item()) if isinstance(v, torch.Tensor) else float(np.asarray(v).reshape(-1)[0])
        self._engine.run_device(n_columns, n_layers, tensors, out, adjes=scalar(st["flux_adjustment_for_earth_sun_distance"]),
                                dyofyr=day_of_year, solcycfrac=scalar(st["solar_cycle_fraction"]))
        device_state.finish(self._engine, self._asynchronous)
        tendencies = {"air_temperature": out["hr"]}
        diagnostics = {
            "upwelling_shortwave_flux_in_air": out["uflx"], "downwelling_shortwave_flux_in_air": out["dflx"],
            "upwelling_shortwave_flux_in_air_assuming_clear_sky": out["uflxc"],
            "downwelling_shortwave_flux_in_air_assuming_clear_sky": out["dflxc"],
            "air_temperature_tendency_from_shortwave_assuming_clear_sky": out["hrc"],
            "air_temperature_tendency_from_shortwave": out["hr"],
        }
        return tendencies, diagnostics
