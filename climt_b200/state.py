"""Host-side state helpers: the inputs behind every golden output.

Restates (numpy, no sympl) the parts of climt that define the hot path's inputs:
  * get_hybrid_sigma_pressure_levels / pressure from a,b  (climt/_core/initialization.py:574-727)
  * default values of every RRTMG / Gray input            (initialization.py:139-233, 740-1030)
  * init_ozone                                            (initialization.py:1130-1141)
  * get_interface_values, mass_to_volume_mixing_ratio     (climt/_core/util.py:47-142)
Raw arrays use the layout sympl hands to `array_call`: (levels, columns), columns flattened
from (lat, lon), SI units unless the component's input_properties ask otherwise.
"""
import os

import numpy as np

from .constants import get_constant

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def get_exponent_for_sigma(b_half, num_sigma_levels):
    r_p, S = 2.2, 5
    r_sigma = 1 if num_sigma_levels > 0 else 1.35
    return r_p + (r_sigma - r_p) * np.arctan(S * b_half) / np.arctan(S)


def hybrid_sigma_pressure_levels(num_levels, reference_pressure, model_top_pressure,
                                 proportion_isobaric_levels=0.25, proportion_sigma_levels=0.1):
    """a_k, b_k on interface levels, surface first (initialization.py:625-727)."""
    thickness_dist = np.sin(np.linspace(0.1, np.pi - 0.1, num_levels - 1))
    thickness_dist /= np.sum(thickness_dist)
    thickness_dist *= reference_pressure - model_top_pressure
    pressure_levels = np.zeros(num_levels)
    pressure_levels[0] = model_top_pressure
    pressure_levels[1:] = model_top_pressure + np.cumsum(thickness_dist)
    sigma_interface = (pressure_levels - model_top_pressure) / (reference_pressure - model_top_pressure)
    ak = np.zeros(num_levels)
    bk = np.zeros(num_levels)
    num_isobaric_levels = int(proportion_isobaric_levels * num_levels)
    num_sigma_levels = int(proportion_sigma_levels * num_levels)
    ak[0:num_isobaric_levels] = pressure_levels[0:num_isobaric_levels]
    isobaric_sigma_level = sigma_interface[num_isobaric_levels - 1]
    for level in range(num_isobaric_levels, num_levels - num_sigma_levels):
        sigma_value = sigma_interface[level]
        b_level = (sigma_value - isobaric_sigma_level) / (1 - isobaric_sigma_level)
        r_level = get_exponent_for_sigma(b_level, num_sigma_levels)
        bk[level] = b_level ** r_level
        ak[level] = model_top_pressure + (sigma_value - bk[level]) * (reference_pressure - model_top_pressure)
    for level in range(num_levels - num_sigma_levels, num_levels):
        sigma_value = sigma_interface[level]
        bk[level] = (sigma_interface[level] - isobaric_sigma_level) / (1 - isobaric_sigma_level)
        ak[level] = model_top_pressure + (sigma_value - bk[level]) * (reference_pressure - model_top_pressure)
    return ak[::-1].copy(), bk[::-1].copy()


def pressure_from_hybrid(ak, bk, p_surf):
    """(p_mid, p_interface) in Pa, shape (nz, ncol)/(nz+1, ncol) (initialization.py:598-622)."""
    p_surf = np.atleast_1d(np.asarray(p_surf, dtype=np.float64))
    model_top_pressure = get_constant("top_of_model_pressure")
    p_interface = ak[:, None] + bk[:, None] * (p_surf[None, :] - model_top_pressure)
    delta_p = p_interface[1:, :] - p_interface[:-1, :]
    rk = get_constant("gas_constant_of_dry_air") / get_constant("heat_capacity_of_dry_air_at_constant_pressure")
    p = ((p_interface[1:, :] ** (rk + 1) - p_interface[:-1, :] ** (rk + 1)) / ((rk + 1) * delta_p)) ** (1.0 / rk)
    return p, p_interface


def init_ozone(p):
    """Default ozone profile: not-a-knot cubic spline through climt's 30-point reference
    profile (initialization.py:1130-1141; the goldens were generated with scipy's CubicSpline)."""
    from scipy.interpolate import CubicSpline
    p_ref = 1e5 * np.linspace(0.998, 0.001, 30)
    ozone_ref = np.load(os.path.join(_DATA, "ozone_profile.npy"))
    return CubicSpline(p_ref[::-1], ozone_ref[::-1])(p)


def get_interface_values(mid_level_values, surface_value, mid_level_pressure, interface_level_pressure):
    """ln-p weighted interface values (util.py:89-142)."""
    interface_values = np.zeros((mid_level_values.shape[0] + 1, mid_level_values.shape[1]), dtype=np.double)
    log_mid_p = np.log(mid_level_pressure)
    interp_weight = (np.log(interface_level_pressure[1:-1, :]) - log_mid_p[1:, :]) / (
        log_mid_p[:-1, :] - log_mid_p[1::, :])
    interface_values[1:-1, :] = mid_level_values[1:, :] - interp_weight * (
        mid_level_values[1:, :] - mid_level_values[0:-1, :])
    interface_values[0, :] = surface_value[:]
    interface_values[-1, :] = mid_level_values[-1, :]
    return interface_values


def mass_to_volume_mixing_ratio(mass_mixing_ratio, molecular_weight, molecular_weight_air=28.964):
    return mass_mixing_ratio * molecular_weight_air / molecular_weight


def default_grid(nz, ncol=1, p_surf=None):
    """Raw (nz, ncol) mid / (nz+1, ncol) interface pressures in Pa, as climt.get_grid builds them."""
    p_surf = get_constant("reference_air_pressure") if p_surf is None else p_surf
    ak, bk = hybrid_sigma_pressure_levels(nz + 1, get_constant("reference_air_pressure"),
                                          get_constant("top_of_model_pressure"))
    ps = np.full(ncol, p_surf, dtype=np.float64) if np.isscalar(p_surf) else np.asarray(p_surf, dtype=np.float64)
    p, p_int = pressure_from_hybrid(ak, bk, ps)
    return {"air_pressure": p, "air_pressure_on_interface_levels": p_int, "surface_air_pressure": ps}


def default_rrtmg_lw_state(nz, ncol=1, p_surf=None):
    """Raw default state of RRTMGLongwave (SI units: pressures in Pa, cloud paths in kg m^-2)."""
    g = default_grid(nz, ncol, p_surf)
    shp = (nz, ncol)
    st = dict(g)
    st.update({
        "air_temperature": np.full(shp, 290.0),
        "surface_temperature": np.full(ncol, 300.0),
        "specific_humidity": np.zeros(shp),
        "mole_fraction_of_ozone_in_air": init_ozone(g["air_pressure"]),
        "mole_fraction_of_carbon_dioxide_in_air": np.full(shp, 330e-6),
        "mole_fraction_of_methane_in_air": np.zeros(shp),
        "mole_fraction_of_nitrous_oxide_in_air": np.zeros(shp),
        "mole_fraction_of_oxygen_in_air": np.full(shp, 0.21),
        "mole_fraction_of_cfc11_in_air": np.zeros(shp),
        "mole_fraction_of_cfc12_in_air": np.zeros(shp),
        "mole_fraction_of_cfc22_in_air": np.zeros(shp),
        "mole_fraction_of_carbon_tetrachloride_in_air": np.zeros(shp),
        "surface_longwave_emissivity": np.ones((16, ncol)),
        "cloud_area_fraction_in_atmosphere_layer": np.zeros(shp),
        "longwave_optical_thickness_due_to_cloud": np.zeros((nz, ncol, 16)),
        "mass_content_of_cloud_ice_in_atmosphere_layer": np.zeros(shp),
        "mass_content_of_cloud_liquid_water_in_atmosphere_layer": np.zeros(shp),
        "cloud_ice_particle_size": np.full(shp, 20.0),
        "cloud_water_droplet_radius": np.full(shp, 10.0),
        "longwave_optical_thickness_due_to_aerosol": np.zeros((16, nz, ncol)),
    })
    return st


def default_rrtmg_sw_state(nz, ncol=1, p_surf=None):
    """Raw default state of RRTMGShortwave (SI units; initialization.py:173-233, 740-1030).
    `time` = 2000-01-01 (get_grid :520) -> day of year 1."""
    g = default_grid(nz, ncol, p_surf)
    shp = (nz, ncol)
    st = dict(g)
    st.update({
        "air_temperature": np.full(shp, 290.0),
        "surface_temperature": np.full(ncol, 300.0),
        "specific_humidity": np.zeros(shp),
        "mole_fraction_of_ozone_in_air": init_ozone(g["air_pressure"]),
        "mole_fraction_of_carbon_dioxide_in_air": np.full(shp, 330e-6),
        "mole_fraction_of_methane_in_air": np.zeros(shp),
        "mole_fraction_of_nitrous_oxide_in_air": np.zeros(shp),
        "mole_fraction_of_oxygen_in_air": np.full(shp, 0.21),
        "mass_content_of_cloud_ice_in_atmosphere_layer": np.zeros(shp),
        "mass_content_of_cloud_liquid_water_in_atmosphere_layer": np.zeros(shp),
        "cloud_ice_particle_size": np.full(shp, 20.0),
        "cloud_water_droplet_radius": np.full(shp, 10.0),
        "cloud_area_fraction_in_atmosphere_layer": np.zeros(shp),
        "zenith_angle": np.zeros(ncol),
        "surface_albedo_for_direct_shortwave": np.full(ncol, 0.06),
        "surface_albedo_for_direct_near_infrared": np.full(ncol, 0.06),
        "surface_albedo_for_diffuse_near_infrared": np.full(ncol, 0.06),
        "surface_albedo_for_diffuse_shortwave": np.full(ncol, 0.06),
        "shortwave_optical_thickness_due_to_cloud": np.zeros((nz, ncol, 14)),
        "cloud_asymmetry_parameter": 0.85 * np.ones((nz, ncol, 14)),
        "cloud_forward_scattering_fraction": 0.8 * np.ones((nz, ncol, 14)),
        "single_scattering_albedo_due_to_cloud": 0.9 * np.ones((nz, ncol, 14)),
        "shortwave_optical_thickness_due_to_aerosol": np.zeros((14, nz, ncol)),
        "aerosol_asymmetry_parameter": np.zeros((14, nz, ncol)),
        "single_scattering_albedo_due_to_aerosol": 0.5 * np.ones((14, nz, ncol)),
        "aerosol_optical_depth_at_55_micron": np.zeros((6, nz, ncol)),
        "solar_cycle_fraction": 0.0,
        "flux_adjustment_for_earth_sun_distance": 1.0,
        "day_of_year": 1,
    })
    return st


def default_gray_state(nz, ncol=1, p_surf=None):
    """Default state of GrayLongwaveRadiation: tau = 1 - p/ps on interfaces (initialization.py:1144-1150)."""
    g = default_grid(nz, ncol, p_surf)
    ps = g["surface_air_pressure"]
    return {"air_temperature": np.full((nz, ncol), 290.0), "surface_temperature": np.full(ncol, 300.0),
            "air_pressure": g["air_pressure"], "air_pressure_on_interface_levels": g["air_pressure_on_interface_levels"],
            "longwave_optical_depth_on_interface_levels": 1.0 * (1.0 - g["air_pressure_on_interface_levels"] / ps[None, :])}
