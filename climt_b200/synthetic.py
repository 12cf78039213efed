"""Synthetic (lat x lon x level) column states for parity tests and the benchmark.

Follows SURVEY.md section 8(d): hybrid sigma-pressure levels with a random surface pressure per column,
a moist-adiabat-like temperature profile, humidity decaying with height and latitude, climt's
default ozone profile, fixed well-mixed gases; optional random clouds.  Everything is fp64 and in the
units the RRTMG C ABI takes (hPa, K, volume mixing ratios, g m-2, micron).
"""
import numpy as np

from . import state as _s
from .constants import get_constant

LW_FIELDS = ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "cfc11", "cfc12",
             "cfc22", "ccl4", "emis", "cldfr", "taucld", "cicewp", "cliqwp", "reice", "reliq", "tauaer")


def make_lw_state(ncol, nlay, seed=20260925, clouds=False, trace=True, aerosol=False, emis_range=None):
    rng = np.random.default_rng(seed)
    lat = np.deg2rad(rng.uniform(-90, 90, ncol))
    ps = rng.uniform(9.5e4, 1.03e5, ncol)
    ak, bk = _s.hybrid_sigma_pressure_levels(nlay + 1, get_constant("reference_air_pressure"),
                                             get_constant("top_of_model_pressure"))
    p, p_int = _s.pressure_from_hybrid(ak, bk, ps)          # Pa, (nlay, ncol), surface first
    ts = 300.0 - 40.0 * np.sin(lat) ** 2 + rng.normal(0, 2, ncol)
    t = np.maximum(ts[None, :] * (p / ps[None, :]) ** 0.19, 200.0) + rng.normal(0, 0.5, p.shape)
    tsfc = ts + rng.uniform(0, 2, ncol)
    q = 0.018 * (p / ps[None, :]) ** 3 * np.exp(-(np.rad2deg(lat)[None, :] / 30.0) ** 2) + 1e-6
    h2o = _s.mass_to_volume_mixing_ratio(q, 18.02)
    o3 = np.maximum(_s.init_ozone(p), 1e-9)
    shp = (nlay, ncol)
    st = {
        "play": p / 100.0, "plev": p_int / 100.0, "tlay": t,
        "tsfc": tsfc, "h2o": h2o, "o3": o3,
        "co2": np.full(shp, 400e-6), "ch4": np.full(shp, 1.8e-6 if trace else 0.0),
        "n2o": np.full(shp, 3.2e-7 if trace else 0.0), "o2": np.full(shp, 0.21),
        "cfc11": np.full(shp, 2.3e-10 if trace else 0.0), "cfc12": np.full(shp, 5.2e-10 if trace else 0.0),
        "cfc22": np.full(shp, 2.3e-10 if trace else 0.0), "ccl4": np.full(shp, 8.0e-11 if trace else 0.0),
        "emis": np.ones((16, ncol)) if emis_range is None else rng.uniform(*emis_range, (16, ncol)),
        "cldfr": np.zeros(shp), "taucld": np.zeros((nlay, ncol, 16)),
        "cicewp": np.zeros(shp), "cliqwp": np.zeros(shp),
        "reice": np.full(shp, 20.0), "reliq": np.full(shp, 10.0),
        "tauaer": np.zeros((16, nlay, ncol)) if not aerosol else rng.uniform(0, 0.05, (16, nlay, ncol)),
    }
    st["tlev"] = _s.get_interface_values(st["tlay"], st["tsfc"], st["play"], st["plev"])
    if clouds:
        band = (p > 3.0e4) & (p < 9.0e4) & (rng.uniform(0, 1, shp) < 0.3)
        st["cldfr"] = np.where(band, rng.uniform(0.05, 1, shp), 0.0)
        st["cicewp"] = np.where(band, rng.uniform(0, 30, shp), 0.0)
        st["cliqwp"] = np.where(band, rng.uniform(0, 30, shp), 0.0)
        st["reice"] = rng.uniform(15, 120, shp)
        st["reliq"] = rng.uniform(4, 30, shp)
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}

SW_FIELDS = ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "asdir", "asdif", "aldir",
             "aldif", "coszen", "cldfr", "taucld", "ssacld", "asmcld", "fsfcld", "cicewp", "cliqwp", "reice", "reliq",
             "tauaer", "ssaaer", "asmaer", "ecaer")


def make_sw_state(ncol, nlay, seed=20260925, clouds=False, trace=True, aerosol=False, ecmwf=False, overcast_only=True):
    """Shortwave twin of make_lw_state (same atmosphere for the same seed).  Non-McICA RRTMG-SW only accepts cloud
    fractions 0 or 1 (rrtmg_sw_rad.nomcica.f90:616-620) -> overcast_only."""
    lw = make_lw_state(ncol, nlay, seed=seed, clouds=clouds, trace=trace)
    rng = np.random.default_rng(seed + 1)
    shp = (nlay, ncol)
    st = {k: lw[k] for k in ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "cldfr",
                             "cicewp", "cliqwp", "reice", "reliq")}
    if clouds and overcast_only:
        st["cldfr"] = np.where(st["cldfr"] > 0, 1.0, 0.0)
    zen = np.deg2rad(np.clip(rng.uniform(0, 85, ncol), 0, 85))
    st.update({
        "asdir": rng.uniform(0.06, 0.3, ncol), "asdif": rng.uniform(0.06, 0.3, ncol),
        "aldir": rng.uniform(0.06, 0.3, ncol), "aldif": rng.uniform(0.06, 0.3, ncol),
        "coszen": np.cos(zen),
        "taucld": np.zeros((nlay, ncol, 14)), "ssacld": 0.9 * np.ones((nlay, ncol, 14)),
        "asmcld": 0.85 * np.ones((nlay, ncol, 14)), "fsfcld": 0.8 * np.ones((nlay, ncol, 14)),
        "tauaer": np.zeros((14, nlay, ncol)), "ssaaer": 0.5 * np.ones((14, nlay, ncol)),
        "asmaer": np.zeros((14, nlay, ncol)), "ecaer": np.zeros((6, nlay, ncol)),
    })
    if aerosol:
        st["tauaer"] = rng.uniform(0, 0.05, (14, nlay, ncol))
        st["ssaaer"] = rng.uniform(0.7, 0.99, (14, nlay, ncol))
        st["asmaer"] = rng.uniform(0.5, 0.8, (14, nlay, ncol))
    if ecmwf:
        st["ecaer"] = rng.uniform(0, 0.02, (6, nlay, ncol))
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}


def make_emanuel_state(ncol, nlev, seed=20260925):
    """Soundings for the Emanuel convection path, in the component's layout and units: (ncol, nlev) arrays, level 0 at the
    surface, pressures in mbar.  Same grid, surface pressure and latitude law as make_lw_state; conditionally unstable
    lapse rates, a moist boundary layer whose relative humidity varies from column to column (dry columns, columns that
    only just trigger, deep convection), sheared winds, and a cloud-base mass flux carried over from a previous step in
    about half of the columns."""
    rng = np.random.default_rng(seed + 3)
    lat = np.deg2rad(rng.uniform(-90, 90, ncol))
    ps = rng.uniform(9.5e4, 1.03e5, ncol)
    ak, bk = _s.hybrid_sigma_pressure_levels(nlev + 1, get_constant("reference_air_pressure"),
                                             get_constant("top_of_model_pressure"))
    p, p_int = _s.pressure_from_hybrid(ak, bk, ps)          # Pa, (nlev, ncol), surface first
    ts = 302.0 - 45.0 * np.sin(lat) ** 2 + rng.normal(0, 2, ncol)
    gamma = rng.uniform(0.17, 0.215, ncol)                   # R*lapse/g: 5.8 .. 7.3 K/km
    t = np.maximum(ts[None, :] * (p / ps[None, :]) ** gamma[None, :], rng.uniform(190.0, 215.0, ncol)[None, :])
    t = t + rng.normal(0, 0.3, p.shape)
    es = 611.2 * np.exp(17.67 * (t - 273.15) / (t - 29.65))
    qsat = 0.622 * es / (p - 0.378 * es)
    rh_s = rng.uniform(0.2, 0.98, ncol)
    sig = p / ps[None, :]
    rh = np.clip(rh_s[None, :] * (0.25 + 0.75 * sig ** 1.5) * rng.uniform(0.85, 1.1, p.shape), 0.02, 0.99)
    q = np.maximum(rh * np.minimum(qsat, 0.04), 1e-7)
    u = 5.0 + 25.0 * (1.0 - sig) * np.cos(lat)[None, :] + rng.normal(0, 1.0, p.shape)
    v = rng.normal(0, 2.0, p.shape) + 3.0 * (1.0 - sig)
    cbmf = np.where(rng.uniform(size=ncol) < 0.5, rng.uniform(0.0, 0.03, ncol), 0.0)
    st = {"air_temperature": t.T, "specific_humidity": q.T, "eastward_wind": u.T, "northward_wind": v.T,
          "air_pressure": p.T / 100.0, "air_pressure_on_interface_levels": p_int.T / 100.0, "cloud_base_mass_flux": cbmf}
    return {k: np.ascontiguousarray(a, dtype=np.float64) for k, a in st.items()}


def make_surface_state(ncol, seed=20260925, aquaplanet=True):
    """Latitude / longitude of the columns and the slab-surface fields (SlabSurface's inputs other than the radiative fluxes):
    an aquaplanet (every column open sea, 50 m mixed layer -- BASELINE.json configs[4]) or a random mix of the four area types."""
    rng = np.random.default_rng(seed + 5)
    st = {
        "latitude": rng.uniform(-90.0, 90.0, ncol), "longitude": rng.uniform(0.0, 360.0, ncol),
        "surface_upward_latent_heat_flux": rng.uniform(0.0, 250.0, ncol),
        "surface_upward_sensible_heat_flux": rng.uniform(-20.0, 80.0, ncol),
        "surface_thermal_capacity": np.full(ncol, 4.1813e3), "surface_material_density": np.full(ncol, 1.0e3),
        "upward_heat_flux_at_ground_level_in_soil": np.zeros(ncol), "heat_flux_into_sea_water_due_to_sea_ice": np.zeros(ncol),
        "area_type": np.full(ncol, 2, dtype=np.int32) if aquaplanet else rng.integers(0, 4, ncol).astype(np.int32),
        "soil_layer_thickness": np.full(ncol, 50.0), "ocean_mixed_layer_thickness": np.full(ncol, 50.0),
        "heat_capacity_of_soil": np.full(ncol, 2000.0), "sea_water_density": np.full(ncol, 1.029e3),
        "ocean_heat_transport_convergence": rng.uniform(-30.0, 30.0, ncol),
    }
    return {k: np.ascontiguousarray(a) for k, a in st.items()}


def component_states(ncol, nlay, seed=20260925, clouds=False):
    """The same synthetic atmosphere as make_lw_state / make_sw_state, as the raw state dicts sympl hands to
    RRTMGLongwave.array_call / RRTMGShortwave.array_call (the components' own quantity names and units; pressures in mbar)."""
    from .rrtmg_lw import RRTMGLongwave
    from .rrtmg_sw import RRTMGShortwave
    lw, sw = make_lw_state(ncol, nlay, seed=seed, clouds=clouds), make_sw_state(ncol, nlay, seed=seed, clouds=clouds)
    short = {"h2ovmr": "h2o", "o3vmr": "o3", "co2vmr": "co2", "ch4vmr": "ch4", "n2ovmr": "n2o", "o2vmr": "o2", "cfc11vmr": "cfc11",
             "cfc12vmr": "cfc12", "cfc22vmr": "cfc22", "ccl4vmr": "ccl4"}
    out = []
    for cls, st in ((RRTMGLongwave, lw), (RRTMGShortwave, sw)):
        raw = {name: st[short.get(abi, abi)] for abi, name in cls._ABI_FROM_STATE.items()}
        raw["specific_humidity"] = st["h2o"] * 18.02 / 28.964          # inverse of mass_to_volume_mixing_ratio (util.py:84)
        out.append(raw)
    out[1]["zenith_angle"] = np.arccos(sw["coszen"])
    out[1]["solar_cycle_fraction"] = np.array(0.0)
    out[1]["flux_adjustment_for_earth_sun_distance"] = np.array(1.0)
    out[1]["day_of_year"] = 1
    return out[0], out[1]
