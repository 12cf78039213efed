"""CPU restatement of the two column steps either side of the radiation call.  TEST INFRASTRUCTURE ONLY.

  orbit, instellation     climt/_components/instellation/component.py:36-61 (array_call), :84-132 (_instellation_kernel_np),
                          :135-150 (_obliquity_star_jit), :153-177 (_sun_ecliptic_longitude_jit), :180-191 (_gmst_jit)
  slab_surface            climt/_components/slab_surface.py:164-447 (array_call, include_ekman=False), :449-517 (kernel)
  berger                  climt/_components/berger_solar_insolation.py:538-693 (array_call, _driver, the Berger 1978 series in
                          _get_orbital_parameters_functional, the numba kernel _get_solar_parameters_np); coefficients from the
                          data file climt_b200/data/berger1978.npz (tools/extract_berger_tables.py)

Pinned by tests/golden/adjacent_reference.npz, which tests/golden/make_adjacent_golden.py produced by running the reference's
own component classes (tests/test_adjacent_cpu.py).
"""
import datetime

import numpy as np

AREA_MAP = {"land": 0, "land_ice": 1, "sea": 2, "sea_ice": 3}


def julian_centuries(model_time):
    d = model_time - datetime.datetime(2000, 1, 1, 12, 0)
    return (d.days + (d.seconds + d.microseconds / 1000000.0) / (24 * 3600.0)) / 36525.0


def orbit(t):
    """-> (sin_dec, cos_dec, right_ascension, gmst)"""
    eps = np.deg2rad(23.0 + 26.0 / 60 + 21.406 / 3600.0
                     - (46.836769 * t - 0.0001831 * (t ** 2) + 0.00200340 * (t ** 3) - 0.576e-6 * (t ** 4) - 4.34e-8 * (t ** 5)) / 3600.0)
    mean_anomaly = np.deg2rad(357.52910 + 35999.05030 * t - 0.0001559 * t * t - 0.00000048 * t * t * t)
    mean_longitude = np.deg2rad(280.46645 + 36000.76983 * t + 0.0003032 * (t ** 2))
    d_l = np.deg2rad((1.914600 - 0.004817 * t - 0.000014 * (t ** 2)) * np.sin(mean_anomaly)
                     + (0.019993 - 0.000101 * t) * np.sin(2 * mean_anomaly) + 0.000290 * np.sin(3 * mean_anomaly))
    eclon = mean_longitude + d_l
    x, y, z = np.cos(eclon), np.cos(eps) * np.sin(eclon), np.sin(eps) * np.sin(eclon)
    r = np.sqrt(1.0 - z * z)
    dec = np.arctan2(z, r)
    ra = 2.0 * np.arctan2(y, x + r)
    theta = 67310.54841 + t * (876600 * 3600 + 8640184.812866 + t * (0.093104 - t * 6.2 * 10e-6))
    gmst = np.deg2rad(theta / 240.0) % (2.0 * np.pi)
    if gmst < 0:
        gmst += 2.0 * np.pi
    return float(np.sin(dec)), float(np.cos(dec)), float(ra), float(gmst)


def instellation(lat_deg, lon_deg, model_time):
    lat_deg, lon_deg = np.asarray(lat_deg, dtype=np.float64), np.asarray(lon_deg, dtype=np.float64)
    sin_dec, cos_dec, ra, gmst = orbit(julian_centuries(model_time))
    lat = np.deg2rad(lat_deg)
    h = gmst + lon_deg * (np.pi / 180.0) - ra
    cos_mu = np.clip(np.sin(lat) * sin_dec + np.cos(lat) * cos_dec * np.cos(h), -1.0, 1.0)
    return np.minimum(np.arccos(cos_mu), np.pi / 2.0)


def slab_surface(state):
    """state keyed by the component's input names -> (tendency, depth, ocean_heat_transport_convergence), area_type's shape"""
    at = np.asarray(state["area_type"])
    if at.dtype.kind in "iu":
        code = at.astype(np.int32)
    else:
        code = np.zeros(at.shape, dtype=np.int32)
        for k, v in AREA_MAP.items():
            code[at.astype(str) == k] = v
    shape = code.shape
    code = code.reshape(-1)

    def surf(name):
        a = np.asarray(state[name], dtype=np.float64)
        return (a[..., 0] if a.ndim > 1 else a).reshape(-1)

    def vec(name):
        return np.asarray(state[name], dtype=np.float64).reshape(-1)
    net = (surf("downwelling_shortwave_flux_in_air") + surf("downwelling_longwave_flux_in_air") - surf("upwelling_shortwave_flux_in_air")
           - surf("upwelling_longwave_flux_in_air") - vec("surface_upward_sensible_heat_flux") - vec("surface_upward_latent_heat_flux"))
    land, sea, land_ice, sea_ice = (code == 0) | (code == 1), (code == 2) | (code == 3), code == 1, code == 3
    net = np.where(land_ice, -vec("upward_heat_flux_at_ground_level_in_soil"), np.where(sea_ice, vec("heat_flux_into_sea_water_due_to_sea_ice"), net))
    oht = vec("ocean_heat_transport_convergence")
    net = np.where(sea & ~sea_ice, net + oht, net)
    dens = np.where(sea, vec("sea_water_density"), vec("surface_material_density"))
    d = np.where(sea, vec("ocean_mixed_layer_thickness"), np.where(land, vec("soil_layer_thickness"), 0.0))
    cap = np.where(land, vec("heat_capacity_of_soil"), vec("surface_thermal_capacity"))
    hc = (dens * d) * cap
    with np.errstate(divide="ignore", invalid="ignore"):
        val = np.where(hc != 0, net / np.where(hc != 0, hc, 1.0), 0.0)
    val = np.where(land_ice | sea_ice, 0.0, val)
    return val.reshape(shape), d.reshape(shape), oht.reshape(shape)


def berger_orbit(t, tables):
    """_get_orbital_parameters_functional (:579-625): t = years since 1950 -> lambda_m0, eccentricity, omega_tilde, obliquity"""
    a2d = float(tables["arcsec_to_degree"])
    obliquity = 23.320556 + np.sum(tables["A"] * a2d * np.cos((tables["f"] * a2d * t + tables["delta"]) * np.pi / 180.0))
    obliquity = obliquity * np.pi / 180.0
    cos_sum = np.sum(tables["P"] * np.cos(tables["alpha"] * a2d * t + tables["zeta"]))
    sin_sum = np.sum(tables["P"] * np.sin(tables["alpha"] * a2d * t + tables["zeta"]))
    e2 = cos_sum * cos_sum + sin_sum * sin_sum
    e = np.sqrt(e2)
    e3 = e * e2
    pi_val = np.arctan2(sin_sum, cos_sum)
    if pi_val < 0:
        pi_val += 2.0 * np.pi
    omega = pi_val * 180.0 / np.pi + 50.439273 * a2d * t + 3.392506
    omega += np.sum(tables["F"] * np.sin((tables["f_prime"] * a2d * t + tables["delta_prime"]) * np.pi / 180.0))
    omega = (omega % 360.0) * np.pi / 180.0
    beta = np.sqrt(1.0 - e2)
    lambda_m0 = 2.0 * ((0.5 * e + 0.125 * e3) * (1.0 + beta) * np.sin(omega + np.pi) - 0.25 * e2 * (0.5 + beta) * np.sin(2 * (omega + np.pi))
                       + 0.125 * e3 * (1.0 / 3.0 + beta) * np.sin(3 * (omega + np.pi)))
    return lambda_m0, e, omega, obliquity


def berger(lat, lon, time, solar_constant, tables):
    """BergerSolarInsolation.array_call: -> dict of the five diagnostics (lat / lon as given: the reference does not convert them)"""
    lat, lon = np.asarray(lat, dtype=np.float64), np.asarray(lon, dtype=np.float64)
    lambda_m0, e, omega, obliquity = berger_orbit(float(time.year - 1950), tables)
    y0, y1 = type(time)(time.year, 3, 20, 12), type(time)(time.year + 1, 3, 20, 12)
    ysve = (time - y0).total_seconds() / (y1 - y0).total_seconds()
    fday = (time - type(time)(time.year, time.month, time.day)).total_seconds() / 86400.0
    lambda_m = lambda_m0 + ysve * 2.0 * np.pi
    temp = lambda_m - (omega + np.pi)
    st = np.sin(temp)
    lmbda = lambda_m + e * (2.0 * st + e * (1.25 * np.sin(2 * temp) + e * ((13.0 / 12.0) * np.sin(3 * temp) - 0.25 * st)))
    inv_rho = (1 + e * np.cos(lmbda - (omega + np.pi))) / (1 - e * e)
    decl = np.arcsin(np.sin(obliquity) * np.sin(lmbda))
    H = 2 * np.pi * (fday + lon / 360.0)
    cos_mu = np.sin(lat) * np.sin(decl) - np.cos(lat) * np.cos(decl) * np.cos(H)
    return {"solar_insolation": solar_constant * (inv_rho * inv_rho) * cos_mu, "solar_zenith_angle": np.arccos(cos_mu),
            "obliquity": obliquity, "eccentricity": e, "normalized_earth_sun_distance": 1.0 / inv_rho}
