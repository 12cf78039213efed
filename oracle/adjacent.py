"""CPU restatement of the two column steps either side of the radiation call.  TEST INFRASTRUCTURE ONLY.

  orbit, instellation     climt/_components/instellation/component.py:36-61 (array_call), :84-132 (_instellation_kernel_np),
                          :135-150 (_obliquity_star_jit), :153-177 (_sun_ecliptic_longitude_jit), :180-191 (_gmst_jit)
  slab_surface            climt/_components/slab_surface.py:164-447 (array_call, include_ekman=False), :449-517 (kernel)

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
