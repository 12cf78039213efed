"""ctypes front-end of the C++ CORK restatement (oracle/cork_oracle.cpp) plus the component glue.  TEST INFRASTRUCTURE ONLY.

`lw_call` / `sw_call` follow CorkLongwaveRadiation.array_call (cork/lw/component.py:208-373) and
CorkShortwaveRadiation.array_call (cork/sw/component.py:223-496) for optics="correlated_k" with additive overlap,
including the host-side preparation of the interpolation coordinates (`_additive_co2_fast`, correlated_k.py:526-561).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)

MOLAR_MASS_DRY_AIR = 28.970   # cork/common.py:9-17
MOLAR_MASS_H2O = 18.015
MOLAR_MASS = {"h2o": 18.015, "co2": 44.010, "o3": 47.998, "ch4": 16.043, "n2o": 44.013, "o2": 31.998}


def build(force=False):
    so = os.path.join(_HERE, "liborc_cork.so")
    src = os.path.join(_HERE, "cork_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liborc_cork.so"])
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


def load_table(path):
    """.npz k-table -> dict of arrays (load_k_table / _load_npz_table, correlated_k.py:159-218)."""
    with np.load(path, allow_pickle=True) as z:
        return {k: z[k] for k in z.files}


def column_amount(q, p_int, g):
    nlev, ncol = q.shape
    out = np.zeros((nlev, ncol))
    lib().orc_cork_column_amount(_p(_c(q)), _p(_c(p_int)), ctypes.c_double(g), nlev, ncol, _p(out))
    return out


def heating_rate(net, p_int, g, cpd):
    nlev, ncol = net.shape[0] - 1, net.shape[1]
    out = np.zeros((nlev, ncol))
    lib().orc_cork_heating(_p(_c(net)), _p(_c(p_int)), ctypes.c_double(g), ctypes.c_double(cpd), nlev, ncol, _p(out))
    return out


def optical_depth(table, T, p, gas_amounts, h2o_vmr=None, co2_vmr=None, co2_logk=True):
    """compute_ck_optical_depth for additive overlap (correlated_k.py:378-470, 526-561)."""
    k = np.ascontiguousarray(table["k_coefficients"])
    if k.dtype not in (np.float32, np.float64):
        k = k.astype(np.float64)
    ngas, nband, ngpt, nT, nP = k.shape[:5]
    nX = k.shape[5] if k.ndim >= 6 else 0
    nC = k.shape[6] if k.ndim == 7 else 0
    nlev, ncol = T.shape
    T_grid = _c(table["temperature_grid"])
    p_grid_log = _c(table["pressure_grid_log"])
    log_p = np.log(np.maximum(p, 1.0))
    zero = np.zeros(1)
    log_x_grid = log_c_grid = zero
    log_x = log_c = zero
    if nX:
        x_grid = np.asarray(table["h2o_vmr_grid"], dtype=np.float64)
        log_x_grid = np.log(np.maximum(x_grid, 1e-30))
        log_x = np.log(np.maximum(np.clip(h2o_vmr, float(x_grid[0]), float(x_grid[-1])), 1e-30))
    if nC:
        c_grid = np.asarray(table["co2_vmr_grid"], dtype=np.float64)
        log_c_grid = np.log(np.maximum(c_grid, 1e-30))
        log_c = np.log(np.maximum(np.clip(co2_vmr, float(c_grid[0]), float(c_grid[-1])), 1e-30))
    has_cont = "continuum_kappa" in table and np.asarray(table["continuum_kappa"]).ndim == 4 and nX > 0
    log_cont = np.log(np.maximum(np.asarray(table["continuum_kappa"], dtype=np.float64), 1e-40)) if has_cont else zero
    tau = np.zeros((nband, ngpt, nlev, ncol))
    ga = _c(gas_amounts)
    lib().orc_cork_tau(k.ctypes.data_as(ctypes.c_void_p), int(k.dtype == np.float64), ngas, nband, ngpt, nT, nP, nX, nC, _p(T_grid), _p(p_grid_log), _p(_c(log_x_grid)),
                       _p(_c(log_c_grid)), _p(_c(T)), _p(_c(log_p)), _p(_c(log_x)), _p(_c(log_c)), _p(ga), int(has_cont),
                       _p(_c(log_cont)), int(co2_logk), nlev, ncol, _p(tau))
    return tau


def planck_sources(table, T, T_surf, sigma, nband, ngpt):
    pf = np.ascontiguousarray(table["planck_fraction"], dtype=np.float64)
    nlev, ncol = T.shape
    planck_src = np.zeros((nband, ngpt, nlev, ncol))
    surf_src = np.zeros((nband, ngpt, ncol))
    lib().orc_cork_planck(pf.ctypes.data_as(_dp), pf.shape[0], pf.shape[1], pf.shape[2], _p(_c(table["temperature_grid"])),
                          _p(_c(T)), _p(_c(T_surf)), ctypes.c_double(sigma), nband, ngpt, 0, nlev, ncol, _p(planck_src), _p(surf_src))
    return planck_src, surf_src


def lw_transport(tau, planck_src, surf_src, emissivity, weights, D):
    nband, ngpt, nlev, ncol = tau.shape
    ub, db = np.zeros((nband, nlev + 1, ncol)), np.zeros((nband, nlev + 1, ncol))
    u, d = np.zeros((nlev + 1, ncol)), np.zeros((nlev + 1, ncol))
    lib().orc_cork_lw_transport(_p(_c(tau)), _p(_c(planck_src)), _p(_c(surf_src)), _p(_c(emissivity)), _p(_c(weights)), nband,
                                ngpt, nlev, ncol, ctypes.c_double(D), _p(ub), _p(db), _p(u), _p(d))
    return ub, db, u, d


def sw_two_stream(tau, ssa, asym, zenith, albedo, solar_flux, weights):
    nband, ngpt, nlev, ncol = tau.shape
    ub, db = np.zeros((nband, nlev + 1, ncol)), np.zeros((nband, nlev + 1, ncol))
    u, d = np.zeros((nlev + 1, ncol)), np.zeros((nlev + 1, ncol))
    lib().orc_cork_sw_two_stream(_p(_c(tau)), _p(_c(ssa)), _p(_c(asym)), _p(_c(zenith)), _p(_c(albedo)), _p(_c(solar_flux)),
                                 _p(_c(weights)), nband, ngpt, nlev, ncol, _p(ub), _p(db), _p(u), _p(d))
    return ub, db, u, d


def _band_diag(tau, weights, up_band, down_band, p_int, g, cpd):
    nband, ngpt, nlev, ncol = tau.shape
    tau_band = np.zeros((nband, nlev, ncol))
    for b in range(nband):
        for gp in range(ngpt):
            tau_band[b] += weights[b, gp] * tau[b, gp]
    hr_band = np.zeros((nband, nlev, ncol))
    for b in range(nband):
        hr_band[b] = heating_rate(up_band[b] - down_band[b], p_int, g, cpd) * 86400.0
    return tau_band, hr_band


def gas_setup(table, s, p_int, g, with_co2):
    """Table classification (cork/lw/component.py:46-59) and the gas part of array_call (:243-287; cork/sw/component.py:308-347):
    -> gas_amounts (ngas, nlev, ncol), h2o_vmr or None, co2_vmr or None.  s["q"] (or s["h2o"]) is the specific humidity, every other
    gas s[<name>] a mole fraction."""
    names = [str(x) for x in table["gas_names"]] if "gas_names" in table else ["effective"]
    has_h2o, has_co2 = "h2o_vmr_grid" in table, "co2_vmr_grid" in table
    fully = names == ["effective"] and not has_h2o
    bg = (names == ["effective"] and has_h2o) or str(np.asarray(table.get("background_is_premixed", ""))).lower() == "true"
    nlev, ncol = p_int.shape[0] - 1, p_int.shape[1]
    gas_amounts = np.zeros((len(names), nlev, ncol))
    h2o_vmr = co2_vmr = None
    q = s["q"] if "q" in s else s.get("h2o")
    if fully:
        gas_amounts[0] = column_amount(np.ones((nlev, ncol)), p_int, g)
    elif bg:
        gas_amounts[0] = column_amount(np.ones((nlev, ncol)), p_int, g)
        M = MOLAR_MASS_H2O / MOLAR_MASS_DRY_AIR
        h2o_vmr = q / np.maximum(q + (1.0 - q) * M, 1e-30)
        if has_co2 and with_co2:
            co2_vmr = s["co2"]
    else:
        for ig, gas in enumerate(names):
            qg = q if gas == "h2o" else s[gas] * (MOLAR_MASS.get(gas, MOLAR_MASS_DRY_AIR) / MOLAR_MASS_DRY_AIR)
            gas_amounts[ig] = column_amount(np.ascontiguousarray(qg), p_int, g)
    return gas_amounts, h2o_vmr, co2_vmr


def lw_call(table, s, g, cpd, sigma, D=1.66):
    """s: T, p, p_int [Pa], T_surf, q (specific humidity), co2 (VMR), emissivity (nband, ncol), tau_cloud_lw (nlev, ncol, nband)."""
    T, p, p_int = s["T"], s["p"], s["p_int"]
    nlev, ncol = T.shape
    gas_amounts, h2o_vmr, co2_vmr = gas_setup(table, s, p_int, g, with_co2=True)
    tau_gas = optical_depth(table, T, p, gas_amounts, h2o_vmr=h2o_vmr, co2_vmr=co2_vmr)
    weights = np.asarray(table["gpoint_weights"], dtype=np.float64)
    nband, ngpt = tau_gas.shape[:2]
    planck_src, surf_src = planck_sources(table, T, s["T_surf"], sigma, nband, ngpt)
    tau = tau_gas + s["tau_cloud_lw"].transpose(2, 0, 1)[:, np.newaxis, :, :]
    ub, db, u, d = lw_transport(tau, planck_src, surf_src, s["emissivity"], weights, D)
    hr = heating_rate(u - d, p_int, g, cpd)
    tau_band, hr_band = _band_diag(tau, weights, ub, db, p_int, g, cpd)
    return {"tau_gas": tau_gas, "planck_src": planck_src, "surf_src": surf_src, "up_band": ub, "down_band": db, "up_broad": u,
            "down_broad": d, "heating_rate": hr, "tau_band": tau_band, "trans_band": np.exp(-D * tau_band), "hr_band": hr_band}


def sw_call(table, s, g, cpd):
    """s: T, p, p_int, q, zenith [rad], albedo, earth_sun_factor, tau_cloud_sw / ssa_cloud / g_cloud (nlev, ncol, nband)."""
    T, p, p_int = s["T"], s["p"], s["p_int"]
    nlev, ncol = T.shape
    gas_amounts, h2o_vmr, _ = gas_setup(table, s, p_int, g, with_co2=False)
    tau_abs = optical_depth(table, T, p, gas_amounts, h2o_vmr=h2o_vmr)
    weights = np.asarray(table["gpoint_weights"], dtype=np.float64)
    nband, ngpt = tau_abs.shape[:2]
    ssa = np.zeros((nband, ngpt, nlev, ncol))
    asym = np.zeros((nband, ngpt, nlev, ncol))
    tau = tau_abs.copy()
    if table.get("rayleigh_coefficient") is not None:
        ray = table["rayleigh_coefficient"]
        for b in range(nband):
            for k in range(nlev):
                dp = abs(p_int[k + 1, :] - p_int[k, :])
                tau_ray = ray[b] * dp / g
                for gp in range(ngpt):
                    tot = tau_abs[b, gp, k, :] + tau_ray
                    ssa[b, gp, k, :] = np.where(tot > 0, tau_ray / np.where(tot > 0, tot, 1.0), 0.0)
                    tau[b, gp, k, :] = tot
    solar_flux = np.asarray(table["solar_source_per_gpoint"]) * float(np.asarray(s["earth_sun_factor"]).reshape(-1)[0])
    tau_c = s["tau_cloud_sw"].transpose(2, 0, 1)[:, np.newaxis, :, :]
    ssa_c = s["ssa_cloud"].transpose(2, 0, 1)[:, np.newaxis, :, :]
    g_c = s["g_cloud"].transpose(2, 0, 1)[:, np.newaxis, :, :]
    tau_total = tau + tau_c
    scat_gas = tau * ssa
    scat_cloud = tau_c * ssa_c
    scat_total = scat_gas + scat_cloud
    ssa_total = np.divide(scat_total, tau_total, out=np.zeros_like(tau_total), where=tau_total > 0)
    g_total = np.divide(scat_gas * asym + scat_cloud * g_c, scat_total, out=np.zeros_like(scat_total), where=scat_total > 0)
    ub, db, u, d = sw_two_stream(tau_total, ssa_total, g_total, s["zenith"], s["albedo"], solar_flux, weights)
    hr = heating_rate(u - d, p_int, g, cpd)
    tau_band, hr_band = _band_diag(tau_total, weights, ub, db, p_int, g, cpd)
    return {"tau_abs": tau_abs, "up_band": ub, "down_band": db, "up_broad": u, "down_broad": d, "heating_rate": hr,
            "tau_band": tau_band, "hr_band": hr_band}
