"""CPU restatement of CORK's picket-fence optics (optics="parmentier").  TEST INFRASTRUCTURE ONLY.

Follows, vectorised over columns and levels:
  compute_rosseland_mean_opacity   cork/optics/parmentier.py:55-74     (Freedman et al. 2014 fit)
  lookup_ratio_coefficients        cork/optics/parmentier.py:100-153   (Parmentier & Guillot 2014 ratios; region search quirk kept)
  compute_thermal_opacities        cork/optics/parmentier.py:8-33
  bond_albedo_from_fluxes          cork/optics/parmentier.py:156-163
  CorkLongwaveRadiation._parmentier_optics      cork/lw/component.py:375-422 + array_call :208-373
  CorkShortwaveRadiation._parmentier_sw_optics  cork/sw/component.py:498-532 + array_call :242-306, 414-447
The transport sweeps are the C++ restatements of the numba kernels in oracle/cork_oracle.cpp.

Pinned by tests/golden/parmentier_reference.npz, which tests/golden/make_parmentier_golden.py produced by running the
reference's own component classes (tests/test_parmentier_cpu.py).
"""
import numpy as np

from . import cork as OC


def rosseland_mean_opacity(T, p, fr):
    log_T = np.log10(np.maximum(T, 10.0))
    log_P = np.log10(np.maximum(p * 10.0, 1.0))
    lo = float(fr["a_lo"]) * log_T + float(fr["b_lo"]) * log_P + float(fr["c_lo"])
    hi = float(fr["a_hi"]) * log_T + float(fr["b_hi"]) * log_P + float(fr["c_hi"])
    return 10.0 ** np.where(T < float(fr["T_boundary"]), lo, hi) * 0.1


def ratio_coefficients(co, T_eff):
    """-> gamma_v1, gamma_v2, gamma_v3, beta, gamma_P, R, each (ncol,)"""
    T_eff = np.asarray(T_eff, dtype=np.float64)
    X = np.log10(np.maximum(T_eff, 10.0))
    b = np.asarray(co["T_eff_boundaries"], dtype=np.float64)
    region = np.zeros(T_eff.shape, dtype=np.int64)   # stays 0 when no interval matches (parmentier.py:115-119)
    found = np.zeros(T_eff.shape, dtype=bool)
    for i in range(len(b) - 1):
        hit = (T_eff >= b[i]) & (T_eff < b[i + 1]) & ~found
        region[hit] = i
        found |= hit

    def lin(name):
        ab = np.asarray(co[name], dtype=np.float64)
        return ab[region, 0] + ab[region, 1] * X

    gv3, gv2, gv1 = 10.0 ** lin("log10_gamma_v3_ab"), 10.0 ** lin("log10_gamma_v2_ab"), 10.0 ** lin("log10_gamma_v1_ab")
    beta = np.clip(lin("beta_ab"), 0.01, 0.99)
    quad = np.asarray(co["log10_gamma_P_quad"], dtype=np.float64)
    gamma_P = np.maximum(10.0 ** (quad[0] + quad[1] * X + quad[2] * X ** 2), 1.0)
    gm1 = gamma_P - 1.0
    disc = gm1 ** 2 + 4.0 * beta * (1.0 - beta) * gm1
    with np.errstate(invalid="ignore"):
        R = 1.0 + gm1 / (2.0 * beta * (1.0 - beta)) + np.sqrt(disc) / (2.0 * beta * (1.0 - beta))
    R = np.where(disc < 0, 1.0, np.maximum(R, 1.0))
    return gv1, gv2, gv3, beta, gamma_P, R


def effective_temperature(T_irr, T_int, A_B=0.0):
    return np.maximum((T_int ** 4 + (1.0 - A_B) * 0.25 * T_irr ** 4) ** 0.25, 100.0)


def lw_call(co, fr, s, g, cpd, sigma, D=1.66):
    """s: T, p, p_int, T_surf, T_irr, T_int, emissivity (2, ncol), tau_cloud_lw (nlev, ncol, 2)."""
    T, p, p_int = s["T"], s["p"], s["p_int"]
    nlev, ncol = T.shape
    gv1, gv2, gv3, beta, gamma_P, R = ratio_coefficients(co, effective_temperature(s["T_irr"], s["T_int"]))
    kappa_R = rosseland_mean_opacity(T, p, fr)
    kappa_2 = kappa_R * (beta / R + 1.0 - beta)[None, :]
    kappa_1 = R[None, :] * kappa_2
    mass = np.abs(p_int[1:] - p_int[:-1]) / g
    tau = np.zeros((2, 1, nlev, ncol))
    tau[0, 0], tau[1, 0] = kappa_1 * mass, kappa_2 * mass
    planck = sigma * T ** 4
    planck_src = np.stack([beta[None, :] * planck, (1.0 - beta)[None, :] * planck])[:, None]
    sp = sigma * s["T_surf"] ** 4
    surf_src = np.stack([beta * sp, (1.0 - beta) * sp])[:, None]
    weights = np.ones((2, 1))
    tau = tau + s["tau_cloud_lw"].transpose(2, 0, 1)[:, np.newaxis, :, :]
    ub, db, u, d = OC.lw_transport(tau, planck_src, surf_src, s["emissivity"], weights, D)
    hr = OC.heating_rate(u - d, p_int, g, cpd)
    tau_band, hr_band = OC._band_diag(tau, weights, ub, db, p_int, g, cpd)
    return {"up_band": ub, "down_band": db, "up_broad": u, "down_broad": d, "heating_rate": hr, "tau_band": tau_band,
            "trans_band": np.exp(-D * tau_band), "hr_band": hr_band}


def sw_call(co, fr, s, g, cpd, sigma, bond_albedo_feedback=False, default_solar_flux_per_band=None):
    """s: T, p, p_int, T_irr, T_int, zenith, albedo, earth_sun_factor, tau_cloud_sw / ssa_cloud / g_cloud (nlev, ncol, 3)."""
    T, p, p_int = s["T"], s["p"], s["p_int"]
    nlev, ncol = T.shape
    T_irr_max = s["T_irr"].max()
    if T_irr_max > 0:
        spb = np.array([sigma * T_irr_max ** 4 / 3.0] * 3)
    else:
        spb = np.asarray(default_solar_flux_per_band, dtype=np.float64)
    solar_flux = spb.reshape(3, 1) * np.ones((3, 1)) * float(np.asarray(s["earth_sun_factor"]).reshape(-1)[0])
    weights = np.ones((3, 1))
    tau_c = s["tau_cloud_sw"].transpose(2, 0, 1)[:, np.newaxis, :, :]
    ssa_c = s["ssa_cloud"].transpose(2, 0, 1)[:, np.newaxis, :, :]
    g_c = s["g_cloud"].transpose(2, 0, 1)[:, np.newaxis, :, :]
    kappa_R = rosseland_mean_opacity(T, p, fr)
    mass = np.abs(p_int[1:] - p_int[:-1]) / g
    A_B = np.zeros(ncol)
    for _ in range(2 if bond_albedo_feedback else 1):
        gv = ratio_coefficients(co, effective_temperature(s["T_irr"], s["T_int"], A_B))[:3]
        tau = np.stack([(gv[b][None, :] * kappa_R) * mass for b in range(3)])[:, None]
        ssa = np.zeros_like(tau)
        asym = np.zeros_like(tau)
        tau_total = tau + tau_c
        scat_gas = tau * ssa
        scat_cloud = tau_c * ssa_c
        scat_total = scat_gas + scat_cloud
        ssa_total = np.divide(scat_total, tau_total, out=np.zeros_like(tau_total), where=tau_total > 0)
        g_total = np.divide(scat_gas * asym + scat_cloud * g_c, scat_total, out=np.zeros_like(scat_total), where=scat_total > 0)
        ub, db, u, d = OC.sw_two_stream(tau_total, ssa_total, g_total, s["zenith"], s["albedo"], solar_flux, weights)
        with np.errstate(divide="ignore", invalid="ignore"):
            A_B = np.clip(np.where(d[-1] > 0, u[-1] / d[-1], 0.0), 0.0, 1.0)
    hr = OC.heating_rate(u - d, p_int, g, cpd)
    tau_band, hr_band = OC._band_diag(tau_total, weights, ub, db, p_int, g, cpd)
    return {"up_band": ub, "down_band": db, "up_broad": u, "down_broad": d, "heating_rate": hr, "tau_band": tau_band,
            "hr_band": hr_band, "bond_albedo": A_B, "solar_flux": solar_flux}
