"""CPU restatement of the Reed-Jablonowski simple-physics package as climt drives it.  TEST INFRASTRUCTURE ONLY.

  simple_physics            climt/_lib/simple_physics/simple_physics_custom.f90:59-565 (arrays top-down, Fortran index k = 1 .. pver)
  get_new_state             climt/_components/simple_physics/_simple_physics.pyx:84-134 (level flip, pdel / rpdel)
  component_call            climt/_components/simple_physics/component.py:224-271 (array_call: constants, latent-heat clamp)

Vectorised over columns, explicit loops over levels in the Fortran's order; the Fortran's un-suffixed (single precision) literals
are kept as float32 values.

PARITY ONLY PARTLY PINNED: the Fortran cannot be compiled or run here (no Fortran compiler, no numba port of this package).  What
pins this restatement: (a) the reference's cached outputs TestSimplePhysics-{column,3d} (default state: calm, dry -- the trivial
path, tests/golden/reference_caches.npz); (b) the reference's own property tests, restated: column moist enthalpy changes by exactly
the surface fluxes (tests/test_conservation.py:295-313); (c) closed-form checks of each process in tests/test_simple_physics.py.
"""
import numpy as np

F32 = lambda x: float(np.float32(x))  # noqa: E731

DEFAULTS = dict(simulate_cyclone=False, large_scale_condensation=True, boundary_layer=True, surface_fluxes=True,
                use_external_surface_temperature=True, use_external_surface_specific_humidity=False,
                top_of_boundary_layer=85000.0, boundary_layer_influence_height=20000.0, drag_coefficient_heat_fluxes=0.0011,
                base_momentum_drag_coefficient=0.0007, wind_dependent_momentum_drag_coefficient=0.000065,
                maximum_momentum_drag_coefficient=0.002)


def simple_physics(dtime, lat, t, q, u, v, pmid, pint, ps, ts, qsurf, K, test, do_lsc, do_pbl, do_surf_flux, use_ts_ext,
                   use_qsurf_ext):
    """Arrays (pver[+1], pcols), model top first (the Fortran's order); t, q, u, v are updated in place.
    K: gravit cpair rair latvap rh2o radius omega rhow pbltop pblconst C Cd0 Cd1 Cm.  -> precl, sens_ht_flux, lat_ht_flux"""
    pver, pcols = t.shape
    gravit, cpair, rair, latvap, rh2o = K["gravit"], K["cpair"], K["rair"], K["latvap"], K["rh2o"]
    epsilo, zvir = rair / rh2o, (rh2o / rair) - 1.0
    pi = F32(4.0 * np.arctan(np.float32(1.0)))
    T0, e0, v20, p0 = 273.16, 610.78, 20.0, 100000.0
    pdel = pint[1:] - pint[:-1]
    rpdel = 1.0 / pdel
    za = rair / gravit * t[-1] * (1.0 + zvir * q[-1]) * 0.5 * (np.log(ps) - np.log(pint[pver - 1]))
    if use_ts_ext:
        Tsurf = np.array(ts, dtype=np.float64)
    elif test == 1:
        T00, u0, eta0, q0 = 288.0, 35.0, 0.252, F32(0.021)
        latw, etav = 2.0 * pi / 9.0, (1.0 - eta0) * 0.5 * pi
        Tsurf = (T00 + pi * u0 / rair * 1.5 * np.sin(etav) * np.cos(etav) ** 0.5
                 * ((-2.0 * np.sin(lat) ** 6 * (np.cos(lat) ** 2 + 1.0 / 3.0) + 10.0 / 63.0) * u0 * np.cos(etav) ** 1.5
                    + (8.0 / 5.0 * np.cos(lat) ** 3 * (np.sin(lat) ** 2 + 2.0 / 3.0) - pi / 4.0) * K["radius"] * K["omega"] * 0.5)
                 ) / (1.0 + zvir * q0 * np.exp(-(lat / latw) ** 4))
    else:
        Tsurf = np.full(pcols, 302.15)
    precl = np.zeros(pcols)
    if do_lsc:
        dtdt, dqdt = np.zeros_like(t), np.zeros_like(t)
        for k in range(pver):
            qsat = epsilo * e0 / pmid[k] * np.exp(-latvap / rh2o * ((1.0 / t[k]) - 1.0 / T0))
            sat = q[k] > qsat
            tmp = 1.0 / dtime * (q[k] - qsat) / (1.0 + (latvap / cpair) * (epsilo * latvap * qsat / (rair * t[k] ** 2)))
            dtdt[k] = np.where(sat, latvap / cpair * tmp, 0.0)
            dqdt[k] = np.where(sat, -tmp, 0.0)
            precl = precl + np.where(sat, tmp * pdel[k] / (gravit * K["rhow"]), 0.0)
        t += dtdt * dtime
        q += dqdt * dtime
    sens, lath = np.zeros(pcols), np.zeros(pcols)
    Km = np.zeros((pver + 1, pcols))
    Ke = np.zeros((pver + 1, pcols))
    if do_surf_flux:
        wind = np.sqrt(u[-1] ** 2 + v[-1] ** 2)
        Ke[pver] = K["C"] * wind * za
        Cd = np.where(wind < v20, K["Cd0"] + K["Cd1"] * wind, K["Cm"])
        Km[pver] = np.where(wind < v20, Cd * wind * za, K["Cm"] * wind * za)
        for k in range(pver):
            above = pint[k] >= K["pbltop"]
            taper = np.exp(-(K["pbltop"] - pint[k]) ** 2 / K["pblconst"] ** 2)
            Km[k] = np.where(above, Km[pver], Km[pver] * taper)
            Ke[k] = np.where(above, Ke[pver], Ke[pver] * taper)
        damp = 1.0 + Cd * wind * dtime / za
        u[-1] = u[-1] / damp
        v[-1] = v[-1] / damp
        rho = pmid[-1] / (rair * t[-1])
        flux = K["C"] * wind * (Tsurf - t[-1])
        sens = rho * cpair * flux
        t[-1] = t[-1] + flux * (rho * gravit) / (pint[pver] - pint[pver - 1]) * dtime
        if use_qsurf_ext:
            qsats = np.array(qsurf, dtype=np.float64)
        else:
            dT = Tsurf - 273.0
            warm = (F32(1.0007) + F32(3.46e-8) * ps) * F32(611.21) * np.exp(F32(17.966) * dT / (F32(247.15) + dT))
            cold = (F32(1.0003) + F32(4.18e-8) * ps) * F32(611.15) * np.exp(F32(22.452) * dT / (F32(272.5) + dT))
            esats = np.where(Tsurf > 271, warm, cold)
            qsats = epsilo * esats / (ps - F32(0.378) * esats)
        rho = pmid[-1] / (rair * t[-1])
        flux = K["C"] * wind * (qsats - q[-1])
        lath = latvap * rho * flux
        q[-1] = q[-1] + flux * (rho * gravit) / (pint[pver] - pint[pver - 1]) * dtime
    if do_pbl:
        CA, CC, CAm, CCm = (np.zeros((pver, pcols)) for _ in range(4))
        for k in range(pver - 1):
            rho = pint[k + 1] / (rair * (t[k + 1] + t[k]) / 2.0)
            dpm = pmid[k + 1] - pmid[k]
            CAm[k] = rpdel[k] * dtime * gravit * gravit * Km[k + 1] * rho * rho / dpm
            CCm[k + 1] = rpdel[k + 1] * dtime * gravit * gravit * Km[k + 1] * rho * rho / dpm
            CA[k] = rpdel[k] * dtime * gravit * gravit * Ke[k + 1] * rho * rho / dpm
            CC[k + 1] = rpdel[k + 1] * dtime * gravit * gravit * Ke[k + 1] * rho * rho / dpm
        CE, CEm, CFu, CFv, CFt, CFq = (np.zeros((pver + 1, pcols)) for _ in range(6))
        kap = rair / cpair
        for k in range(pver - 1, -1, -1):
            den = 1.0 + CA[k] + CC[k] - CA[k] * CE[k + 1]
            denm = 1.0 + CAm[k] + CCm[k] - CAm[k] * CEm[k + 1]
            CE[k] = CC[k] / den
            CEm[k] = CCm[k] / denm
            CFu[k] = (u[k] + CAm[k] * CFu[k + 1]) / denm
            CFv[k] = (v[k] + CAm[k] * CFv[k + 1]) / denm
            CFt[k] = ((p0 / pmid[k]) ** kap * t[k] + CA[k] * CFt[k + 1]) / den
            CFq[k] = (q[k] + CA[k] * CFq[k + 1]) / den
        u[0], v[0], q[0] = CFu[0], CFv[0], CFq[0]
        t[0] = CFt[0] * (pmid[0] / p0) ** kap
        for k in range(1, pver):
            u[k] = CEm[k] * u[k - 1] + CFu[k]
            v[k] = CEm[k] * v[k - 1] + CFv[k]
            t[k] = (CE[k] * t[k - 1] * (p0 / pmid[k - 1]) ** kap + CFt[k]) * (pmid[k] / p0) ** kap
            q[k] = CE[k] * q[k - 1] + CFq[k]
    return precl, sens, lath


def constants(options, C):
    """set_physical_constants as the component calls it (component.py:196-222); C = the sympl constants by name"""
    return dict(gravit=C["gravitational_acceleration"], cpair=C["heat_capacity_of_dry_air_at_constant_pressure"],
                rair=C["gas_constant_of_dry_air"], latvap=C["latent_heat_of_condensation"], rh2o=C["gas_constant_of_vapor_phase"],
                radius=C["planetary_radius"], omega=C["planetary_rotation_rate"], rhow=C["density_of_liquid_water"],
                pbltop=options["top_of_boundary_layer"], pblconst=options["boundary_layer_influence_height"],
                C=options["drag_coefficient_heat_fluxes"], Cd0=options["base_momentum_drag_coefficient"],
                Cd1=options["wind_dependent_momentum_drag_coefficient"], Cm=options["maximum_momentum_drag_coefficient"])


def component_call(state, dtime, C, **options):
    """SimplePhysics(**options).array_call(state, timestep): state arrays (nlev[+1], ncol), level 0 at the surface.
    -> (diagnostics, new_state)"""
    o = dict(DEFAULTS, **options)
    flip = lambda a: np.array(a[::-1], dtype=np.float64)  # noqa: E731
    t, q, u, v = (flip(state[k]) for k in ("air_temperature", "specific_humidity", "eastward_wind", "northward_wind"))
    pmid, pint = flip(state["air_pressure"]), flip(state["air_pressure_on_interface_levels"])
    precl, sens, lath = simple_physics(
        float(dtime), np.asarray(state["latitude"], dtype=np.float64), t, q, u, v, pmid, pint,
        np.asarray(state["surface_air_pressure"], dtype=np.float64), state["surface_temperature"],
        state["surface_specific_humidity"], constants(o, C), int(o["simulate_cyclone"]), int(o["large_scale_condensation"]),
        int(o["boundary_layer"]), int(o["surface_fluxes"]), int(o["use_external_surface_temperature"]),
        int(o["use_external_surface_specific_humidity"]))
    lath = np.where(lath < 0, 0.0, lath)
    return ({"stratiform_precipitation_rate": precl, "surface_upward_sensible_heat_flux": sens, "surface_upward_latent_heat_flux": lath},
            {"eastward_wind": u[::-1].copy(), "northward_wind": v[::-1].copy(), "air_temperature": t[::-1].copy(),
             "specific_humidity": q[::-1].copy()})
