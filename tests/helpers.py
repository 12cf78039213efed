"""Shared test helpers: oracle drivers, default states in the C-ABI layout, the host emulation."""
import ctypes
import os
import subprocess

import numpy as np

from climt_b200 import constants as C
from climt_b200 import rrtmg_tables as RT
from climt_b200 import state as S
from climt_b200 import synthetic as SY
from climt_b200 import tables as T

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_caches.npz")
_dp = ctypes.POINTER(ctypes.c_double)

ORC_ARG_ORDER = ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "cfc11", "cfc12",
                 "cfc22", "ccl4", "emis", "cldfr", "tauaer", "taucld", "cicewp", "cliqwp", "reice", "reliq")
# synthetic/oracle short names -> C ABI names
ABI_NAME = {"h2o": "h2ovmr", "o3": "o3vmr", "co2": "co2vmr", "ch4": "ch4vmr", "n2o": "n2ovmr", "o2": "o2vmr",
            "cfc11": "cfc11vmr", "cfc12": "cfc12vmr", "cfc22": "cfc22vmr", "ccl4": "ccl4vmr"}


def to_abi(st):
    return {ABI_NAME.get(k, k): v for k, v in st.items()}


def lw_oracle(**flags):
    from oracle.rrtmg import LWOracle
    return LWOracle(C.rrtmg_constants(), T.raw_blob_path("lw"), **flags)


def run_lw_oracle(orc, st):
    return orc(*[st[k] for k in ORC_ARG_ORDER])


def default_lw_abi_state(nz, ncol=1, external_tint=False):
    """climt's default RRTMGLongwave state converted to the C-ABI arrays (what array_call hands down)."""
    d = S.default_rrtmg_lw_state(nz, ncol)
    p, pi = d["air_pressure"] / 100.0, d["air_pressure_on_interface_levels"] / 100.0
    st = {
        "play": p, "plev": pi, "tlay": d["air_temperature"], "tsfc": d["surface_temperature"],
        "h2o": S.mass_to_volume_mixing_ratio(d["specific_humidity"], 18.02),
        "o3": d["mole_fraction_of_ozone_in_air"], "co2": d["mole_fraction_of_carbon_dioxide_in_air"],
        "ch4": d["mole_fraction_of_methane_in_air"], "n2o": d["mole_fraction_of_nitrous_oxide_in_air"],
        "o2": d["mole_fraction_of_oxygen_in_air"], "cfc11": d["mole_fraction_of_cfc11_in_air"],
        "cfc12": d["mole_fraction_of_cfc12_in_air"], "cfc22": d["mole_fraction_of_cfc22_in_air"],
        "ccl4": d["mole_fraction_of_carbon_tetrachloride_in_air"], "emis": d["surface_longwave_emissivity"],
        "cldfr": d["cloud_area_fraction_in_atmosphere_layer"],
        "taucld": d["longwave_optical_thickness_due_to_cloud"],
        "cicewp": d["mass_content_of_cloud_ice_in_atmosphere_layer"] * 1e3,
        "cliqwp": d["mass_content_of_cloud_liquid_water_in_atmosphere_layer"] * 1e3,
        "reice": d["cloud_ice_particle_size"], "reliq": d["cloud_water_droplet_radius"],
        "tauaer": d["longwave_optical_thickness_due_to_aerosol"],
    }
    if external_tint:
        st["tlev"] = np.full((nz + 1, ncol), 290.0)
    else:
        st["tlev"] = S.get_interface_values(st["tlay"], st["tsfc"], p, pi)
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}


SW_ABI_NAME = dict(ABI_NAME)


def to_abi_sw(st):
    return {ABI_NAME.get(k, k): v for k, v in st.items()}


def sw_oracle(**flags):
    from oracle.rrtmg import SWOracle
    return SWOracle(C.rrtmg_constants(), T.raw_blob_path("sw"), **flags)


def default_sw_abi_state(nz, ncol=1):
    d = S.default_rrtmg_sw_state(nz, ncol)
    p, pi = d["air_pressure"] / 100.0, d["air_pressure_on_interface_levels"] / 100.0
    st = dict(
        play=p, plev=pi, tlay=d["air_temperature"], tsfc=d["surface_temperature"],
        h2o=S.mass_to_volume_mixing_ratio(d["specific_humidity"], 18.02), o3=d["mole_fraction_of_ozone_in_air"],
        co2=d["mole_fraction_of_carbon_dioxide_in_air"], ch4=d["mole_fraction_of_methane_in_air"],
        n2o=d["mole_fraction_of_nitrous_oxide_in_air"], o2=d["mole_fraction_of_oxygen_in_air"],
        asdir=d["surface_albedo_for_direct_shortwave"], asdif=d["surface_albedo_for_diffuse_shortwave"],
        aldir=d["surface_albedo_for_direct_near_infrared"], aldif=d["surface_albedo_for_diffuse_near_infrared"],
        coszen=np.cos(d["zenith_angle"]), cldfr=d["cloud_area_fraction_in_atmosphere_layer"],
        taucld=d["shortwave_optical_thickness_due_to_cloud"], ssacld=d["single_scattering_albedo_due_to_cloud"],
        asmcld=d["cloud_asymmetry_parameter"], fsfcld=d["cloud_forward_scattering_fraction"],
        cicewp=d["mass_content_of_cloud_ice_in_atmosphere_layer"] * 1e3,
        cliqwp=d["mass_content_of_cloud_liquid_water_in_atmosphere_layer"] * 1e3,
        reice=d["cloud_ice_particle_size"], reliq=d["cloud_water_droplet_radius"],
        tauaer=d["shortwave_optical_thickness_due_to_aerosol"], ssaaer=d["single_scattering_albedo_due_to_aerosol"],
        asmaer=d["aerosol_asymmetry_parameter"], ecaer=d["aerosol_optical_depth_at_55_micron"])
    st["tlev"] = S.get_interface_values(st["tlay"], st["tsfc"], p, pi)
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}


def golden():
    return np.load(GOLDEN)


def rel_err(got, ref, floor=1e-3):
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor)))


def _build_shared(so, src, deps, flags):
    """Compile src -> so when it is older than a dependency.  Several pytest-xdist workers may want the same library at once:
    the build is serialised by a file lock, written to a temporary name and moved into place, so nobody ever loads a half-written
    file."""
    import fcntl

    def stale():
        return not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps)
    if stale():
        with open(so + ".lock", "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            try:
                if stale():
                    tmp = f"{so}.{os.getpid()}.tmp"
                    subprocess.check_call(["g++"] + flags + ["-shared", "-o", tmp, src])
                    os.replace(tmp, so)
            finally:
                fcntl.flock(lk, fcntl.LOCK_UN)
    return ctypes.CDLL(so)


# ---- host emulation of the kernel code (tests/emul/lw_emul.cpp) ---------------------------------
def emul_lib(which="lw"):
    so = os.path.join(HERE, "emul", "libcb_emul.so" if which == "lw" else "libcb_emul_sw.so")
    src = os.path.join(HERE, "emul", f"{which}_emul.cpp")
    deps = [src] + [os.path.join(HERE, "..", "climt_b200", "csrc", f)
                    for f in ("lw_core.cuh", "lw_tables.h", "cb_common.h", "sw_core.cuh", "sw_tables.h", "mcica_core.cuh",
                              "mcica_host.h")]
    return _build_shared(so, src, deps, ["-O1", "-std=c++17", "-fPIC", "-ffp-contract=off"])


def run_lw_emul(st, flags=(1, 0, 2, 1, 1), mcica=(0, 1, 0), tile=False):
    """flags = (icld, idrv, inflag, iceflag, liqflag); mcica = (enabled, irng, permuteseed); tile: the column-tile form of the
    transfer (lw_tile_cell / lw_tile_sweeps) instead of lw_transfer_unit"""
    flags = tuple(flags) + tuple(mcica)
    lib = emul_lib()
    lib.emul_lw_set_tile(1 if tile else 0)
    k = C.rrtmg_constants()
    consts = np.array([k[n] for n in ("pi", "grav", "planck", "boltz", "clight", "avogad", "alosmt", "gascon",
                                      "sbcnst", "secdy", "cpdair")])
    nlay, ncol = st["play"].shape
    inp = (_dp * 23)(*[st[f].ctypes.data_as(_dp) for f in SY.LW_FIELDS])
    out = {n: np.zeros((nlay + 1, ncol)) for n in ("uflx", "dflx", "uflxc", "dflxc")}
    out.update({n: np.zeros((nlay, ncol)) for n in ("hr", "hrc")})
    out.update({n: np.zeros((nlay + 1, ncol)) for n in ("duflx_dt", "duflxc_dt")})
    outp = (_dp * 8)(*[out[n].ctypes.data_as(_dp) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc", "duflx_dt", "duflxc_dt")])
    rc = lib.emul_lw_run(RT.lw_blob_path().encode(), consts.ctypes.data_as(_dp), (ctypes.c_int * 8)(*flags),
                         ncol, nlay, inp, outp)
    return rc, out


def run_sw_emul(st, iopt=(1, 0, 2, 1, 1, 0, 1), scal=None, mcica=(0, 1, 0), tile=False):
    """iopt = (icld, iaer, inflag, iceflag, liqflag, isolvar, dyofyr); scal = [adjes, scon, solcycfrac, ind0, ind1, bnd[14]];
    tile: the column-tile form of the transfer (sw_tile_cell / sw_tile_sweeps) instead of sw_transfer_unit"""
    lib = emul_lib("sw")
    lib.emul_sw_set_tile(int(tile))  # 0 unit form, 1 tile form with serial sweeps, 2 tile form with the scan-form sweeps
    k = C.rrtmg_constants()
    consts = np.array([k[n] for n in ("pi", "grav", "planck", "boltz", "clight", "avogad", "alosmt", "gascon",
                                      "sbcnst", "secdy", "cpdair")])
    scal = np.array(scal if scal is not None else [1.0, 1367.0, 0.0, 1.0, 1.0] + [1.0] * 14, dtype=np.float64)
    nlay, ncol = st["play"].shape
    inp = (_dp * 29)(*[st[f].ctypes.data_as(_dp) for f in SY.SW_FIELDS])
    out = {n: np.zeros((nlay + 1, ncol)) for n in ("uflx", "dflx", "uflxc", "dflxc")}
    out.update({n: np.zeros((nlay, ncol)) for n in ("hr", "hrc")})
    outp = (_dp * 6)(*[out[n].ctypes.data_as(_dp) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")])
    iopt = tuple(iopt) + tuple(mcica)
    rc = lib.emul_sw_run(RT.sw_blob_path().encode(), consts.ctypes.data_as(_dp), (ctypes.c_int * 10)(*iopt),
                         scal.ctypes.data_as(_dp), ncol, nlay, inp, outp)
    return rc, out


SW_KEYS = {"uflx": "swuflx", "dflx": "swdflx", "uflxc": "swuflxc", "dflxc": "swdflxc", "hr": "swhr", "hrc": "swhrc"}


# ---- CORK -----------------------------------------------------------------------------------------------------------
CORK_G, CORK_CPD, CORK_SIGMA = 9.80665, 1004.64, 5.670367e-08   # sympl defaults


def cork_table(name):
    from climt_b200 import cork
    return cork.load_k_table(name)


def cork_emul_lib():
    so = os.path.join(HERE, "emul", "libcb_emul_cork.so")
    src = os.path.join(HERE, "emul", "cork_emul.cpp")
    deps = [src] + [os.path.join(HERE, "..", "climt_b200", "csrc", f) for f in ("cork_core.cuh", "cork_tables.h", "cb_common.h")]
    deps.append(os.path.join(HERE, "..", "include", "climt_b200.h"))
    return _build_shared(so, src, deps, ["-O1", "-std=c++17", "-fPIC", "-ffp-contract=off"])


def cork_arrays(s, which):
    """golden/oracle state dict -> engine input arrays (cb200_cork_inputs names)"""
    a = {"T": s["T"], "p": s["p"], "p_int": s["p_int"], "q_h2o": s["q"]}
    if which == "lw":
        a.update(T_surf=s["T_surf"], co2_vmr=s["co2"], emissivity=s["emissivity"], tau_cloud=s["tau_cloud_lw"])
    else:
        a.update(zenith=s["zenith"], albedo=s["albedo"], tau_cloud=s["tau_cloud_sw"], ssa_cloud=s["ssa_cloud"], g_cloud=s["g_cloud"])
    return a


def cork_solar_flux(table, earth_sun_factor):
    """solar_source * earth_sun_factor with the reference's dtype semantics (sw/component.py:371-372), as float64"""
    return np.ascontiguousarray(np.asarray(table["solar_source_per_gpoint"]) * float(earth_sun_factor), dtype=np.float64)


def run_cork_emul(table, which, arrays, scalar, umax=4, diagnostics_level=0):
    """The CUDA engine's per-thread code, compiled for the host.  scalar = D (lw) or earth_sun_factor (sw).
    diagnostics_level >= 1: returns (out, {component diagnostic name: (nband, nlev[+1], ncol)})"""
    from climt_b200 import cork
    lib = cork_emul_lib()
    if cork.is_esft(table):
        table = cork.expand_esft_table(table)
    ct, keep = cork.make_ctable(table)
    nlev, ncol = arrays["T"].shape
    nb = ct.nband
    shapes = {"up_broad": (nlev + 1, ncol), "down_broad": (nlev + 1, ncol), "heating_rate": (nlev, ncol), "up_band": (nb, nlev + 1, ncol),
              "down_band": (nb, nlev + 1, ncol), "tau_band": (nb, nlev, ncol), "trans_band": (nb, nlev, ncol), "hr_band": (nb, nlev, ncol)}
    out = {k: np.zeros(v) for k, v in shapes.items()}
    ins = []
    for k in cork.CORK_IN:
        a = arrays.get(k)
        if a is not None:
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
        ins.append(a)
    inp = (_dp * len(ins))(*[a.ctypes.data_as(_dp) if a is not None else None for a in ins])
    outp = (_dp * 8)(*[out[k].ctypes.data_as(_dp) for k in cork.CORK_OUT])
    scal = np.array([CORK_G, CORK_CPD, CORK_SIGMA, scalar if which == "lw" else 1.66])
    solar = cork_solar_flux(table, scalar) if which == "sw" else np.zeros(1)
    if diagnostics_level:
        fields, ptrs = {}, [None] * 10
        for j, (name, iface, minlevel) in (cork.LW_DIAG if which == "lw" else cork.SW_DIAG).items():
            if diagnostics_level >= minlevel:
                fields[name] = np.full((nb, nlev + 1 if iface else nlev, ncol), np.nan)
                ptrs[j] = fields[name].ctypes.data_as(_dp)
        diagp = (_dp * 10)(*ptrs)
        wsum = np.ascontiguousarray(np.asarray(table["gpoint_weights"]).sum(axis=1), dtype=np.float64)
        lib.emul_cork_run_diag.argtypes = [ctypes.POINTER(cork.CorkTable), ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_int,
                                           ctypes.POINTER(_dp), ctypes.POINTER(_dp), ctypes.c_int, ctypes.POINTER(_dp), _dp]
        rc = lib.emul_cork_run_diag(ctypes.byref(ct), 1 if which == "lw" else 0, umax, scal.ctypes.data_as(_dp), solar.ctypes.data_as(_dp),
                                    ncol, nlev, ctypes.cast(inp, ctypes.POINTER(_dp)), ctypes.cast(outp, ctypes.POINTER(_dp)),
                                    int(diagnostics_level), ctypes.cast(diagp, ctypes.POINTER(_dp)), wsum.ctypes.data_as(_dp))
        assert rc == 0
        return out, fields
    lib.emul_cork_run.argtypes = [ctypes.POINTER(cork.CorkTable), ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_int,
                                  ctypes.POINTER(_dp), ctypes.POINTER(_dp)]
    rc = lib.emul_cork_run(ctypes.byref(ct), 1 if which == "lw" else 0, umax, scal.ctypes.data_as(_dp), solar.ctypes.data_as(_dp), ncol, nlev,
                           ctypes.cast(inp, ctypes.POINTER(_dp)), ctypes.cast(outp, ctypes.POINTER(_dp)))
    assert rc == 0
    return out


# ---- CORK picket fence (optics="parmentier") ------------------------------------------------------------------------
PARMENTIER_GOLDEN = os.path.join(HERE, "golden", "parmentier_reference.npz")


def picket_coefficients():
    from climt_b200 import cork
    return cork.load_parmentier_coefficients("solar_composition"), cork.load_freedman2014_coefficients()


def parmentier_case(z, case):
    return {k.split("/")[-1]: z[k] for k in z.files if k.startswith(case + "/in/")}


def picket_arrays(s, which, bond_albedo=None):
    a = {"T": s["T"], "p": s["p"], "p_int": s["p_int"], "T_irr": s["T_irr"], "T_int": s["T_int"]}
    if which == "lw":
        a.update(T_surf=s["T_surf"], emissivity=s["emissivity"], tau_cloud=s["tau_cloud_lw"])
    else:
        a.update(zenith=s["zenith"], albedo=s["albedo"], tau_cloud=s["tau_cloud_sw"], ssa_cloud=s["ssa_cloud"], g_cloud=s["g_cloud"])
        if bond_albedo is not None:
            a["bond_albedo"] = bond_albedo
    return a


def run_picket_emul(which, arrays, scalar, solar_flux=None):
    """The picket-fence engine's per-thread code compiled for the host.  scalar = D (lw); solar_flux (3, 1) for sw."""
    from climt_b200 import cork
    lib = cork_emul_lib()
    co, fr = picket_coefficients()
    pc = cork.make_picket_coeffs(co, fr)
    nlev, ncol = arrays["T"].shape
    nb = 2 if which == "lw" else 3
    shapes = {"up_broad": (nlev + 1, ncol), "down_broad": (nlev + 1, ncol), "heating_rate": (nlev, ncol), "up_band": (nb, nlev + 1, ncol),
              "down_band": (nb, nlev + 1, ncol), "tau_band": (nb, nlev, ncol), "trans_band": (nb, nlev, ncol), "hr_band": (nb, nlev, ncol)}
    out = {k: np.zeros(v) for k, v in shapes.items()}
    ins = [None if arrays.get(k) is None else np.ascontiguousarray(arrays[k], dtype=np.float64) for k in cork.CORK_IN]
    inp = (_dp * len(ins))(*[a.ctypes.data_as(_dp) if a is not None else None for a in ins])
    outp = (_dp * 8)(*[out[k].ctypes.data_as(_dp) for k in cork.CORK_OUT])
    scal = np.array([CORK_G, CORK_CPD, CORK_SIGMA, scalar if which == "lw" else 1.66])
    solar = np.ascontiguousarray(solar_flux, dtype=np.float64) if solar_flux is not None else np.zeros(3)
    lib.emul_picket_run.argtypes = [ctypes.POINTER(cork.PicketCoeffs), ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_int,
                                    ctypes.POINTER(_dp), ctypes.POINTER(_dp)]
    rc = lib.emul_picket_run(ctypes.byref(pc), 1 if which == "lw" else 0, scal.ctypes.data_as(_dp), solar.ctypes.data_as(_dp), ncol, nlev,
                             ctypes.cast(inp, ctypes.POINTER(_dp)), ctypes.cast(outp, ctypes.POINTER(_dp)))
    assert rc == 0
    return out


# golden diagnostic name -> (engine output, band axis moved to the end?)
PICKET_LW_DIAG = {"T": ("heating_rate", False), "upwelling_longwave_flux_in_air": ("up_broad", False),
                  "downwelling_longwave_flux_in_air": ("down_broad", False),
                  "upwelling_longwave_flux_in_air_per_band": ("up_band", True),
                  "downwelling_longwave_flux_in_air_per_band": ("down_band", True),
                  "longwave_optical_depth_per_band": ("tau_band", True), "longwave_transmittance_per_band": ("trans_band", True),
                  "air_temperature_tendency_from_longwave_per_band": ("hr_band", True)}
PICKET_SW_DIAG = {"T": ("heating_rate", False), "upwelling_shortwave_flux_in_air": ("up_broad", False),
                  "downwelling_shortwave_flux_in_air": ("down_broad", False),
                  "upwelling_shortwave_flux_in_air_per_band": ("up_band", True),
                  "downwelling_shortwave_flux_in_air_per_band": ("down_band", True),
                  "shortwave_optical_depth_per_band": ("tau_band", True),
                  "air_temperature_tendency_from_shortwave_per_band": ("hr_band", True)}


def flux_scaled_err(got, ref, scale):
    """max |got - ref| / scale: for quantities that are differences of fluxes (heating rates), measured against the flux scale"""
    return float(np.max(np.abs(got - ref)) / scale)


# ---- Emanuel convection ---------------------------------------------------------------------------------------------
EMANUEL_GOLDEN = os.path.join(HERE, "golden", "emanuel_reference.npz")


def emanuel_case(z, case):
    return {k.split("/")[-1]: z[k] for k in z.files if k.startswith(case + "/in/")}


def emanuel_arrays(st, qs=None):
    a = {"t": st["air_temperature"], "q": st["specific_humidity"], "u": st["eastward_wind"], "v": st["northward_wind"],
         "p": st["air_pressure"], "ph": st["air_pressure_on_interface_levels"], "cbmf": st["cloud_base_mass_flux"]}
    if qs is not None:
        a["qs"] = qs
    return a


def emanuel_emul_lib():
    so = os.path.join(HERE, "emul", "libcb_emul_emanuel.so")
    src = os.path.join(HERE, "emul", "emanuel_emul.cpp")
    deps = [src, os.path.join(HERE, "..", "climt_b200", "csrc", "emanuel_core.cuh"), os.path.join(HERE, "..", "climt_b200", "csrc", "cb_common.h"),
            os.path.join(HERE, "..", "include", "climt_b200.h")]
    return _build_shared(so, src, deps, ["-O1", "-std=c++17", "-fPIC", "-ffp-contract=off"])


def run_emanuel_emul(params, arrays, dt, qs_mode, max_conv_lev=None, layout=1):
    """The CUDA engine's column code compiled for the host; arrays in the component's (ncol, nlev) layout (layout 1) or the
    radiation engines' (nlev, ncol) layout (layout 0)."""
    from climt_b200 import emanuel as EM
    lib = emanuel_emul_lib()
    ncol, nlev = arrays["t"].shape if layout == 1 else arrays["t"].shape[::-1]
    nl = nlev - 3 if max_conv_lev is None else max_conv_lev
    p = EM.make_params(**params)
    pin, keep = EM.EmanuelInputs(), []
    for k in EM.EM_IN:
        if arrays.get(k) is not None:
            a = np.ascontiguousarray(arrays[k], dtype=np.float64)
            keep.append(a)
            setattr(pin, k, a.ctypes.data_as(_dp))
    ins, outs = EM.EmanuelEngine.shapes(ncol, nlev, layout)
    out = {k: np.zeros(outs[k]) for k in EM.EM_OUT}
    out["iflag"] = np.zeros(ncol, dtype=np.int32)
    pout = EM.EmanuelOutputs()
    for k in EM.EM_OUT:
        setattr(pout, k, out[k].ctypes.data_as(_dp))
    pout.iflag = out["iflag"].ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    lib.emul_emanuel_run.argtypes = [ctypes.POINTER(EM.EmanuelParams), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                     ctypes.c_int, ctypes.POINTER(EM.EmanuelInputs), ctypes.POINTER(EM.EmanuelOutputs)]
    rc = lib.emul_emanuel_run(ctypes.byref(p), ncol, nlev, nl, float(dt), int(qs_mode), int(layout), ctypes.byref(pin), ctypes.byref(pout))
    assert rc == 0
    return out


# engine output -> (tendency or diagnostic name of the components)
EMANUEL_OUT = {"ft": "tendency_air_temperature", "fq": "tendency_specific_humidity", "fu": "tendency_eastward_wind",
               "fv": "tendency_northward_wind", "iflag": "convective_state", "precip": "convective_precipitation_rate",
               "wd": "convective_downdraft_velocity_scale", "tprime": "convective_downdraft_temperature_scale",
               "qprime": "convective_downdraft_specific_humidity_scale", "cbmf": "cloud_base_mass_flux",
               "cape": "atmosphere_convective_available_potential_energy"}


def emanuel_compare(got, ref, rtol, what=""):
    """got/ref: dicts keyed like the engine outputs.  The convective state must agree exactly; every other field to rtol of the
    field's own scale (tendencies are sums of terms of either sign: a per-element relative test would measure cancellation)."""
    assert np.array_equal(got["iflag"], ref["iflag"]), (what, "convective_state differs in", int((got["iflag"] != ref["iflag"]).sum()), "columns")
    worst = 0.0
    for k in ("ft", "fq", "fu", "fv", "precip", "wd", "tprime", "qprime", "cbmf", "cape"):
        scale = float(np.max(np.abs(ref[k])))
        err = float(np.max(np.abs(got[k] - ref[k]))) / scale if scale > 0 else float(np.max(np.abs(got[k])))
        assert err < rtol, (what, k, err)
        worst = max(worst, err)
    return worst


def host_pipe_emul_lib():
    """host-side helpers of the engines' host-pointer path (worker pool, all-zero scan), compiled without CUDA"""
    so = os.path.join(HERE, "emul", "libcb_emul_hostpipe.so")
    src = os.path.join(HERE, "emul", "host_pipe_emul.cpp")
    deps = [src, os.path.join(HERE, "..", "climt_b200", "csrc", "engine_common.h")]
    return _build_shared(so, src, deps, ["-O2", "-std=c++17", "-fPIC", "-pthread"])


def simple_physics_emul_lib():
    """k_simple_physics' per-column code compiled for the host (tests/emul/simple_physics_emul.cpp)"""
    so = os.path.join(HERE, "emul", "libcb_emul_simple_physics.so")
    src = os.path.join(HERE, "emul", "simple_physics_emul.cpp")
    deps = [src] + [os.path.join(HERE, "..", "climt_b200", "csrc", f) for f in ("simple_physics_core.cuh", "cb_common.h")]
    return _build_shared(so, src, deps, ["-O1", "-std=c++17", "-fPIC", "-ffp-contract=off"])


def run_simple_physics_emul(state, dtime, params, order=0):
    """state keyed by the component's input names, (nlev[+1], ncol) arrays -> the seven outputs of the C ABI"""
    from climt_b200 import simple_physics as SP
    lib = simple_physics_emul_lib()
    nlev, ncol = state["air_temperature"].shape
    keep, s = [], SP._InHost()
    for k, name in SP._STATE.items():
        a = np.ascontiguousarray(state[name], dtype=np.float64)
        keep.append(a)
        setattr(s, k, a.ctypes.data_as(_dp))
    out = {k: np.full((nlev, ncol) if k in ("t", "q", "u", "v") else (ncol,), np.nan) for k in SP._OUT}
    o = SP._OutHost()
    for k in SP._OUT:
        setattr(o, k, out[k].ctypes.data_as(_dp))
    lib.emul_simple_physics_run.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double, ctypes.POINTER(SP.Params), ctypes.POINTER(SP._InHost),
                                                                 ctypes.POINTER(SP._OutHost)]
    rc = lib.emul_simple_physics_run(ncol, nlev, order, float(dtime), ctypes.byref(params), ctypes.byref(s), ctypes.byref(o))
    assert rc == 0
    return out
