"""The reference's two-step McICA C ABI on the GPU, called through ctypes in the order the unchanged Cython shims use
(climt/_components/rrtmg/lw/_rrtmg_lw.pyx:261-320, sw/_rrtmg_sw.pyx:341-417): set constants -> ini -> mcica_subcol_*_wrapper
fills the (ngpt, ncol, nlay) arrays -> rrtmg_*_mcica_wrapper consumes them.  Pinned by the reference's own McICA goldens
(TestRRTMG{Longwave,Shortwave}MCICA-3d, atol 1e-8 like tests/test_components.py:355-356) and by the oracle for kissvec."""
import ctypes

import numpy as np
import pytest

import helpers as H
from climt_b200 import _native, constants as C, synthetic as SY

pytestmark = pytest.mark.gpu
_dp = ctypes.POINTER(ctypes.c_double)

LW_PRE = ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "cfc11", "cfc12", "cfc22", "ccl4", "emis")
SW_PRE = ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "asdir", "asdif", "aldir", "aldif", "coszen")


def _d(x):
    return ctypes.byref(ctypes.c_double(float(x)))


def _i(x):
    return ctypes.byref(ctypes.c_int(int(x)))


def _p(a):
    return a.ctypes.data_as(_dp)


def _lib():
    L = _native.lib()
    for n in ("rrtmg_set_constants", "rrtmg_lw_ini_wrapper", "mcica_subcol_lw_wrapper", "rrtmg_lw_mcica_wrapper",
              "rrtmg_lw_nomcica_wrapper", "rrtmg_sw_set_constants", "rrtmg_sw_ini_wrapper", "mcica_subcol_sw_wrapper",
              "rrtmg_sw_mcica_wrapper"):
        getattr(L, n).restype = None
        getattr(L, n).argtypes = None
    L.cb200_global_error.restype = ctypes.c_char_p
    return L


def _init(L, which):
    from climt_b200 import rrtmg_tables
    k = C.rrtmg_constants()
    args = [_d(k[n]) for n in ("pi", "grav", "planck", "boltz", "clight", "avogad", "alosmt", "gascon", "sbcnst", "secdy")]
    if which == "lw":
        rrtmg_tables.lw_blob_path()
        L.rrtmg_set_constants(*args)
        L.rrtmg_lw_ini_wrapper(_d(k["cpdair"]))
    else:
        rrtmg_tables.sw_blob_path()
        L.rrtmg_sw_set_constants(*args)
        L.rrtmg_sw_ini_wrapper(_d(k["cpdair"]))


def _lw_two_step(L, st, icld, irng, seed, idrv=0):
    nlay, ncol = st["play"].shape
    st = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}
    mc = {k: np.zeros((nlay, ncol, 140)) for k in ("cldf", "ciwp", "clwp", "tauc")}   # what _get_mcica_arrays allocates
    rei, rel = np.zeros((nlay, ncol)), np.zeros((nlay, ncol))
    irng_io = ctypes.c_int(irng)
    L.mcica_subcol_lw_wrapper(_i(1), _i(ncol), _i(nlay), _i(icld), _i(seed), ctypes.byref(irng_io),
                              *[_p(st[k]) for k in ("play", "cldfr", "cicewp", "cliqwp", "reice", "reliq", "taucld")],
                              _p(mc["cldf"]), _p(mc["ciwp"]), _p(mc["clwp"]), _p(rei), _p(rel), _p(mc["tauc"]))
    out = {n: np.zeros((nlay + 1, ncol)) for n in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dt", "duflxc_dt")}
    out.update({n: np.zeros((nlay, ncol)) for n in ("hr", "hrc")})
    icld_io = ctypes.c_int(icld)
    L.rrtmg_lw_mcica_wrapper(_i(ncol), _i(nlay), ctypes.byref(icld_io), _i(idrv), *[_p(st[k]) for k in LW_PRE],
                             _i(2), _i(1), _i(1), _p(mc["cldf"]), _p(mc["tauc"]), _p(mc["ciwp"]), _p(mc["clwp"]), _p(rei), _p(rel),
                             _p(st["tauaer"]), *[_p(out[n]) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc", "duflx_dt",
                                                                      "duflxc_dt")])
    return out, mc


def test_lw_two_step_mcica_matches_reference_golden():
    L = _lib()
    _init(L, "lw")
    g = H.golden()
    np.random.seed(0)
    seed = int(np.random.randint(0, 2 ** 31 - 1))     # the draw of lw/component.py:415-423 under the harness's seed(0)
    st = H.default_lw_abi_state(28, 50)
    st["cldfr"][16:19] = 0.5
    st["cicewp"][16:19] = 0.3 * 1e3
    out, mc = _lw_two_step(L, st, 1, 1, seed)
    for name, k in (("upwelling_longwave_flux_in_air", "uflx"), ("downwelling_longwave_flux_in_air", "dflx"),
                    ("downwelling_longwave_flux_in_air_assuming_clear_sky", "dflxc"),
                    ("air_temperature_tendency_from_longwave", "hr")):
        np.testing.assert_allclose(out[k], g[f"TestRRTMGLongwaveMCICA-3d/diag/{name}"].reshape(-1, 50), rtol=0, atol=1e-8)
    assert 0.3 < mc["cldf"][16:19].mean() < 0.7 and mc["cldf"][:16].max() == 0.0


@pytest.mark.parametrize("icld", [1, 2, 3])
def test_lw_two_step_mcica_kissvec_matches_oracle(icld):
    from oracle.rrtmg import lw_mcica
    L = _lib()
    _init(L, "lw")
    st = SY.make_lw_state(300, 45, seed=60 + icld, clouds=True, aerosol=True)
    ref = lw_mcica(H.lw_oracle(cloud_overlap=icld), st, 33, irng=0)
    out, _ = _lw_two_step(L, st, icld, 0, 33)
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(out[k], ref[k]) < 1e-6, (k, H.rel_err(out[k], ref[k]))


def test_lw_two_step_refuses_what_the_bit_mask_cannot_hold():
    L = _lib()
    _init(L, "lw")
    st = SY.make_lw_state(40, 30, seed=2, clouds=True)
    nlay, ncol = st["play"].shape
    st = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}
    mc = {k: np.zeros((nlay, ncol, 140)) for k in ("cldf", "ciwp", "clwp", "tauc")}
    mc["cldf"][12, :, :70] = 0.5          # a fractional sub-column
    mc["ciwp"][12] = 20.0
    out = {n: np.zeros((nlay + 1, ncol)) for n in ("uflx", "dflx", "uflxc", "dflxc")}
    out.update({n: np.zeros((nlay, ncol)) for n in ("hr", "hrc")})
    dummy = np.zeros((ncol, 1))
    L.rrtmg_lw_mcica_wrapper(_i(ncol), _i(nlay), _i(1), _i(0), *[_p(st[k]) for k in LW_PRE], _i(2), _i(1), _i(1),
                             _p(mc["cldf"]), _p(mc["tauc"]), _p(mc["ciwp"]), _p(mc["clwp"]), _p(st["reice"]), _p(st["reliq"]),
                             _p(st["tauaer"]), *[_p(out[n]) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")],
                             _p(dummy), _p(dummy))
    assert all(np.isnan(v).all() for v in out.values())          # never stale, never silently wrong
    assert b"0 or 1" in L.cb200_global_error()


def test_sw_two_step_mcica_matches_reference_golden():
    L = _lib()
    _init(L, "sw")
    g = H.golden()
    st = H.default_sw_abi_state(15, 6)
    st["cldfr"][10:12] = 0.5
    st["cicewp"][10:12] = 0.3e3
    np.random.seed(0)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    nlay, ncol = st["play"].shape
    st = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}
    mc = {k: np.zeros((nlay, ncol, 112)) for k in ("cldf", "ciwp", "clwp", "tauc", "ssac", "asmc", "fsfc")}
    rei, rel = np.zeros((nlay, ncol)), np.zeros((nlay, ncol))
    irng_io = ctypes.c_int(1)
    L.mcica_subcol_sw_wrapper(_i(1), _i(ncol), _i(nlay), _i(1), _i(seed), ctypes.byref(irng_io),
                              *[_p(st[k]) for k in ("play", "cldfr", "cicewp", "cliqwp", "reice", "reliq", "taucld", "ssacld",
                                                    "asmcld", "fsfcld")],
                              _p(mc["cldf"]), _p(mc["ciwp"]), _p(mc["clwp"]), _p(rei), _p(rel),
                              _p(mc["tauc"]), _p(mc["ssac"]), _p(mc["asmc"]), _p(mc["fsfc"]))
    out = {n: np.zeros((nlay + 1, ncol)) for n in ("uflx", "dflx", "uflxc", "dflxc")}
    out.update({n: np.zeros((nlay, ncol)) for n in ("hr", "hrc")})
    bnd, ind = np.ones(14), np.ones(2)
    L.rrtmg_sw_mcica_wrapper(
        _i(ncol), _i(nlay), _i(1), _i(0), *[_p(st[k]) for k in SW_PRE], _d(1.0), _i(1), _d(1367.0), _i(0), _i(2), _i(1), _i(1),
        _p(mc["cldf"]), _p(mc["tauc"]), _p(mc["ssac"]), _p(mc["asmc"]), _p(mc["fsfc"]), _p(mc["ciwp"]), _p(mc["clwp"]), _p(rei),
        _p(rel), *[_p(st[k]) for k in ("tauaer", "ssaaer", "asmaer", "ecaer")],
        *[_p(out[n]) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")], _p(bnd), _p(ind), _d(0.0))
    for name, k in (("upwelling_shortwave_flux_in_air", "uflx"), ("downwelling_shortwave_flux_in_air", "dflx"),
                    ("air_temperature_tendency_from_shortwave", "hr")):
        np.testing.assert_allclose(out[k], g[f"TestRRTMGShortwaveMCICA-3d/diag/{name}"].reshape(-1, 6), rtol=0, atol=1e-8)


# ---- idrv = 1 through the reference-named symbol and through the component ---------------------------------------------
@pytest.mark.parametrize("icld", [1, 2])
def test_nomcica_wrapper_fills_flux_derivatives(icld):
    L = _lib()
    _init(L, "lw")
    ncol, nlay = 700, 40     # two host chunks
    st = SY.make_lw_state(ncol, nlay, seed=70 + icld, clouds=True, aerosol=True, emis_range=(0.9, 1.0))
    a = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}
    out = {n: np.full((nlay + 1, ncol), -1.0) for n in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dt", "duflxc_dt")}
    out.update({n: np.zeros((nlay, ncol)) for n in ("hr", "hrc")})
    L.rrtmg_lw_nomcica_wrapper(_i(ncol), _i(nlay), _i(icld), _i(1), *[_p(a[k]) for k in LW_PRE], _i(2), _i(1), _i(1),
                               *[_p(a[k]) for k in ("cldfr", "taucld", "cicewp", "cliqwp", "reice", "reliq", "tauaer")],
                               *[_p(out[n]) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc", "duflx_dt", "duflxc_dt")])
    ref = H.run_lw_oracle(H.lw_oracle(cloud_overlap=icld, idrv=1), st)
    for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dt", "duflxc_dt"):
        assert H.rel_err(out[k], ref[k]) < 1e-6, (k, H.rel_err(out[k], ref[k]))


def test_component_with_calculate_change_up_flux():
    from climt_b200.rrtmg_lw import RRTMGLongwave
    from climt_b200 import state as S
    st = S.default_rrtmg_lw_state(30, 4)
    raw = dict(st)
    raw["air_pressure"] = st["air_pressure"] / 100.0
    raw["air_pressure_on_interface_levels"] = st["air_pressure_on_interface_levels"] / 100.0
    raw["mass_content_of_cloud_ice_in_atmosphere_layer"] = st["mass_content_of_cloud_ice_in_atmosphere_layer"] * 1e3
    raw["mass_content_of_cloud_liquid_water_in_atmosphere_layer"] = st["mass_content_of_cloud_liquid_water_in_atmosphere_layer"] * 1e3
    comp = RRTMGLongwave(calculate_change_up_flux=True)
    tend, diag = comp.array_call(raw)
    g = H.golden()   # the fluxes themselves are unchanged by the option
    np.testing.assert_allclose(diag["upwelling_longwave_flux_in_air"][:, 0],
                               g["TestRRTMGLongwave-column/diag/upwelling_longwave_flux_in_air"].reshape(-1), rtol=0, atol=1e-8)
    d = comp.change_up_flux["duflx_dt"]
    assert d.shape == (31, 4) and 5.5 < d[0, 0] < 6.5     # 4 sigma T^3 at 300 K = 6.12 W m-2 K-1
    assert np.all(np.diff(d[:, 0]) < 0)
