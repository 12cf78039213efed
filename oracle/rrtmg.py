"""ctypes front-end of the C++ RRTMG restatement (oracle/rrtmg_lw_oracle.cpp).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    so = os.path.join(_HERE, "liborc_rrtmg.so")
    srcs = [os.path.join(_HERE, f) for f in ("rrtmg_lw_oracle.cpp", "rrtmg_sw_oracle.cpp", "ftn.hpp", "mcica_gen.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_last_error.restype = ctypes.c_char_p
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp)


def _c(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


class LWOracle:
    """Mirrors _rrtmg_lw.pyx: set_constants + initialise_rrtm_radiation + rrtm_calculate_longwave_fluxes."""

    def __init__(self, constants, raw_blob, cloud_overlap=1, idrv=0, inflag=2, iceflag=1, liqflag=1):
        L = lib()
        c = constants
        L.orc_lw_set_constants.argtypes = [ctypes.c_double] * 10
        L.orc_lw_set_constants(c["pi"], c["grav"], c["planck"], c["boltz"], c["clight"], c["avogad"],
                               c["alosmt"], c["gascon"], c["sbcnst"], c["secdy"])
        L.orc_lw_ini.argtypes = [ctypes.c_char_p, ctypes.c_double]
        if L.orc_lw_ini(raw_blob.encode(), c["cpdair"]):
            raise RuntimeError(L.orc_last_error().decode())
        self.flags = dict(icld=cloud_overlap, idrv=idrv, inflag=inflag, iceflag=iceflag, liqflag=liqflag)

    def exp_tables(self):
        t = [np.zeros(10001) for _ in range(3)]
        lib().orc_lw_get_exp_tables(*[_p(x) for x in t])
        return t

    def reduced(self, band, name):
        L = lib()
        L.orc_lw_get_reduced.argtypes = [ctypes.c_int, ctypes.c_char_p, _dp, ctypes.c_int64]
        n = L.orc_lw_get_reduced(band, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n)
        L.orc_lw_get_reduced(band, name.encode(), _p(out), n)
        return out

    def __call__(self, play, plev, tlay, tlev, tsfc, h2o, o3, co2, ch4, n2o, o2, cfc11, cfc12, cfc22, ccl4, emis,
                 cldfr, tauaer, taucld, cicewp, cliqwp, reice, reliq, debug=False):
        """Arrays as the Cython shim takes them: (nlay, ncol) [(nlay+1, ncol) for interfaces], emis (16, ncol),
        tauaer (16, nlay, ncol), taucld (nlay, ncol, 16); pressures in hPa."""
        nlay, ncol = play.shape
        a = dict(play=_c(play), plev=_c(plev, (nlay + 1, ncol)), tlay=_c(tlay, (nlay, ncol)),
                 tlev=_c(tlev, (nlay + 1, ncol)), tsfc=_c(tsfc, (ncol,)))
        gases = [_c(x, (nlay, ncol)) for x in (h2o, o3, co2, ch4, n2o, o2, cfc11, cfc12, cfc22, ccl4)]
        emis = _c(emis, (16, ncol)); cldfr = _c(cldfr, (nlay, ncol)); taucld = _c(taucld, (nlay, ncol, 16))
        cl = [_c(x, (nlay, ncol)) for x in (cicewp, cliqwp, reice, reliq)]
        tauaer = _c(tauaer, (16, nlay, ncol))
        out = {k: np.zeros((nlay + 1, ncol)) for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dt", "duflxc_dt")}
        out.update({k: np.zeros((nlay, ncol)) for k in ("hr", "hrc")})
        dbg_t = np.zeros((140, nlay, ncol)) if debug else None
        dbg_f = np.zeros((140, nlay, ncol)) if debug else None
        icld = ctypes.c_int(self.flags["icld"])
        L = lib()
        f = self.flags
        rc = L.orc_lw_nomcica(
            ctypes.c_int(ncol), ctypes.c_int(nlay), ctypes.byref(icld), ctypes.c_int(f["idrv"]),
            _p(a["play"]), _p(a["plev"]), _p(a["tlay"]), _p(a["tlev"]), _p(a["tsfc"]),
            *[_p(g) for g in gases], _p(emis),
            ctypes.c_int(f["inflag"]), ctypes.c_int(f["iceflag"]), ctypes.c_int(f["liqflag"]),
            _p(cldfr), _p(taucld), *[_p(x) for x in cl], _p(tauaer),
            _p(out["uflx"]), _p(out["dflx"]), _p(out["hr"]), _p(out["uflxc"]), _p(out["dflxc"]), _p(out["hrc"]),
            _p(out["duflx_dt"]), _p(out["duflxc_dt"]),
            _p(dbg_t) if debug else None, _p(dbg_f) if debug else None)
        if rc:
            raise RuntimeError(L.orc_last_error().decode())
        if debug:
            out["taug"], out["fracs"] = dbg_t, dbg_f
        return out


SW_ARRAYS = ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "asdir", "asdif", "aldir",
             "aldif", "coszen", "cldfr", "taucld", "ssacld", "asmcld", "fsfcld", "cicewp", "cliqwp", "reice", "reliq",
             "tauaer", "ssaaer", "asmaer", "ecaer")


class SWOracle:
    """Mirrors _rrtmg_sw.pyx: set_constants + initialise_rrtm_radiation + rrtm_calculate_shortwave_fluxes."""

    def __init__(self, constants, raw_blob, cloud_overlap=1, iaer=0, inflag=2, iceflag=1, liqflag=1, isolvar=0,
                 scon=1367.0, indsolvar=(1.0, 1.0), bndsolvar=None):
        L = lib()
        c = constants
        L.orc_sw_last_error.restype = ctypes.c_char_p
        L.orc_sw_set_constants.argtypes = [ctypes.c_double] * 10
        L.orc_sw_set_constants(c["pi"], c["grav"], c["planck"], c["boltz"], c["clight"], c["avogad"],
                               c["alosmt"], c["gascon"], c["sbcnst"], c["secdy"])
        L.orc_sw_ini.argtypes = [ctypes.c_char_p, ctypes.c_double]
        if L.orc_sw_ini(raw_blob.encode(), c["cpdair"]):
            raise RuntimeError(L.orc_sw_last_error().decode())
        self.flags = dict(icld=cloud_overlap, iaer=iaer, inflag=inflag, iceflag=iceflag, liqflag=liqflag,
                          isolvar=isolvar, scon=scon)
        self.indsolvar = np.array(indsolvar, dtype=np.float64)
        self.bndsolvar = np.ones(16) if bndsolvar is None else np.asarray(bndsolvar, dtype=np.float64)

    def reduced(self, band, name):
        L = lib()
        L.orc_sw_get_reduced.argtypes = [ctypes.c_int, ctypes.c_char_p, _dp, ctypes.c_int64]
        n = L.orc_sw_get_reduced(band, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n)
        L.orc_sw_get_reduced(band, name.encode(), _p(out), n)
        return out

    def __call__(self, st, adjes=1.0, dyofyr=1, solcycfrac=0.0):
        """st: dict with SW_ARRAYS in the C-ABI layout ((nlay, ncol); cloud (nlay, ncol, 14); aerosol (14, nlay, ncol);
        ecaer (6, nlay, ncol))."""
        nlay, ncol = st["play"].shape
        shapes = {k: (nlay, ncol) for k in SW_ARRAYS}
        shapes.update(plev=(nlay + 1, ncol), tlev=(nlay + 1, ncol), tsfc=(ncol,), asdir=(ncol,), asdif=(ncol,),
                      aldir=(ncol,), aldif=(ncol,), coszen=(ncol,), taucld=(nlay, ncol, 14), ssacld=(nlay, ncol, 14),
                      asmcld=(nlay, ncol, 14), fsfcld=(nlay, ncol, 14), tauaer=(14, nlay, ncol),
                      ssaaer=(14, nlay, ncol), asmaer=(14, nlay, ncol), ecaer=(6, nlay, ncol))
        a = {k: _c(st[k], shapes[k]) for k in SW_ARRAYS}
        out = {k: np.zeros((nlay + 1, ncol)) for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")}
        out.update({k: np.zeros((nlay, ncol)) for k in ("swhr", "swhrc")})
        f = self.flags
        icld, iaer = ctypes.c_int(f["icld"]), ctypes.c_int(f["iaer"])
        ind = self.indsolvar.copy()
        L = lib()
        pre = [a[k] for k in ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "asdir",
                              "asdif", "aldir", "aldif", "coszen")]
        cl = [a[k] for k in ("cldfr", "taucld", "ssacld", "asmcld", "fsfcld", "cicewp", "cliqwp", "reice", "reliq",
                             "tauaer", "ssaaer", "asmaer", "ecaer")]
        rc = L.orc_sw_nomcica(
            ctypes.c_int(ncol), ctypes.c_int(nlay), ctypes.byref(icld), ctypes.byref(iaer), *[_p(x) for x in pre],
            ctypes.c_double(adjes), ctypes.c_int(dyofyr), ctypes.c_double(f["scon"]), ctypes.c_int(f["isolvar"]),
            ctypes.c_int(f["inflag"]), ctypes.c_int(f["iceflag"]), ctypes.c_int(f["liqflag"]), *[_p(x) for x in cl],
            _p(out["swuflx"]), _p(out["swdflx"]), _p(out["swhr"]), _p(out["swuflxc"]), _p(out["swdflxc"]),
            _p(out["swhrc"]), _p(self.bndsolvar), _p(ind), ctypes.c_double(solcycfrac))
        if rc:
            raise RuntimeError(L.orc_sw_last_error().decode())
        return out


def lw_mcica(oracle, st, permuteseed, irng=1, return_mask=False):
    """McICA longwave through the oracle (mcica_subcol_lw + rrtmg_lw mcica).  `oracle` is an initialised LWOracle
    (its flags give icld / inflag / iceflag / liqflag); st uses the oracle's short names."""
    nlay, ncol = st["play"].shape
    f = oracle.flags
    order = ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "cfc11", "cfc12", "cfc22",
             "ccl4", "emis")
    a = [_c(st[k]) for k in order]
    cl = [_c(st[k]) for k in ("cldfr", "taucld", "cicewp", "cliqwp", "reice", "reliq", "tauaer")]
    out = {k: np.zeros((nlay + 1, ncol)) for k in ("uflx", "dflx", "uflxc", "dflxc")}
    out.update({k: np.zeros((nlay, ncol)) for k in ("hr", "hrc")})
    mask = np.zeros((nlay, ncol, 140)) if return_mask else None
    icld = ctypes.c_int(f["icld"])
    L = lib()
    rc = L.orc_lw_mcica(ctypes.c_int(ncol), ctypes.c_int(nlay), ctypes.byref(icld), ctypes.c_int(int(permuteseed)),
                        ctypes.c_int(irng), *[_p(x) for x in a], ctypes.c_int(f["inflag"]), ctypes.c_int(f["iceflag"]),
                        ctypes.c_int(f["liqflag"]), *[_p(x) for x in cl],
                        _p(out["uflx"]), _p(out["dflx"]), _p(out["hr"]), _p(out["uflxc"]), _p(out["dflxc"]), _p(out["hrc"]),
                        _p(mask) if return_mask else None)
    if rc:
        raise RuntimeError(L.orc_last_error().decode())
    if return_mask:
        out["mask"] = mask
    return out


def sw_mcica(oracle, st, permuteseed, irng=1, adjes=1.0, dyofyr=1, solcycfrac=0.0):
    """McICA shortwave through the oracle (mcica_subcol_sw + rrtmg_sw mcica); `oracle` is an initialised SWOracle."""
    nlay, ncol = st["play"].shape
    f = oracle.flags
    a = {k: _c(st[k]) for k in SW_ARRAYS}
    out = {k: np.zeros((nlay + 1, ncol)) for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")}
    out.update({k: np.zeros((nlay, ncol)) for k in ("swhr", "swhrc")})
    icld, iaer = ctypes.c_int(f["icld"]), ctypes.c_int(f["iaer"])
    ind = oracle.indsolvar.copy()
    L = lib()
    pre = [a[k] for k in ("play", "plev", "tlay", "tlev", "tsfc", "h2o", "o3", "co2", "ch4", "n2o", "o2", "asdir",
                          "asdif", "aldir", "aldif", "coszen")]
    cl = [a[k] for k in ("cldfr", "taucld", "ssacld", "asmcld", "fsfcld", "cicewp", "cliqwp", "reice", "reliq",
                         "tauaer", "ssaaer", "asmaer", "ecaer")]
    rc = L.orc_sw_mcica(
        ctypes.c_int(ncol), ctypes.c_int(nlay), ctypes.byref(icld), ctypes.byref(iaer), ctypes.c_int(int(permuteseed)),
        ctypes.c_int(irng), *[_p(x) for x in pre], ctypes.c_double(adjes), ctypes.c_int(dyofyr),
        ctypes.c_double(f["scon"]), ctypes.c_int(f["isolvar"]), ctypes.c_int(f["inflag"]), ctypes.c_int(f["iceflag"]),
        ctypes.c_int(f["liqflag"]), *[_p(x) for x in cl], _p(out["swuflx"]), _p(out["swdflx"]), _p(out["swhr"]),
        _p(out["swuflxc"]), _p(out["swdflxc"]), _p(out["swhrc"]), _p(oracle.bndsolvar), _p(ind),
        ctypes.c_double(solcycfrac))
    if rc:
        raise RuntimeError(L.orc_sw_last_error().decode())
    return out
