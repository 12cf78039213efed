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
    srcs = [os.path.join(_HERE, f) for f in ("rrtmg_lw_oracle.cpp", "ftn.hpp")]
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
