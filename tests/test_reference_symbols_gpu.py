"""The reference's own C symbols (INTEGRATION.md option A): libclimt_b200.so called exactly the way climt's Cython shims call
the Fortran wrappers -- every argument by pointer, module-global configuration, numpy buffers in, results in the caller's
arrays (climt/_components/rrtmg/lw/_rrtmg_lw.pyx:19-80,174-212; sw/_rrtmg_sw.pyx:22-105)."""
import ctypes

import numpy as np
import pytest

import helpers as H
from climt_b200 import _native, constants as C, synthetic as SY

pytestmark = pytest.mark.gpu
_dp = ctypes.POINTER(ctypes.c_double)


def _d(x):
    return ctypes.byref(ctypes.c_double(float(x)))


def _i(x):
    return ctypes.byref(ctypes.c_int(int(x)))


def _p(a):
    return a.ctypes.data_as(_dp)


def _set_constants(fn):
    k = C.rrtmg_constants()
    fn(*[_d(k[n]) for n in ("pi", "grav", "planck", "boltz", "clight", "avogad", "alosmt", "gascon", "sbcnst", "secdy")])
    return k


def test_lw_link_swap_symbols_match_oracle():
    L = _native.lib()
    for f in (L.rrtmg_set_constants, L.rrtmg_lw_ini_wrapper, L.rrtmg_lw_nomcica_wrapper):
        f.restype = None
        f.argtypes = None
    from climt_b200 import rrtmg_tables
    rrtmg_tables.lw_blob_path()  # the wrapper reads the table blob next to the library (written by build())
    k = _set_constants(L.rrtmg_set_constants)
    L.rrtmg_lw_ini_wrapper(_d(k["cpdair"]))
    ncol, nlay = 300, 45
    st = SY.make_lw_state(ncol, nlay, seed=77, clouds=True, aerosol=True)
    a = {n: np.ascontiguousarray(v) for n, v in H.to_abi(st).items()}
    out = {n: np.zeros((nlay + 1, ncol)) for n in ("uflx", "dflx", "uflxc", "dflxc")}
    out.update({n: np.zeros((nlay, ncol)) for n in ("hr", "hrc")})
    dummy = np.zeros((ncol, 1))
    order = ("play", "plev", "tlay", "tlev", "tsfc", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr", "cfc11vmr",
             "cfc12vmr", "cfc22vmr", "ccl4vmr", "emis")
    L.rrtmg_lw_nomcica_wrapper(_i(ncol), _i(nlay), _i(1), _i(0), *[_p(a[n]) for n in order], _i(2), _i(1), _i(1),
                               *[_p(a[n]) for n in ("cldfr", "taucld", "cicewp", "cliqwp", "reice", "reliq", "tauaer")],
                               *[_p(out[n]) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")], _p(dummy), _p(dummy))
    ref = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), st)
    for n in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(out[n], ref[n]) < 1e-6, n


def test_sw_link_swap_symbols_match_oracle():
    L = _native.lib()
    for f in (L.rrtmg_sw_set_constants, L.rrtmg_sw_ini_wrapper, L.rrtmg_sw_nomcica_wrapper):
        f.restype = None
        f.argtypes = None
    from climt_b200 import rrtmg_tables
    rrtmg_tables.sw_blob_path()
    k = _set_constants(L.rrtmg_sw_set_constants)
    L.rrtmg_sw_ini_wrapper(_d(k["cpdair"]))
    ncol, nlay = 260, 40
    st = SY.make_sw_state(ncol, nlay, seed=78, clouds=True)
    a = {n: np.ascontiguousarray(v) for n, v in H.to_abi_sw(st).items()}
    out = {n: np.zeros((nlay + 1, ncol)) for n in ("uflx", "dflx", "uflxc", "dflxc")}
    out.update({n: np.zeros((nlay, ncol)) for n in ("hr", "hrc")})
    bnd, ind = np.ones(14), np.ones(2)
    L.rrtmg_sw_nomcica_wrapper(
        _i(ncol), _i(nlay), _i(1), _i(0),
        *[_p(a[n]) for n in ("play", "plev", "tlay", "tlev", "tsfc", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
                             "asdir", "asdif", "aldir", "aldif", "coszen")],
        _d(1.0), _i(120), _d(1367.0), _i(0), _i(2), _i(1), _i(1),
        *[_p(a[n]) for n in ("cldfr", "taucld", "ssacld", "asmcld", "fsfcld", "cicewp", "cliqwp", "reice", "reliq", "tauaer",
                             "ssaaer", "asmaer", "ecaer")],
        *[_p(out[n]) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")], _p(bnd), _p(ind), _d(0.0))
    ref = H.sw_oracle()(st, adjes=1.0, dyofyr=120, solcycfrac=0.0)
    for n, nn in (("uflx", "swuflx"), ("dflx", "swdflx"), ("uflxc", "swuflxc"), ("dflxc", "swdflxc")):
        assert H.rel_err(out[n], ref[nn]) < 1e-6, n
