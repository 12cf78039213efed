"""The reference's two-step McICA C ABI, the parts that need no GPU.

* every symbol the unchanged `_rrtmg_lw.pyx` / `_rrtmg_sw.pyx` declare `extern` resolves in libclimt_b200.so (a Cython extension
  linked against it would otherwise fail at import: CPython dlopens with RTLD_NOW) -- fixture tests/golden/pyx_externs.json,
  produced from the reference by tests/golden/make_pyx_externs.py and re-derived here when the reference tree is present;
* `mcica_subcol_{lw,sw}_wrapper` (host-side generators, rrtmg_lw_c_binder.f90:50-92 / rrtmg_sw_c_binder.f90:59-107) fill the
  (ngpt, ncol, nlay) arrays exactly like the oracle's restatement of generate_stochastic_clouds (oracle/mcica_gen.hpp), for both
  random-number generators and all three overlap assumptions.
"""
import ctypes
import json
import os

import numpy as np
import pytest

import helpers as H
from climt_b200 import _native, synthetic as SY

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)


def _i(x):
    return ctypes.byref(ctypes.c_int(int(x)))


def _p(a):
    return a.ctypes.data_as(_dp)


def test_every_pyx_extern_resolves_in_the_library():
    fx = json.load(open(os.path.join(HERE, "golden", "pyx_externs.json")))
    L = _native.lib()
    for which in ("lw", "sw"):
        assert len(fx[which]["symbols"]) == 5
        for name in fx[which]["symbols"]:
            assert hasattr(L, name), f"{name} ({fx[which]['file']}) is not exported by libclimt_b200.so"


def test_pyx_extern_fixture_is_current():
    ref = os.environ.get("CLIMT_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_pyx_externs", os.path.join(HERE, "golden", "make_pyx_externs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    fx = json.load(open(os.path.join(HERE, "golden", "pyx_externs.json")))
    for which, rel in m.FILES.items():
        assert m.externs(os.path.join(ref, rel)) == fx[which]["symbols"]


@pytest.mark.parametrize("icld,irng", [(1, 1), (2, 1), (3, 1), (1, 0), (2, 0), (3, 0)])
def test_lw_subcolumn_generator_matches_oracle(icld, irng):
    from oracle.rrtmg import lw_mcica
    ncol, nlay, seed = 37, 33, 91
    st = SY.make_lw_state(ncol, nlay, seed=3 + icld, clouds=True)
    st["taucld"] = np.random.default_rng(2).uniform(0.1, 3.0, (nlay, ncol, 16))
    ref = lw_mcica(H.lw_oracle(cloud_overlap=icld), st, seed, irng=irng, return_mask=True)["mask"]  # (nlay, ncol, 140)
    L = _native.lib()
    L.mcica_subcol_lw_wrapper.restype = None
    L.mcica_subcol_lw_wrapper.argtypes = None
    big = [np.full((nlay, ncol, 140), -7.0) for _ in range(4)]   # cldfmcl, ciwpmcl, clwpmcl, taucmcl
    rei, rel = np.full((nlay, ncol), -7.0), np.full((nlay, ncol), -7.0)
    irng_io = ctypes.c_int(5 if irng else 0)   # anything but 0 -> 1 (mcica_subcol_gen_lw.f90:303)
    L.mcica_subcol_lw_wrapper(_i(1), _i(ncol), _i(nlay), _i(icld), _i(seed), ctypes.byref(irng_io),
                              *[_p(st[k]) for k in ("play", "cldfr", "cicewp", "cliqwp", "reice", "reliq", "taucld")],
                              _p(big[0]), _p(big[1]), _p(big[2]), _p(rei), _p(rel), _p(big[3]))
    assert irng_io.value == (1 if irng else 0)
    np.testing.assert_array_equal(big[0], ref)
    assert 0.02 < big[0].mean() < 0.9
    np.testing.assert_array_equal(big[1], ref * st["cicewp"][:, :, None])
    np.testing.assert_array_equal(big[2], ref * st["cliqwp"][:, :, None])
    ngb = np.repeat(np.arange(16), [10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2])
    np.testing.assert_array_equal(big[3], ref * st["taucld"][:, :, ngb])
    np.testing.assert_array_equal(rei, st["reice"])
    np.testing.assert_array_equal(rel, st["reliq"])


def test_sw_subcolumn_generator_fills_clear_values():
    ncol, nlay, seed = 21, 26, 5
    st = SY.make_sw_state(ncol, nlay, seed=8, clouds=True)
    rng = np.random.default_rng(3)
    for k, (lo, hi) in (("taucld", (0.1, 4.0)), ("ssacld", (0.5, 0.99)), ("asmcld", (0.6, 0.9)), ("fsfcld", (0.3, 0.8))):
        st[k] = rng.uniform(lo, hi, (nlay, ncol, 14))
    L = _native.lib()
    L.mcica_subcol_sw_wrapper.restype = None
    L.mcica_subcol_sw_wrapper.argtypes = None
    big = [np.full((nlay, ncol, 112), -7.0) for _ in range(7)]   # cldf, ciwp, clwp, tauc, ssac, asmc, fsfc
    rei, rel = np.zeros((nlay, ncol)), np.zeros((nlay, ncol))
    irng_io = ctypes.c_int(1)
    L.mcica_subcol_sw_wrapper(_i(1), _i(ncol), _i(nlay), _i(2), _i(seed), ctypes.byref(irng_io),
                              *[_p(st[k]) for k in ("play", "cldfr", "cicewp", "cliqwp", "reice", "reliq", "taucld", "ssacld",
                                                    "asmcld", "fsfcld")],
                              _p(big[0]), _p(big[1]), _p(big[2]), _p(rei), _p(rel), *[_p(b) for b in big[3:]])
    m = big[0]
    assert set(np.unique(m)) == {0.0, 1.0}
    # (the mask itself is pinned by the reference's ShortwaveMCICA golden through the two-step call, tests/test_mcica_symbols_gpu.py)
    ngb = np.repeat(np.arange(14), [6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12])
    np.testing.assert_array_equal(big[1], m * st["cicewp"][:, :, None])
    np.testing.assert_array_equal(big[3], m * st["taucld"][:, :, ngb])
    np.testing.assert_array_equal(big[4], np.where(m == 1.0, st["ssacld"][:, :, ngb], 1.0))   # clear: ssa 1
    np.testing.assert_array_equal(big[5], m * st["asmcld"][:, :, ngb])                        # clear: asm 0
    np.testing.assert_array_equal(big[6], m * st["fsfcld"][:, :, ngb])
    # layers without cloud stay clear in every sub-column; a layer with fraction f is cloudy in about f of them
    cf = st["cldfr"]
    assert np.all(m[cf < 1e-20] == 0.0)
    sel = cf > 0.3
    assert abs(m[sel].mean() - cf[sel].mean()) < 0.05


def test_subcolumn_generator_returns_untouched_outputs_for_clear_sky_flag():
    L = _native.lib()
    L.mcica_subcol_lw_wrapper.restype = None
    L.mcica_subcol_lw_wrapper.argtypes = None
    ncol, nlay = 4, 6
    z2 = np.zeros((nlay, ncol))
    tc = np.zeros((nlay, ncol, 16))
    big = [np.full((nlay, ncol, 140), 3.0) for _ in range(4)]
    rei, rel = np.full((nlay, ncol), 3.0), np.full((nlay, ncol), 3.0)
    L.mcica_subcol_lw_wrapper(_i(1), _i(ncol), _i(nlay), _i(0), _i(1), _i(1), _p(z2 + 500.0), _p(z2), _p(z2), _p(z2), _p(z2),
                              _p(z2), _p(tc), _p(big[0]), _p(big[1]), _p(big[2]), _p(rei), _p(rel), _p(big[3]))
    assert all(np.all(b == 3.0) for b in big) and np.all(rei == 3.0)   # mcica_subcol_gen_lw.f90:119 `if (icld.eq.0) return`
