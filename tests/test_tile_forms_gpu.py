"""The column-tile transfer kernels (k_lw_tile: the default longwave path; k_sw_tile: built, measured, off by default) on the GPU:
against the oracle and against the unit form they replace, clear sky / clouds / McICA, ragged column counts (tiles of 32 and 16
columns, 32-column supertiles choosing between the cloud-free and the cloudy form).  The forms are chosen by environment variables
read when an engine is created."""
import os

import numpy as np
import pytest

import helpers as H
from climt_b200 import synthetic as SY

pytestmark = pytest.mark.gpu


def _subset(st, sub, ncol):
    """the same state restricted to the columns `sub` (the column axis is the one of length ncol)"""
    out = {}
    for k, v in st.items():
        ax = [i for i, n in enumerate(v.shape) if n == ncol][-1]
        idx = [slice(None)] * v.ndim
        idx[ax] = sub
        out[k] = np.ascontiguousarray(v[tuple(idx)])
    return out


def _engine(cls, env, **kw):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return cls(**kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("clouds,ncol,nlay", [(False, 1000, 60), (True, 533, 47), (True, 4100, 72), (True, 300, 100), (False, 200, 150)])
def test_lw_tile_and_unit_forms_agree_with_the_oracle(clouds, ncol, nlay):
    """(100 layers: 9 cells per producer thread; 150 layers: the tile form's shared memory does not fit -> the unit form runs)"""
    from climt_b200.engine import LWEngine
    st = SY.make_lw_state(ncol, nlay, seed=31, clouds=clouds, aerosol=True)
    if clouds:  # a band of cloud-free columns: whole 32-column supertiles take the cloud-free form
        for k in ("cldfr", "cicewp", "cliqwp"):
            st[k][:, 96:224] = 0.0
    abi = H.to_abi(st)
    tile = _engine(LWEngine, {"CLIMT_B200_LW_TILE": "1"}, device=0).run_host(ncol, nlay, abi)
    unit = _engine(LWEngine, {"CLIMT_B200_LW_TILE": "0"}, device=0).run_host(ncol, nlay, abi)
    sub = slice(0, ncol, max(1, ncol // 97))
    ref = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), _subset(st, sub, ncol))
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(tile[k], unit[k]) < 1e-11, k
        assert H.rel_err(tile[k][:, sub], ref[k]) < 1e-9, k
    np.testing.assert_allclose(tile["hr"], unit["hr"], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("icld", [1, 2])
def test_lw_tile_form_mcica_matches_the_unit_form(icld):
    from climt_b200.engine import LWEngine
    ncol, nlay = 2100, 72
    st = SY.make_lw_state(ncol, nlay, seed=7, clouds=True)
    for k in ("cldfr", "cicewp", "cliqwp"):
        st[k][:, 512:700] = 0.0
    abi = H.to_abi(st)
    kw = dict(device=0, icld=icld, mcica=True, irng=0, permuteseed=112)
    tile = _engine(LWEngine, {"CLIMT_B200_LW_TILE": "1"}, **kw).run_host(ncol, nlay, abi)
    unit = _engine(LWEngine, {"CLIMT_B200_LW_TILE": "0"}, **kw).run_host(ncol, nlay, abi)
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(tile[k], unit[k]) < 1e-11, k


@pytest.mark.parametrize("form", ["scan", "tile"])
@pytest.mark.parametrize("clouds,ncol,nlay", [(False, 1000, 60), (True, 533, 47), (True, 2100, 72), (False, 300, 100)])
def test_sw_scan_and_tile_forms_match_the_unit_form_and_the_oracle(clouds, ncol, nlay, form):
    """scan: k_sw_scan (default) takes the cloud-free 32-column supertiles, k_sw_transfer the rest; tile: k_sw_tile (off by default)"""
    from climt_b200.engine import SWEngine
    if form == "tile" and nlay > 62:
        pytest.skip("the cloudy tile form needs 16-column tiles: up to 62 layers")
    st = SY.make_sw_state(ncol, nlay, seed=33, clouds=clouds)
    if clouds:
        for k in ("cldfr", "cicewp", "cliqwp"):
            st[k][:, 96:224] = 0.0
    abi = H.to_abi_sw(st)
    env = {"CLIMT_B200_SW_TILE": "1" if form == "tile" else "0", "CLIMT_B200_SW_SCAN": "1" if form == "scan" else "0"}
    tile = _engine(SWEngine, env, device=0).run_host(ncol, nlay, abi, dyofyr=80)
    unit = _engine(SWEngine, {"CLIMT_B200_SW_TILE": "0", "CLIMT_B200_SW_SCAN": "0"}, device=0).run_host(ncol, nlay, abi, dyofyr=80)
    sub = slice(0, ncol, max(1, ncol // 61))
    ref = H.sw_oracle()(_subset(st, sub, ncol), dyofyr=80)
    for k, kk in (("uflx", "swuflx"), ("dflx", "swdflx"), ("uflxc", "swuflxc"), ("dflxc", "swdflxc")):
        assert np.isfinite(tile[k]).all(), k
        assert H.rel_err(tile[k], unit[k]) < 1e-11, k
        assert H.rel_err(tile[k][:, sub], ref[kk]) < 1e-9, k


def test_sw_scan_form_mcica_mixed_supertiles():
    from climt_b200.engine import SWEngine
    ncol, nlay = 2100, 60
    st = SY.make_sw_state(ncol, nlay, seed=9, clouds=True, overcast_only=False)
    for k in ("cldfr", "cicewp", "cliqwp"):
        st[k][:, 512:700] = 0.0
    abi = H.to_abi_sw(st)
    kw = dict(device=0, icld=2, mcica=True, irng=0, permuteseed=112)
    scan = _engine(SWEngine, {"CLIMT_B200_SW_SCAN": "1"}, **kw).run_host(ncol, nlay, abi, dyofyr=80)
    unit = _engine(SWEngine, {"CLIMT_B200_SW_SCAN": "0"}, **kw).run_host(ncol, nlay, abi, dyofyr=80)
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(scan[k], unit[k]) < 1e-11, k
