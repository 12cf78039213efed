"""The column-tile form of the longwave transfer (lw_core.cuh: lw_tile_cell + lw_tile_sweeps, CUDA kernel k_lw_tile): the same
code stepped serially on the CPU over NaN-poisoned row buffers (tests/emul/lw_emul.cpp: run_tile), against the oracle, the
reference's goldens and the unit form it replaces."""
import numpy as np
import pytest

import helpers as H
from climt_b200 import synthetic as SY

TOL = 1e-10


@pytest.mark.parametrize("clouds,icld,nlay", [(False, 0, 60), (False, 1, 45), (True, 1, 33), (True, 1, 72)])
def test_tile_form_matches_oracle_and_unit_form(clouds, icld, nlay):
    st = SY.make_lw_state(24, nlay, seed=11 + nlay, clouds=clouds, trace=True, aerosol=True, emis_range=(0.85, 1.0))
    ref = H.run_lw_oracle(H.lw_oracle(cloud_overlap=icld), st)
    rc, got = H.run_lw_emul(st, (icld, 0, 2, 1, 1), tile=True)
    rc2, unit = H.run_lw_emul(st, (icld, 0, 2, 1, 1), tile=False)
    assert rc == 0 and rc2 == 0
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert np.isfinite(got[k]).all()
        assert H.rel_err(got[k], ref[k]) < TOL, k
        assert H.rel_err(got[k], unit[k]) < 1e-12, k
    for k in ("hr", "hrc"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-7, atol=1e-9)
    if not clouds:
        np.testing.assert_array_equal(got["uflx"], got["uflxc"])


def test_tile_form_matches_golden_default_state():
    g = H.golden()
    st = H.default_lw_abi_state(30, 1)
    rc, got = H.run_lw_emul(st, tile=True)
    assert rc == 0
    ref = g["TestRRTMGLongwave-column/diag/upwelling_longwave_flux_in_air"][:, 0, 0]
    np.testing.assert_allclose(got["uflx"][:, 0], ref, rtol=0, atol=1e-8)


def test_tile_form_mcica_matches_reference_golden():
    """TestRRTMGLongwaveMCICA-3d through the tile form (see tests/test_lw_emul.py for the state)."""
    g = H.golden()
    np.random.seed(0)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    st = H.default_lw_abi_state(28, 50)
    st["cldfr"][16:19] = 0.5
    st["cicewp"][16:19] = 0.3 * 1e3
    rc, got = H.run_lw_emul(st, (1, 0, 2, 1, 1), mcica=(1, 1, seed), tile=True)
    assert rc == 0
    for name, k in (("upwelling_longwave_flux_in_air", "uflx"), ("downwelling_longwave_flux_in_air", "dflx"),
                    ("downwelling_longwave_flux_in_air_assuming_clear_sky", "dflxc"),
                    ("air_temperature_tendency_from_longwave", "hr")):
        ref = g[f"TestRRTMGLongwaveMCICA-3d/diag/{name}"].reshape(-1, 50)
        np.testing.assert_allclose(got[k], ref, rtol=0, atol=1e-8)


@pytest.mark.parametrize("icld,irng", [(1, 0), (2, 0), (3, 1)])
def test_tile_form_mcica_matches_oracle(icld, irng):
    from oracle.rrtmg import lw_mcica
    st = SY.make_lw_state(10, 36, seed=5 + icld, clouds=True, aerosol=True)
    ref = lw_mcica(H.lw_oracle(cloud_overlap=icld), st, 77, irng=irng)
    rc, got = H.run_lw_emul(st, (icld, 0, 2, 1, 1), mcica=(1, irng, 77), tile=True)
    assert rc == 0
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(got[k], ref[k]) < TOL, k
