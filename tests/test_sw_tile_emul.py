"""The column-tile form of the shortwave transfer (sw_core.cuh: sw_tile_cell + sw_tile_sweeps, CUDA kernel k_sw_tile): the same code
stepped serially on the CPU over NaN-poisoned row buffers (tests/emul/sw_emul.cpp: run_tile), against the oracle, the reference's
goldens and the unit form it replaces."""
import numpy as np
import pytest

import helpers as H
from climt_b200 import synthetic as SY


@pytest.mark.parametrize("tile", [1, 2])
@pytest.mark.parametrize("mode", ["clear", "clear_icld0", "clouds", "aerosol", "ecmwf"])
def test_tile_form_matches_oracle_and_unit_form(mode, tile):
    iaer = {"aerosol": 10, "ecmwf": 6}.get(mode, 0)
    icld = 0 if mode == "clear_icld0" else 1
    st = SY.make_sw_state(20, 45, seed=19, clouds=(mode == "clouds"), aerosol=(iaer == 10), ecmwf=(iaer == 6))
    ref = H.sw_oracle(iaer=iaer)(st, adjes=1.0, dyofyr=200, solcycfrac=0.3)
    args = ((icld, iaer, 2, 1, 1, 0, 200), [1.0, 1367.0, 0.3, 1.0, 1.0] + [1.0] * 14)
    rc, got = H.run_sw_emul(st, *args, tile=tile)
    rc2, unit = H.run_sw_emul(st, *args, tile=0)
    assert rc == 0 and rc2 == 0
    for k, kk in H.SW_KEYS.items():
        assert np.isfinite(got[k]).all()
        if k.startswith("hr"):
            np.testing.assert_allclose(got[k], ref[kk], rtol=1e-7, atol=1e-9)
        else:
            assert H.rel_err(got[k], ref[kk]) < 1e-10, k
            assert H.rel_err(got[k], unit[k]) < 1e-12, k
    if mode != "clouds":
        np.testing.assert_array_equal(got["uflx"], got["uflxc"])


def test_tile_form_mcica_matches_reference_golden():
    g = H.golden()
    st = H.default_sw_abi_state(15, 6)
    st["cldfr"][10:12] = 0.5
    st["cicewp"][10:12] = 0.3e3
    np.random.seed(0)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    rc, e = H.run_sw_emul(st, (1, 0, 2, 1, 1, 0, 1), mcica=(1, 1, seed), tile=True)
    assert rc == 0
    for name, k in (("upwelling_shortwave_flux_in_air", "uflx"), ("downwelling_shortwave_flux_in_air", "dflx"),
                    ("air_temperature_tendency_from_shortwave", "hr")):
        ref = g[f"TestRRTMGShortwaveMCICA-3d/diag/{name}"].reshape(-1, 6)
        np.testing.assert_allclose(e[k], ref, rtol=0, atol=1e-8)


@pytest.mark.parametrize("nlay", [7, 8, 63, 64, 72, 100])
def test_scan_form_at_awkward_layer_counts(nlay):
    """interfaces per lane = ceil((nlay + 1) / 8): blocks that end exactly at, before and after the top of the atmosphere"""
    st = SY.make_sw_state(6, nlay, seed=40 + nlay, clouds=True)
    args = ((1, 0, 2, 1, 1, 0, 120), None)
    rc, got = H.run_sw_emul(st, *args, tile=2)
    rc2, unit = H.run_sw_emul(st, *args, tile=0)
    assert rc == 0 and rc2 == 0
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert np.isfinite(got[k]).all()
        assert H.rel_err(got[k], unit[k]) < 1e-12, k


@pytest.mark.parametrize("icld,irng", [(1, 0), (2, 1), (3, 0)])
def test_tile_form_mcica_matches_oracle(icld, irng):
    from oracle.rrtmg import sw_mcica
    st = SY.make_sw_state(8, 33, seed=3 + icld, clouds=True, overcast_only=False, aerosol=True)
    ref = sw_mcica(H.sw_oracle(cloud_overlap=icld, iaer=10), st, 55, irng=irng, dyofyr=150)
    rc, got = H.run_sw_emul(st, (icld, 10, 2, 1, 1, 0, 150), mcica=(1, irng, 55), tile=1 + (icld % 2))
    assert rc == 0
    for k, kk in H.SW_KEYS.items():
        if not k.startswith("hr"):
            assert H.rel_err(got[k], ref[kk]) < 1e-10, k
