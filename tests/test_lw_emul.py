"""Host logic of the kernels: the very per-thread code the CUDA kernels run (climt_b200/csrc/lw_core.cuh),
stepped serially on the CPU by tests/emul/lw_emul.cpp, against the oracle.  This checks the algorithm and
the generic band evaluator without a GPU; the -m gpu tests check the compiled sm_100a kernels."""
import numpy as np
import pytest

import helpers as H
from climt_b200 import synthetic as SY

TOL = 1e-10   # two independent fp64 implementations; the acceptance bar for the product is 1e-6


@pytest.mark.parametrize("clouds,trace,nlay", [(False, True, 60), (True, True, 33), (False, False, 72)])
def test_emulated_kernels_match_oracle(clouds, trace, nlay):
    st = SY.make_lw_state(24, nlay, seed=7 + nlay, clouds=clouds, trace=trace, aerosol=True, emis_range=(0.85, 1.0))
    ref = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), st)
    rc, got = H.run_lw_emul(st)
    assert rc == 0
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(got[k], ref[k]) < TOL, k
    for k in ("hr", "hrc"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-7, atol=1e-9)
    if not clouds:
        np.testing.assert_array_equal(got["uflx"], got["uflxc"])


def test_emulated_kernels_match_golden_default_state():
    g = H.golden()
    st = H.default_lw_abi_state(30, 1)
    rc, got = H.run_lw_emul(st)
    assert rc == 0
    ref = g["TestRRTMGLongwave-column/diag/upwelling_longwave_flux_in_air"][:, 0, 0]
    np.testing.assert_allclose(got["uflx"][:, 0], ref, rtol=0, atol=1e-8)


def test_high_co2_and_n2o_trigger_adjusted_columns():
    """Minor-gas 'too abundant' branches (taumol.f90:529-535, 1336, 1464, 1702, 1860, 2478)."""
    st = SY.make_lw_state(8, 40, seed=3)
    st["co2"][:] = 2000e-6
    st["n2o"][:] = 1.5e-6
    ref = H.run_lw_oracle(H.lw_oracle(), st)
    rc, got = H.run_lw_emul(st)
    assert rc == 0
    assert H.rel_err(got["dflx"], ref["dflx"]) < TOL and H.rel_err(got["uflx"], ref["uflx"]) < TOL


def test_cloud_radius_out_of_bounds_is_reported_not_fatal():
    st = SY.make_lw_state(4, 30, seed=5, clouds=True)
    st["cldfr"][10, :] = 0.5
    st["cicewp"][10, :] = 10.0
    st["reice"][10, :] = 500.0
    rc, _ = H.run_lw_emul(st)
    assert rc == 2     # 'ICE RADIUS OUT OF BOUNDS' where the Fortran would `stop`


def test_mcica_emulated_kernels_match_reference_golden():
    """TestRRTMGLongwaveMCICA-3d: 10x5 columns, 28 levels, cloud fraction 0.5 / ice 0.3 kg m-2 in layers 16:19,
    Mersenne twister seeded like the reference harness (np.random.seed(0); randint(0, 2**31-1)), tests/test_components.py:148,440-452."""
    g = H.golden()
    np.random.seed(0)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    st = H.default_lw_abi_state(28, 50)
    st["cldfr"][16:19] = 0.5
    st["cicewp"][16:19] = 0.3 * 1e3
    rc, got = H.run_lw_emul(st, (1, 0, 2, 1, 1), mcica=(1, 1, seed))
    assert rc == 0
    for name, k in (("upwelling_longwave_flux_in_air", "uflx"), ("downwelling_longwave_flux_in_air", "dflx"),
                    ("downwelling_longwave_flux_in_air_assuming_clear_sky", "dflxc"),
                    ("air_temperature_tendency_from_longwave", "hr")):
        ref = g[f"TestRRTMGLongwaveMCICA-3d/diag/{name}"].reshape(-1, 50)
        np.testing.assert_allclose(got[k], ref, rtol=0, atol=1e-8)
    # the masks differ from column to column: a direct pin on the MT stream order
    assert np.ptp(got["dflx"][0]) > 10.0


@pytest.mark.parametrize("icld,irng", [(1, 0), (2, 0), (3, 0), (1, 1), (2, 1), (3, 1)])
def test_mcica_emulated_kernels_match_oracle(icld, irng):
    from oracle.rrtmg import lw_mcica
    st = SY.make_lw_state(10, 36, seed=5 + icld, clouds=True, aerosol=True)
    ref = lw_mcica(H.lw_oracle(cloud_overlap=icld), st, 77, irng=irng)
    rc, got = H.run_lw_emul(st, (icld, 0, 2, 1, 1), mcica=(1, irng, 77))
    assert rc == 0
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(got[k], ref[k]) < TOL, k
    assert np.abs(got["uflx"] - got["uflxc"]).max() > 1e-3     # clouds do something


def test_oracle_mcica_golden():
    from oracle.rrtmg import lw_mcica
    g = H.golden()
    np.random.seed(0)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    st = H.default_lw_abi_state(28, 50)
    st["cldfr"][16:19] = 0.5
    st["cicewp"][16:19] = 0.3 * 1e3
    o = lw_mcica(H.lw_oracle(cloud_overlap=1), st, seed, irng=1)
    for name, k in (("downwelling_longwave_flux_in_air", "dflx"), ("air_temperature_tendency_from_longwave", "hr")):
        np.testing.assert_allclose(o[k], g[f"TestRRTMGLongwaveMCICA-3d/diag/{name}"].reshape(-1, 50), rtol=0, atol=1e-8)
