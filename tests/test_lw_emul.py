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


# ---- maximum-random overlap without McICA: rtrnmr (rrtmg_lw_rtrnmr.f90).  No golden exists for it in the reference, so the
# restatement is pinned by (a) agreement of the two independently structured implementations and (b) its physics: it must
# reduce to the golden-pinned rtrn when no two cloudy layers touch, and to a single thicker cloud when equal fractions stack.
@pytest.mark.parametrize("icld", [2, 3])
def test_maximum_random_overlap_kernels_match_oracle(icld):
    st = SY.make_lw_state(40, 60, seed=31 + icld, clouds=True, aerosol=True)
    ref = H.run_lw_oracle(H.lw_oracle(cloud_overlap=icld), st)
    rc, got = H.run_lw_emul(st, flags=(icld, 0, 2, 1, 1))
    assert rc == 0
    for k in ("uflx", "dflx", "uflxc", "dflxc"):
        assert H.rel_err(got[k], ref[k]) < TOL, k
    rnd = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), st)
    np.testing.assert_array_equal(ref["uflxc"], rnd["uflxc"])          # clear-sky stream does not know about overlap
    assert np.abs(ref["dflx"] - rnd["dflx"]).max() > 1.0                 # ... the total-sky one does (W m-2)


def test_maximum_random_overlap_reduces_to_random_for_isolated_cloud_layers():
    st = SY.make_lw_state(16, 40, seed=3, clouds=True)
    rng = np.random.default_rng(0)
    st["cldfr"][:] = 0.0
    for l in (8, 12, 17, 25):                                             # no two cloudy layers adjacent
        st["cldfr"][l, :] = rng.uniform(0.1, 0.9, 16)
    st["cicewp"][:] = rng.uniform(5, 30, st["cicewp"].shape)
    st["cliqwp"][:] = rng.uniform(5, 30, st["cliqwp"].shape)
    rnd = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), st)              # rtrn: pinned by the reference goldens
    for impl in ("oracle", "kernel"):
        mr = H.run_lw_oracle(H.lw_oracle(cloud_overlap=2), st) if impl == "oracle" else H.run_lw_emul(st, flags=(2, 0, 2, 1, 1))[1]
        for k in ("uflx", "dflx"):
            # rtrn folds the cloud into an effective fraction with exp(), rtrnmr splits the radiance with the tabulated
            # transmittance: same physics, ~5e-6 apart
            assert H.rel_err(mr[k], rnd[k]) < 2e-5, (impl, k, H.rel_err(mr[k], rnd[k]))


def test_maximum_overlap_of_equal_fractions_beats_random_overlap_cloud_cover():
    st = SY.make_lw_state(8, 40, seed=5, clouds=False)
    for k in ("cicewp", "cliqwp"):
        st[k][:] = 20.0
    st["reice"][:] = 40.0
    st["reliq"][:] = 10.0
    st["cldfr"][10:13, :] = 0.4                                           # three stacked layers, same fraction
    rnd = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), st)
    rc, mr = H.run_lw_emul(st, flags=(2, 0, 2, 1, 1))
    assert rc == 0
    # maximum overlap: total cover 0.4; random: 1 - 0.6^3 = 0.78 -> less downward longwave at the surface, more OLR
    assert np.all(mr["dflx"][0] < rnd["dflx"][0] - 5.0) and np.all(mr["uflx"][-1] > rnd["uflx"][-1])
    assert np.all(mr["dflx"][0] > mr["dflxc"][0])


# ---- idrv = 1: d(upward flux)/d(surface temperature) (rrtmg_lw_rtrn.f90:279-296,458-512,546-555).  The reference holds no golden
# for it (its component cannot even request it, _rrtmg_lw.pyx:164-165), so the restatement is pinned by what the quantity IS:
# a centred finite difference of the oracle's own upward flux with respect to the surface temperature.
def test_oracle_flux_derivative_is_the_finite_difference_of_its_upward_flux():
    st = SY.make_lw_state(6, 40, seed=21, clouds=True, aerosol=True, emis_range=(0.85, 1.0))
    o = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1, idrv=1), st)
    h = 0.05
    up, dn = dict(st), dict(st)
    up["tsfc"], dn["tsfc"] = st["tsfc"] + h, st["tsfc"] - h
    # the interface temperature at the surface is a separate input (tlev[0]); only tbound moves, as in the Fortran's derivative
    fu = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), up)
    fd = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), dn)
    for k, kk in (("duflx_dt", "uflx"), ("duflxc_dt", "uflxc")):
        fdiff = (fu[kk] - fd[kk]) / (2 * h)
        # 1 %: the flux uses a Planck table that is piecewise linear in 1-K steps (its finite difference is the step's mean slope)
        # while the derivative interpolates a table of true derivatives (totplnkderiv) -- they differ by the curvature over < 1 K
        assert np.max(np.abs(o[k] - fdiff)) < 1e-2 * np.max(np.abs(fdiff)), k
    assert o["duflx_dt"][0].min() > 3.0          # ~ 4 sigma T^3 at the surface
    assert np.all(np.diff(o["duflxc_dt"], axis=0) <= 1e-12)   # the clear-sky derivative can only be attenuated upwards


@pytest.mark.parametrize("icld", [1, 2])
def test_flux_derivative_kernels_match_oracle(icld):
    st = SY.make_lw_state(9, 33, seed=30 + icld, clouds=True, aerosol=True, emis_range=(0.9, 1.0))
    ref = H.run_lw_oracle(H.lw_oracle(cloud_overlap=icld, idrv=1), st)
    rc, got = H.run_lw_emul(st, (icld, 1, 2, 1, 1))
    assert rc == 0
    for k in ("uflx", "dflx", "uflxc", "dflxc", "duflx_dt", "duflxc_dt"):
        assert H.rel_err(got[k], ref[k]) < TOL, k
    assert np.abs(got["duflx_dt"] - got["duflxc_dt"]).max() > 1e-3


def test_mcica_flux_derivative_with_overcast_layers_equals_the_deterministic_one():
    """rtrnmc's derivative (rrtmg_lw_rtrnmc.f90:447-510) has no oracle restatement; with every cloudy layer overcast all
    sub-columns are cloudy, so it must reproduce rtrn's (cloud fraction 1) exactly."""
    st = SY.make_lw_state(7, 30, seed=44, clouds=True)
    st["cldfr"] = np.where(st["cldfr"] > 0, 1.0, 0.0)
    rc0, det = H.run_lw_emul(st, (1, 1, 2, 1, 1))
    rc1, mc = H.run_lw_emul(st, (1, 1, 2, 1, 1), mcica=(1, 1, 3))
    assert rc0 == 0 and rc1 == 0
    for k in ("uflx", "duflx_dt", "duflxc_dt"):
        assert H.rel_err(mc[k], det[k]) < 1e-12, k
