"""The oracle is pinned by the reference's own golden outputs (tests/golden/reference_caches.npz, made by
tests/golden/make_golden.py from /root/reference/tests/cached_component_output).  The reference's own
tolerance is atol 1e-8 (tests/test_components.py:355-356); the restatement reproduces them to ~1e-13."""
import numpy as np
import pytest

import helpers as H

CASES = [
    ("TestRRTMGLongwave", "column", 30, {}, False),
    ("TestRRTMGLongwave", "3d", 28, {}, False),
    ("TestRRTMGLongwaveWithClouds", "column", 30, {"inflag": 1}, False),
    ("TestRRTMGLongwaveWithClouds", "3d", 28, {"inflag": 1}, False),
    ("TestRRTMGLongwaveWithExternalInterfaceTemperature", "column", 30, {}, True),
    ("TestRRTMGLongwaveWithExternalInterfaceTemperature", "3d", 28, {}, True),
]
DIAG = {"upwelling_longwave_flux_in_air": "uflx", "downwelling_longwave_flux_in_air": "dflx",
        "upwelling_longwave_flux_in_air_assuming_clear_sky": "uflxc",
        "downwelling_longwave_flux_in_air_assuming_clear_sky": "dflxc",
        "air_temperature_tendency_from_longwave_assuming_clear_sky": "hrc",
        "air_temperature_tendency_from_longwave": "hr"}


@pytest.mark.parametrize("cls,kind,nz,flags,ext", CASES)
def test_lw_oracle_matches_reference_golden(cls, kind, nz, flags, ext):
    g = H.golden()
    st = H.default_lw_abi_state(nz, 1, external_tint=ext)
    out = H.run_lw_oracle(H.lw_oracle(**flags), st)
    for name, key in DIAG.items():
        ref = g[f"{cls}-{kind}/diag/{name}"]
        # every column of the 3d golden is identical (horizontally uniform default state)
        assert np.allclose(ref, ref[:, :1, :1], rtol=0, atol=1e-11)
        np.testing.assert_allclose(out[key][:, 0], ref[:, 0, 0], rtol=0, atol=1e-8)
        assert H.rel_err(out[key][:, 0], ref[:, 0, 0], floor=1e-2) < 1e-10
    ref = g[f"{cls}-{kind}/tend/air_temperature"]
    np.testing.assert_allclose(out["hr"][:, 0], ref[:, 0, 0], rtol=0, atol=1e-8)


def test_crib_values():
    """SURVEY.md appendix B crib."""
    g = H.golden()
    up = g["TestRRTMGLongwave-column/diag/upwelling_longwave_flux_in_air"][:, 0, 0]
    assert abs(up[0] - 459.29431776) < 1e-7 and abs(up[-1] - 443.9333446703) < 1e-7


@pytest.mark.parametrize("cls,builder", [("TestRRTMGLongwave", "default_rrtmg_lw_state"), ("TestRRTMGShortwave", "default_rrtmg_sw_state")])
def test_default_state_inputs_match_the_reference_state(cls, builder):
    """The `*_stepping` caches carry the whole state the reference's get_default_state built for the component (grid pressures,
    ozone profile, gas defaults, cloud and aerosol defaults): the inputs of every golden comparison above, checked one by one."""
    from climt_b200 import state as S
    g = H.golden()
    st = getattr(S, builder)(30, 1)
    pre = f"{cls}-column_stepping/state/"
    checked = 0
    for key in g.files:
        if not key.startswith(pre):
            continue
        name = key[len(pre):]
        if name not in st or name == "air_temperature":      # (the cache's temperature is the stepped one)
            continue
        ref = np.asarray(g[key], dtype=np.float64)
        mine = np.asarray(st[name], dtype=np.float64)
        np.testing.assert_allclose(mine.reshape(-1), ref.reshape(-1), rtol=1e-14, atol=0, err_msg=name)
        checked += 1
    assert checked >= 20, checked
    ak, bk = S.hybrid_sigma_pressure_levels(31, 1.0132e5, 20.0)
    np.testing.assert_allclose(ak, g[pre + "atmosphere_hybrid_sigma_pressure_a_coordinate_on_interface_levels"], rtol=1e-14, atol=1e-12)
    np.testing.assert_allclose(bk, g[pre + "atmosphere_hybrid_sigma_pressure_b_coordinate_on_interface_levels"], rtol=1e-14, atol=1e-16)


def test_one_adams_bashforth_step_of_the_oracle_tendency_matches_the_stepping_cache():
    """TestRRTMGLongwave-column_stepping: AdamsBashforth(RRTMGLongwave) on the default state, dt = 10 s (tests/test_components.py:
    145-152, 241-254).  Its first step is forward Euler on the tendency in K/day: T + 10 s * hr / 86400."""
    g = H.golden()
    out = H.run_lw_oracle(H.lw_oracle(), H.default_lw_abi_state(30, 1))
    ref_T = g["TestRRTMGLongwave-column_stepping/state/air_temperature"][:, 0, 0]
    np.testing.assert_allclose(290.0 + 10.0 * out["hr"][:, 0] / 86400.0, ref_T, rtol=0, atol=2e-13)
    assert np.abs(ref_T - 290.0).max() > 1e-4
    np.testing.assert_allclose(out["uflx"][:, 0], g["TestRRTMGLongwave-column_stepping/diag/upwelling_longwave_flux_in_air"][:, 0, 0], rtol=0, atol=1e-8)
