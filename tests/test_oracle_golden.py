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
