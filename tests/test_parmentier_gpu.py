"""CORK picket-fence optics (optics="parmentier") on the GPU, through the drop-in components:
`CorkLongwaveRadiation(optics="parmentier").array_call(state)` / `CorkShortwaveRadiation(...)` against golden vectors
produced by the reference's own component classes (tests/golden/make_parmentier_golden.py), and the engines against the
oracle at a size the reference's scalar Python loops cannot reach.  Tolerance 1e-6 relative (BASELINE.json); observed ~1e-13."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
RTOL = 1e-6
CASES = ("clear", "cloudy_feedback", "cold")


@pytest.fixture(scope="module")
def gold():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return np.load(H.PARMENTIER_GOLDEN)


def _compare(diag, tend, z, case, which):
    ref_flux = z[f"{case}/{which}/{'upwelling_longwave_flux_in_air' if which == 'lw' else 'downwelling_shortwave_flux_in_air'}"]
    fscale = max(float(np.max(np.abs(ref_flux))), 1e-300)
    names = [k.split("/")[-1] for k in z.files if k.startswith(f"{case}/{which}/")]
    assert set(names) == set(diag) | {"T"}
    for name in names:
        ref = z[f"{case}/{which}/{name}"]
        got = tend["T"] if name == "T" else diag[name]
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        if name == "T" or "tendency" in name:
            s = float(np.max(np.abs(ref))) or 1.0
            assert H.flux_scaled_err(got, ref, s) < 1e-6, (case, which, name)
        elif "flux" in name:
            np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-9 * fscale, err_msg=f"{case} {which} {name}")
        else:
            np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-300, err_msg=f"{case} {which} {name}")


@pytest.mark.parametrize("case", CASES)
def test_drop_in_components_match_the_reference_components(gold, case):
    from climt_b200 import cork
    s = H.parmentier_case(gold, case)
    lw = cork.CorkLongwaveRadiation(diffusivity_factor=float(s["diffusivity"]))   # optics="parmentier" is the default
    assert lw.num_longwave_bands == 2 and "irradiation_temperature" in lw.input_properties
    tend, diag = lw.array_call(dict(s))
    _compare(diag, tend, gold, case, "lw")
    sw = cork.CorkShortwaveRadiation(optics="parmentier", bond_albedo_feedback=bool(s["bond_albedo_feedback"]))
    assert sw.num_shortwave_bands == 3 and "internal_temperature" in sw.input_properties
    tend, diag = sw.array_call(dict(s))
    _compare(diag, tend, gold, case, "sw")
    assert np.all(diag["upwelling_shortwave_flux_in_air"][:, 2] == 0.0)  # night column


def test_three_dimensional_state_and_unirradiated_fallback(gold):
    """(lev, lat, lon) state: outputs keep the horizontal shape; with T_irr = 0 everywhere the solar flux falls back to the
    stellar spectrum integrated over three equal bands (cork/sw/component.py:36-62, 252-256)."""
    from climt_b200 import cork
    s = H.parmentier_case(gold, "clear")
    nlev, ncol = s["T"].shape
    ny, nx = 6, ncol // 6

    def r3(a, lead):
        return a[..., :ny * nx].reshape(lead + (ny, nx)) if a.ndim == len(lead) + 1 else a

    s3 = {"T": r3(s["T"], (nlev,)), "p": r3(s["p"], (nlev,)), "p_int": r3(s["p_int"], (nlev + 1,)), "T_surf": r3(s["T_surf"], ()),
          "T_irr": np.zeros((ny, nx)), "T_int": r3(s["T_int"], ()), "zenith": r3(s["zenith"], ()), "albedo": r3(s["albedo"], ()),
          "earth_sun_factor": r3(s["earth_sun_factor"], ()),
          "tau_cloud_sw": np.zeros((nlev, ny, nx, 3)), "ssa_cloud": np.zeros((nlev, ny, nx, 3)), "g_cloud": np.zeros((nlev, ny, nx, 3))}
    sw = cork.CorkShortwaveRadiation()
    tend, diag = sw.array_call(s3)
    assert tend["T"].shape == (nlev, ny, nx)
    assert diag["upwelling_shortwave_flux_in_air_per_band"].shape == (nlev + 1, ny, nx, 3)
    mu0 = np.cos(s3["zenith"])
    day = mu0 > 1e-4
    toa_down = diag["downwelling_shortwave_flux_in_air"][-1]
    expect = sw._solar_flux_per_band.sum() * float(s3["earth_sun_factor"].reshape(-1)[0]) * mu0
    np.testing.assert_allclose(toa_down[day], expect[day], rtol=1e-12)
    assert np.all(toa_down[~day] == 0.0)
    assert abs(sw._solar_flux_per_band.sum() - 1361.0) < 30.0   # the shipped solar spectrum integrates to about one solar constant


def test_engines_match_the_oracle_at_scale(gold):
    """20 480 columns x 60 levels (10 host-pipeline chunks): engine == oracle; device-pointer call == host call bit for bit."""
    import torch
    import sys
    sys.path.insert(0, H.os.path.join(H.HERE, "golden"))
    import make_parmentier_golden as MG
    from climt_b200 import cork
    from oracle import parmentier as OP
    ncol, nlev = 20480, 60
    s = MG.make_state(ncol, nlev, seed=5, clouds=True)
    co, fr = H.picket_coefficients()
    for which in ("lw", "sw"):
        eng = cork.CorkEngine(None, g=H.CORK_G, cpd=H.CORK_CPD, sigma=H.CORK_SIGMA, picket=(which, co, fr))
        arrays = H.picket_arrays(s, which)
        if which == "lw":
            ref = OP.lw_call(co, fr, s, H.CORK_G, H.CORK_CPD, H.CORK_SIGMA)
            got = eng.lw_host(ncol, nlev, arrays)
        else:
            ref = OP.sw_call(co, fr, s, H.CORK_G, H.CORK_CPD, H.CORK_SIGMA)
            got = eng.sw_host(ncol, nlev, arrays, solar_flux=ref["solar_flux"])
        fscale = float(np.max(np.abs(ref["down_broad"])))
        for k in ("up_broad", "down_broad", "up_band", "down_band"):
            np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=1e-9 * fscale, err_msg=f"{which} {k}")
        np.testing.assert_allclose(got["tau_band"], ref["tau_band"], rtol=RTOL, atol=1e-300)
        ins, outs = eng.shapes(ncol, nlev)
        dev_in = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).cuda() for k, v in arrays.items()}
        dev_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in got}
        if which == "lw":
            eng.lw_device(ncol, nlev, dev_in, dev_out)
        else:
            eng.sw_device(ncol, nlev, dev_in, dev_out, solar_flux=ref["solar_flux"])
        torch.cuda.synchronize()
        for k in got:
            assert np.array_equal(dev_out[k].cpu().numpy(), got[k]), (which, k)
        assert eng.last_launches >= 4
        eng.close()


def test_missing_inputs_are_rejected(gold):
    from climt_b200 import cork
    co, fr = H.picket_coefficients()
    s = H.parmentier_case(gold, "clear")
    nlev, ncol = s["T"].shape
    eng = cork.CorkEngine(None, picket=("sw", co, fr))
    a = H.picket_arrays(s, "sw")
    with pytest.raises(ValueError):
        eng.sw_host(ncol, nlev, {k: v for k, v in a.items() if k != "T_irr"}, solar_flux=np.ones((3, 1)))
    with pytest.raises(ValueError):
        eng.lw_host(ncol, nlev, H.picket_arrays(s, "lw"))   # a visible-band engine has no longwave path
    eng.close()
