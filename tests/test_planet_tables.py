"""Every correlated-k table class the reference ships -- Mars / Venus (H2O axis over a premixed background), Titan and
TRAPPIST-1e hab2 (fully premixed, 5-D), TRAPPIST-1e hab1 (premixed with an H2O axis, "effective" gas), the grey per-gas fixture --
through `optics="correlated_k"`: golden vectors produced by the reference's own component classes on the reference's own NetCDF
tables (tests/golden/make_planet_golden.py), which this package ships converted to its own container (`.cb2k`, same arrays
and dtypes: climt_b200/table_store.py).
CPU: the oracle (table classification + glue of oracle/cork.py, kernels of cork_oracle.cpp) against the goldens.
GPU: the drop-in components against the goldens.  Tolerance 1e-6 relative (BASELINE.json); asserted at 1e-9 of the flux scale
(observed ~1e-11); heating rates are compared as the flux divergence they come from (tendency x dp x cp / g), because a thin
top layer turns a 1e-11 flux difference into a much larger relative one."""
import numpy as np
import pytest

import helpers as H
from climt_b200 import cork
from oracle import cork as OC

CASES = {"mars": ("mars_lw", "mars_sw"), "venus": ("venus_lw", "venus_sw"), "titan": ("titan_lw", "titan_sw"),
         "trappist1e_hab1": ("trappist1e_hab1_lw", "trappist1e_hab1_sw"), "trappist1e_hab2": ("trappist1e_hab2_lw", "trappist1e_hab2_sw"),
         "tour_gray": ("tour_gray_lw", None)}
RTOL = 1e-9
LW_MAP = {"upwelling_longwave_flux_in_air": "up_broad", "downwelling_longwave_flux_in_air": "down_broad",
          "upwelling_longwave_flux_in_air_per_band": "up_band", "downwelling_longwave_flux_in_air_per_band": "down_band",
          "longwave_optical_depth_per_band": "tau_band", "longwave_transmittance_per_band": "trans_band",
          "air_temperature_tendency_from_longwave_per_band": "hr_band"}
SW_MAP = {"upwelling_shortwave_flux_in_air": "up_broad", "downwelling_shortwave_flux_in_air": "down_broad",
          "upwelling_shortwave_flux_in_air_per_band": "up_band", "downwelling_shortwave_flux_in_air_per_band": "down_band",
          "shortwave_optical_depth_per_band": "tau_band", "air_temperature_tendency_from_shortwave_per_band": "hr_band"}


@pytest.fixture(scope="module")
def gold():
    return np.load(H.os.path.join(H.HERE, "golden", "planet_reference.npz"))


def _inputs(z, case):
    pre = f"{case}/in/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def _close(got, ref, name, scale):
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12 * scale, err_msg=name)


def test_shipped_tables_cover_every_class():
    seen = set()
    for lwt, swt in CASES.values():
        for t in (lwt, swt):
            if t is None:
                continue
            tbl = cork.load_k_table(t)
            names, has_h2o, has_co2, fully, bg = cork.table_flags(tbl)
            seen.add((tuple(names), has_h2o, fully, bg))
            assert np.asarray(tbl["k_coefficients"]).dtype == np.float32 and str(tbl["overlap_method"]) == "additive"
    assert {(("h2o",), True, False, True), (("effective",), False, True, False), (("effective",), True, False, True),
            (("co2",), False, False, False)} <= seen


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_the_reference_components(gold, case):
    lwt, swt = CASES[case]
    s = _inputs(gold, case)
    s["q"] = s["h2o"]
    lw = OC.lw_call(cork.load_k_table(lwt), s, H.CORK_G, H.CORK_CPD, H.CORK_SIGMA)
    fscale = float(np.abs(gold[f"{case}/lw/upwelling_longwave_flux_in_air"]).max())
    for name, key in LW_MAP.items():
        ref = gold[f"{case}/lw/{name}"]
        got = lw[key] if lw[key].ndim == 2 else np.moveaxis(lw[key], 0, -1)
        if key == "hr_band":
            _close(got, ref, name, float(np.abs(ref).max()))
        else:
            _close(got, ref, name, fscale if "flux" in name else 0.0)
    _close(lw["heating_rate"], gold[f"{case}/lw/T"], "T", float(np.abs(gold[f"{case}/lw/T"]).max()))
    if swt:
        sw = OC.sw_call(cork.load_k_table(swt), s, H.CORK_G, H.CORK_CPD)
        fscale = float(np.abs(gold[f"{case}/sw/downwelling_shortwave_flux_in_air"]).max())
        for name, key in SW_MAP.items():
            ref = gold[f"{case}/sw/{name}"]
            got = sw[key] if sw[key].ndim == 2 else np.moveaxis(sw[key], 0, -1)
            _close(got, ref, name, float(np.abs(ref).max()) if key == "hr_band" else (fscale if "flux" in name else 0.0))
        _close(sw["heating_rate"], gold[f"{case}/sw/T"], "T", float(np.abs(gold[f"{case}/sw/T"]).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_drop_in_components_match_the_reference_components(gold, case):
    lwt, swt = CASES[case]
    s = _inputs(gold, case)
    comps = [("lw", cork.CorkLongwaveRadiation(optics="correlated_k", table=lwt))]
    if swt:
        comps.append(("sw", cork.CorkShortwaveRadiation(optics="correlated_k", table=swt)))
    for which, comp in comps:
        tend, diag = comp.array_call(dict(s))
        names = [k.split("/")[-1] for k in gold.files if k.startswith(f"{case}/{which}/")]
        assert set(names) == set(diag) | {"T"}
        flux = gold[f"{case}/{which}/{'upwelling_longwave_flux_in_air' if which == 'lw' else 'downwelling_shortwave_flux_in_air'}"]
        fscale = float(np.abs(flux).max())
        dp_w = np.abs(np.diff(s["p_int"], axis=0)) * H.CORK_CPD / H.CORK_G      # heating rate [K/s] x dp_w = flux divergence [W m-2]
        for name in names:
            ref = gold[f"{case}/{which}/{name}"]
            got = np.asarray(tend["T"] if name == "T" else diag[name])
            assert got.shape == ref.shape, (name, got.shape, ref.shape)
            if name == "T" or "tendency" in name:
                w = dp_w if ref.ndim == 2 else dp_w[..., None]
                w = w / 86400.0 if "tendency" in name else w                    # the *_tendency_from_* diagnostics are per day
                assert float(np.abs((got - ref) * w).max()) <= RTOL * fscale, (case, which, name)
            elif "flux" in name:
                np.testing.assert_allclose(got, ref, rtol=RTOL, atol=RTOL * fscale, err_msg=f"{case} {which} {name}")
            else:
                np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-300, err_msg=f"{case} {which} {name}")
