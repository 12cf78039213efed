"""CORK options the engine rejected in round 1, pinned by goldens produced by RUNNING the reference's own components
(tests/golden/make_esft_golden.py):
  * ESFT overlap (cork/optics/correlated_k.py:343-409,564-594): a two-gas and a three-gas per-gas table, and the reference's own
    2-band fixtures with the overlap overridden (one gas: must reduce to the additive path, tests/test_cork_optics.py:220-243);
  * diagnostics_level 1 / 2 (cork/lw/component.py:189-202,341-358; cork/sw/component.py:455-492) on the ESFT tables and on the
    shipped earth tables.
CPU: the ESFT expansion (climt_b200.cork.expand_esft_table) + the oracle against the goldens.
GPU: the drop-in components against the goldens."""
import os

import numpy as np
import pytest

import helpers as H
from climt_b200 import cork
from oracle import cork as OC

RTOL = 1e-9
LW_MAP = {"upwelling_longwave_flux_in_air": "up_broad", "downwelling_longwave_flux_in_air": "down_broad",
          "upwelling_longwave_flux_in_air_per_band": "up_band", "downwelling_longwave_flux_in_air_per_band": "down_band",
          "longwave_optical_depth_per_band": "tau_band", "longwave_transmittance_per_band": "trans_band",
          "air_temperature_tendency_from_longwave_per_band": "hr_band"}
SW_MAP = {"upwelling_shortwave_flux_in_air": "up_broad", "downwelling_shortwave_flux_in_air": "down_broad",
          "upwelling_shortwave_flux_in_air_per_band": "up_band", "downwelling_shortwave_flux_in_air_per_band": "down_band",
          "shortwave_optical_depth_per_band": "tau_band", "air_temperature_tendency_from_shortwave_per_band": "hr_band"}
ESFT_CASES = ("esft2", "esft3", "esft1_2band")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(H.HERE, "golden", "esft_diag_reference.npz"))


@pytest.fixture(scope="module")
def tables():
    z = np.load(os.path.join(H.HERE, "golden", "esft_tables.npz"))

    def get(name):
        pre = name + "/"
        return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    return get


def _inputs(z, case):
    pre = f"{case}/in/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def test_esft_weights_are_the_reference_products():
    w = np.array([[0.3, 0.7], [0.25, 0.75]])
    c = cork.esft_weights(w, 2)
    assert c.shape == (2, 4)
    # combined index = g_gas0 + 2 * g_gas1 (correlated_k.py:362-370)
    np.testing.assert_array_equal(c[0], [0.3 * 0.3, 0.7 * 0.3, 0.3 * 0.7, 0.7 * 0.7])
    np.testing.assert_allclose(c.sum(axis=1), 1.0, rtol=1e-14)
    np.testing.assert_array_equal(cork.esft_weights(w, 1), w)


def test_esft_expansion_maps_digits_to_gases(tables):
    t = tables("esft2_lw")
    x = cork.expand_esft_table(t)
    k, kx = t["k_coefficients"], x["k_coefficients"]
    ngas, nband, ngpt = k.shape[:3]
    assert kx.shape[:3] == (ngas, nband, ngpt ** ngas) and str(x["overlap_method"]) == "additive" and "continuum_kappa" not in x
    for idx in (0, 1, 5, 11, 15):
        np.testing.assert_array_equal(kx[0, :, idx], k[0, :, idx % ngpt])
        np.testing.assert_array_equal(kx[1, :, idx], k[1, :, (idx // ngpt) % ngpt])
        np.testing.assert_array_equal(x["planck_fraction"][:, idx], t["planck_fraction"][:, idx % ngpt])


@pytest.mark.parametrize("case", ESFT_CASES)
def test_oracle_on_the_expanded_table_matches_the_reference_esft_path(gold, tables, case):
    s = _inputs(gold, case)
    s["q"] = s["h2o"]
    lw = OC.lw_call(cork.expand_esft_table(tables(case + "_lw")), s, H.CORK_G, H.CORK_CPD, H.CORK_SIGMA)
    for name, key in LW_MAP.items():
        ref = gold[f"{case}/lw0/{name}"]
        got = lw[key] if lw[key].ndim == 2 else np.moveaxis(lw[key], 0, -1)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12 * float(np.abs(ref).max()), err_msg=f"{case} lw {name}")
    sw = OC.sw_call(cork.expand_esft_table(tables(case + "_sw")), s, H.CORK_G, H.CORK_CPD)
    for name, key in SW_MAP.items():
        ref = gold[f"{case}/sw0/{name}"]
        got = sw[key] if sw[key].ndim == 2 else np.moveaxis(sw[key], 0, -1)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12 * float(np.abs(ref).max()), err_msg=f"{case} sw {name}")


def _engine_arrays(table, s, which):
    """what the components hand to the engine (cork.py _gas_arrays; cork/lw/component.py:243-287)"""
    names, has_h2o, has_co2, fully, bg = cork.table_flags(table)
    a = H.cork_arrays(dict(s, q=s["h2o"]), which)
    a.pop("q_h2o"); a.pop("co2_vmr", None)
    if fully:
        return a
    if bg:
        a["q_h2o"] = s["h2o"]
        if has_co2 and which == "lw":
            a["co2_vmr"] = s["co2"]
        return a
    a["gas_q"] = np.stack([s[g] if g == "h2o" else s[g] * (cork.MOLAR_MASS.get(g, cork.MOLAR_MASS_DRY_AIR) / cork.MOLAR_MASS_DRY_AIR)
                           for g in names])
    return a


@pytest.mark.parametrize("case,lwt,swt", [("esft2", None, None), ("esft3", None, None), ("diag_earth", "earth_low_res_lw", "earth_low_res_sw")])
def test_kernel_code_diagnostics_match_the_reference(gold, tables, case, lwt, swt):
    """The engine's per-thread code (host emulation, NaN-poisoned diagnostics buffers) at diagnostics_level 2 against the
    reference components' extra diagnostics."""
    s = _inputs(gold, case)
    for which, tname in (("lw", lwt), ("sw", swt)):
        table = cork.load_k_table(tname) if tname else tables(f"{case}_{which}")
        esf = float(s["earth_sun_factor"].reshape(-1)[0])
        out, fields = H.run_cork_emul(table, which, _engine_arrays(table, s, which), 1.66 if which == "lw" else esf, diagnostics_level=2)
        pre = f"{case}/{which}2/"
        fscale = float(np.abs(gold[pre + ("upwelling_longwave_flux_in_air" if which == "lw" else "downwelling_shortwave_flux_in_air")]).max())
        assert fields and all(np.isfinite(v).all() for v in fields.values())
        for name, got in fields.items():
            ref = gold[pre + name]
            got = np.moveaxis(got, 0, -1)
            atol = RTOL * fscale if ("per_gpoint" in name or "direct_beam" in name) else 1e-13
            np.testing.assert_allclose(got, ref, rtol=RTOL, atol=atol, err_msg=f"{case} {which} {name}")


def _compare_component(gold, case, which, level, comp, s):
    tend, diag = comp.array_call(dict(s))
    pre = f"{case}/{which}{level}/"
    names = [k[len(pre):] for k in gold.files if k.startswith(pre)]
    assert set(names) == set(diag) | {"T"}, (sorted(set(names) ^ (set(diag) | {"T"})))
    flux = gold[pre + ("upwelling_longwave_flux_in_air" if which == "lw" else "downwelling_shortwave_flux_in_air")]
    fscale = float(np.abs(flux).max())
    dp_w = np.abs(np.diff(s["p_int"], axis=0)) * H.CORK_CPD / H.CORK_G
    for name in names:
        ref = gold[pre + name]
        got = np.asarray(tend["T"] if name == "T" else diag[name])
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        if name == "T" or "tendency" in name:
            w = dp_w if ref.ndim == 2 else dp_w[..., None]
            w = w / 86400.0 if "tendency" in name else w
            assert float(np.abs((got - ref) * w).max()) <= RTOL * fscale, (case, which, name)
        elif "flux" in name or "per_gpoint" in name or "direct_beam" in name:
            np.testing.assert_allclose(got, ref, rtol=RTOL, atol=RTOL * fscale, err_msg=f"{case} {which}{level} {name}")
        else:
            np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-13, err_msg=f"{case} {which}{level} {name}")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ESFT_CASES)
def test_esft_components_match_the_reference(gold, tables, case):
    s = _inputs(gold, case)
    levels = (0,) if case == "esft1_2band" else (0, 1, 2)
    for level in levels:
        lw = cork.CorkLongwaveRadiation(optics="correlated_k", table=tables(case + "_lw"), diagnostics_level=level)
        _compare_component(gold, case, "lw", level, lw, s)
        sw = cork.CorkShortwaveRadiation(optics="correlated_k", table=tables(case + "_sw"), diagnostics_level=level)
        _compare_component(gold, case, "sw", level, sw, s)


@pytest.mark.gpu
@pytest.mark.parametrize("level", (1, 2))
def test_diagnostics_levels_on_the_earth_tables_match_the_reference(gold, level):
    s = _inputs(gold, "diag_earth")
    lw = cork.CorkLongwaveRadiation(optics="correlated_k", table="earth_low_res_lw", diagnostics_level=level)
    _compare_component(gold, "diag_earth", "lw", level, lw, s)
    sw = cork.CorkShortwaveRadiation(optics="correlated_k", table="earth_low_res_sw", diagnostics_level=level)
    _compare_component(gold, "diag_earth", "sw", level, sw, s)
