"""The drop-in boundary on the Python side: every drop-in class keeps the reference class's constructor keywords
(with the same defaults) and its input / tendency / diagnostic / output properties -- names, dims, units and aliases -- as read off
the reference's own classes by tests/golden/make_properties_golden.py (tests/golden/reference_properties.json).  Extra
constructor keywords of the drop-ins (`device`, `asynchronous`, `flux_layout`) are additions; nothing may be missing."""
import inspect
import json

import pytest

import helpers as H

REF = json.load(open(H.os.path.join(H.HERE, "golden", "reference_properties.json")))
KINDS = ("input_properties", "tendency_properties", "diagnostic_properties", "output_properties")
ADDED = {"device", "asynchronous", "flux_layout"}


def _cls(name):
    from climt_b200 import (berger_solar_insolation, cork, emanuel, gray, instellation, rrtmg_lw, rrtmg_sw, simple_physics,
                            slab_surface)
    return {"RRTMGLongwave": rrtmg_lw.RRTMGLongwave, "RRTMGShortwave": rrtmg_sw.RRTMGShortwave,
            "GrayLongwaveRadiation": gray.GrayLongwaveRadiation, "EmanuelConvection": emanuel.EmanuelConvection,
            "EmanuelConvectionPython": emanuel.EmanuelConvectionPython,
            "SimplePhysics": simple_physics.SimplePhysics, "Instellation": instellation.Instellation,
            "BergerSolarInsolation": berger_solar_insolation.BergerSolarInsolation, "SlabSurface": slab_surface.SlabSurface,
            "CorkLongwaveRadiation": cork.CorkLongwaveRadiation, "CorkShortwaveRadiation": cork.CorkShortwaveRadiation}[name]


def _norm(p):
    return {name: {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in d.items() if k in ("dims", "units", "alias")}
            for name, d in p.items()}


def _compare(ours, ref, where):
    for kind in KINDS:
        if kind not in ref:
            continue
        mine = _norm(getattr(ours, kind))
        assert set(mine) == set(ref[kind]), (where, kind, sorted(set(mine) ^ set(ref[kind])))
        for name, want in ref[kind].items():
            assert mine[name] == want, (where, kind, name, mine[name], want)


@pytest.mark.parametrize("name", sorted(REF))
def test_constructor_keywords_and_defaults(name):
    sig = inspect.signature(_cls(name).__init__)
    mine = {n: (None if p.default is inspect.Parameter.empty else repr(p.default)) for n, p in sig.parameters.items()
            if n != "self" and p.kind is not inspect.Parameter.VAR_KEYWORD}
    want = REF[name]["init"]
    assert set(want) <= set(mine), (name, "missing keywords", sorted(set(want) - set(mine)))
    assert set(mine) - set(want) <= ADDED, (name, "unexpected keywords", sorted(set(mine) - set(want) - ADDED))
    for k, v in want.items():
        assert mine[k] == v, (name, k, mine[k], v)


def test_slab_surface_with_ekman_option():
    _compare(_cls("SlabSurface")(include_ekman=True), REF["SlabSurface"]["instances"]["ekman"], "SlabSurface[ekman]")


@pytest.mark.parametrize("name", [n for n in sorted(REF) if not n.startswith("Cork")])
def test_class_level_properties(name):
    _compare(_cls(name), REF[name], name)


@pytest.mark.parametrize("name,label", [(n, lab) for n in sorted(REF) if n.startswith("Cork") for lab in sorted(REF[n]["instances"])])
def test_cork_instance_properties(monkeypatch, name, label):
    """CORK properties depend on the optics mode and on the table class; constructing a drop-in creates an engine, so the native
    create calls are stubbed out here (no GPU in the CPU suite)"""
    from climt_b200 import cork

    class NoEngine:
        def __init__(self, *a, **k):
            pass
    monkeypatch.setattr(cork, "CorkEngine", NoEngine)
    which = "lw" if "Long" in name else "sw"
    kw = {"parmentier": {"optics": "parmentier"}, "earth": {"optics": "correlated_k", "table": f"earth_low_res_{which}"},
          "mars": {"optics": "correlated_k", "table": f"mars_{which}"}, "titan": {"optics": "correlated_k", "table": f"titan_{which}"},
          "tour_gray": {"optics": "correlated_k", "table": "tour_gray_lw"}}[label]
    _compare(_cls(name)(**kw), REF[name]["instances"][label], f"{name}[{label}]")


def test_shim_allocates_outputs_without_explicit_dims():
    """sympl gives a tendency without `dims` the dims of the input quantity of the same name (RRTMGShortwave relies on it)"""
    import numpy as np
    from climt_b200 import sympl_shim as SS
    if SS.HAVE_SYMPL:
        pytest.skip("real sympl installed")
    from climt_b200.rrtmg_sw import RRTMGShortwave
    raw = {"air_temperature": np.zeros((7, 5)), "air_pressure_on_interface_levels": np.zeros((8, 5))}
    out = SS.initialize_numpy_arrays_with_properties(RRTMGShortwave.tendency_properties, raw, RRTMGShortwave.input_properties)
    assert out["air_temperature"].shape == (7, 5)
    out = SS.initialize_numpy_arrays_with_properties(RRTMGShortwave.diagnostic_properties, raw, RRTMGShortwave.input_properties)
    assert out["upwelling_shortwave_flux_in_air"].shape == (8, 5) and out["air_temperature_tendency_from_shortwave"].shape == (7, 5)
