#!/usr/bin/env python
"""The component surface the drop-ins must keep: input / tendency / diagnostic / output properties (names, dims,
units, aliases) and constructor signatures of the reference's component classes, read off the reference's own classes.

The reference modules import once `sympl` and `climt._core` are replaced by permissive stubs (their Cython / Fortran extensions
are guarded by try/except in the reference itself); class-level property dicts are read as they are, the CORK components'
instance-level properties by constructing them for each optics mode / table class.

Run in the build container (the GPU box has no /root/reference):   python tests/golden/make_properties_golden.py
Writes tests/golden/reference_properties.json.
"""
import importlib
import importlib.resources
import inspect
import json
import os
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


class _Permissive(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def dummy(*a, **k):
            return a[0] if (len(a) == 1 and callable(a[0]) and not k) else None   # works as a decorator too
        return dummy


def reference_classes():
    from climt_b200.constants import get_constant

    class Component:
        def __init__(self, **kwargs):
            pass

    sympl = _Permissive("sympl")
    for n in ("TendencyComponent", "DiagnosticComponent", "Stepper", "ImplicitTendencyComponent"):
        setattr(sympl, n, type(n, (Component,), {}))
    sympl.get_constant = get_constant
    sys.modules["sympl"] = sympl
    for name, path in (("climt", "/climt"), ("climt._components", "/climt/_components"), ("climt._components.cork", "/climt/_components/cork"),
                       ("climt._components.rrtmg", "/climt/_components/rrtmg"), ("climt._components.rrtmg.lw", "/climt/_components/rrtmg/lw"),
                       ("climt._components.rrtmg.sw", "/climt/_components/rrtmg/sw"), ("climt._components.emanuel", "/climt/_components/emanuel"),
                       ("climt._components.simple_physics", "/climt/_components/simple_physics"),
                       ("climt._components.instellation", "/climt/_components/instellation"), ("climt._data", "/climt/_data"),
                       ("climt._data.cork", "/climt/_data/cork")):
        m = types.ModuleType(name)
        m.__path__ = [REF + path]
        sys.modules[name] = m
    core = _Permissive("climt._core")
    core.__path__ = [REF + "/climt/_core"]
    sys.modules["climt._core"] = core
    sys.modules["climt._core.initialization"] = _Permissive("climt._core.initialization")
    sys.modules["climt._core.horizontal_operators"] = _Permissive("climt._core.horizontal_operators")
    sys.modules["importlib_resources"] = importlib.resources
    imp = importlib.import_module
    out = {
        "RRTMGLongwave": imp("climt._components.rrtmg.lw.component").RRTMGLongwave,
        "RRTMGShortwave": imp("climt._components.rrtmg.sw.component").RRTMGShortwave,
        "GrayLongwaveRadiation": imp("climt._components.radiation").GrayLongwaveRadiation,
        "EmanuelConvection": imp("climt._components.emanuel.component").EmanuelConvection,
        "EmanuelConvectionPython": imp("climt._components.emanuel.pure_python_v3").EmanuelConvectionPython,
        "SimplePhysics": imp("climt._components.simple_physics.component").SimplePhysics,
        "Instellation": imp("climt._components.instellation.component").Instellation,
        "BergerSolarInsolation": imp("climt._components.berger_solar_insolation").BergerSolarInsolation,
        "SlabSurface": imp("climt._components.slab_surface").SlabSurface,
        "CorkLongwaveRadiation": imp("climt._components.cork.lw.component").CorkLongwaveRadiation,
        "CorkShortwaveRadiation": imp("climt._components.cork.sw.component").CorkShortwaveRadiation,
    }
    return out


def props(obj):
    res = {}
    for kind in ("input_properties", "tendency_properties", "diagnostic_properties", "output_properties"):
        p = getattr(obj, kind, None)
        if isinstance(p, dict):
            res[kind] = {name: {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in d.items() if k in ("dims", "units", "alias")}
                         for name, d in p.items()}
    return res


def signature(cls):
    sig = inspect.signature(cls.__init__)
    return {n: (None if p.default is inspect.Parameter.empty else repr(p.default)) for n, p in sig.parameters.items()
            if n != "self" and p.kind is not inspect.Parameter.VAR_KEYWORD}


def main():
    C = reference_classes()
    out = {}
    for name, cls in C.items():
        out[name] = {"init": signature(cls)}
        if not name.startswith("Cork"):
            out[name].update(props(cls))
    tdir = REF + "/climt/_data/cork/correlated_k/"
    for name in ("CorkLongwaveRadiation", "CorkShortwaveRadiation"):
        which = "lw" if "Long" in name else "sw"
        out[name]["instances"] = {}
        for label, kw in (("parmentier", {"optics": "parmentier"}),
                          ("earth", {"optics": "correlated_k", "table": tdir + f"earth_low_res_{which}.npz"}),
                          ("mars", {"optics": "correlated_k", "table": tdir + f"mars_{which}.nc"}),
                          ("titan", {"optics": "correlated_k", "table": tdir + f"titan_{which}.nc"})):
            out[name]["instances"][label] = props(C[name](**kw))
    out["CorkLongwaveRadiation"]["instances"]["tour_gray"] = props(C["CorkLongwaveRadiation"](optics="correlated_k", table=tdir + "tour_gray_lw.nc"))
    out["SlabSurface"]["instances"] = {"ekman": props(C["SlabSurface"](include_ekman=True))}
    dst = os.path.join(HERE, "reference_properties.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print(dst, os.path.getsize(dst), "bytes;", {k: len(v.get("input_properties", {})) for k, v in out.items()})


if __name__ == "__main__":
    main()
