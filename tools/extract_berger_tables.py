#!/usr/bin/env python
"""Extract the Berger (1978) series coefficients the reference's BergerSolarInsolation holds as module-level arrays
(climt/_components/berger_solar_insolation.py:7-492: obliquity A, f, delta; eccentricity P, alpha, zeta; general precession
F, f_prime, delta_prime; arcsec_to_degree) into climt_b200/data/berger1978.npz.  Data, not code: run once in the build container.

    python tools/extract_berger_tables.py
"""
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    def stub(name, path=None, **attrs):
        m = types.ModuleType(name)
        if path:
            m.__path__ = [REF + path]
        m.__dict__.update(attrs)
        sys.modules[name] = m

    class Component:
        def __init__(self, **kwargs):
            pass
    stub("sympl", DiagnosticComponent=Component, get_constant=lambda *a: 1367.0)
    for name, path in (("climt", "/climt"), ("climt._core", "/climt/_core"), ("climt._components", "/climt/_components")):
        stub(name, path)
    mod = importlib.import_module("climt._components.berger_solar_insolation")
    names = ("A", "f", "delta", "P", "alpha", "zeta", "F", "f_prime", "delta_prime")
    out = {n: np.asarray(getattr(mod, n), dtype=np.float64) for n in names}
    out["arcsec_to_degree"] = np.float64(mod.arcsec_to_degree)
    dst = os.path.join(ROOT, "climt_b200", "data", "berger1978.npz")
    np.savez(dst, **out)
    print(dst, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
