#!/usr/bin/env python
"""Convert the reference's golden NetCDF-3 caches for the radiation hot path into one .npz.

Source: /root/reference/tests/cached_component_output/Test{RRTMG*,GrayLongwaveRadiation,SimplePhysics}-{column,3d}-{0,1}.cache
(written by the reference's own test harness, tests/test_components.py:63-75,186-193;
`-0` = tendencies, `-1` = diagnostics).  The `*_stepping` caches need sympl's
AdamsBashforth stepper and are not used.  Run here (the GPU box has no /root/reference):

    python tests/golden/make_golden.py
Keys: "<TestClass>-<column|3d>/<quantity>" -> float64 array with the file's (lev, lat, lon) shape.
"""
import glob
import os
import sys

import numpy as np
from scipy.io import netcdf_file

SRC = "/root/reference/tests/cached_component_output"
CLASSES = ["TestRRTMGLongwave", "TestRRTMGLongwaveMCICA", "TestRRTMGLongwaveWithClouds",
           "TestRRTMGLongwaveWithExternalInterfaceTemperature", "TestRRTMGShortwave",
           "TestRRTMGShortwaveMCICA", "TestGrayLongwaveRadiation",
           # a Stepper: part 0 = its diagnostics (stored under "tend"), part 1 = the new state (stored under "diag")
           "TestSimplePhysics"]


def main():
    out = {}
    for cls in CLASSES:
        for kind in ("column", "3d"):
            for part in (0, 1):
                path = os.path.join(SRC, f"{cls}-{kind}-{part}.cache")
                if not os.path.exists(path):
                    print("missing", path, file=sys.stderr)
                    continue
                nc = netcdf_file(path, "r", mmap=False)
                for name, var in nc.variables.items():
                    key = f"{cls}-{kind}/{'tend' if part == 0 else 'diag'}/{name}"
                    out[key] = np.array(var[:], dtype=np.float64)
                nc.close()
    # The `*_stepping` caches hold the diagnostics AND the whole state after one Adams-Bashforth step of 10 s on the component's
    # default state: every input the reference's get_default_state produced (pressures, ozone, gases, cloud defaults) plus the
    # stepped temperature.  Column versions only; stored as "<TestClass>-column_stepping/{diag,state}/<quantity>".
    for cls in ("TestRRTMGLongwave", "TestRRTMGShortwave"):
        for part, tag in ((0, "diag"), (1, "state")):
            path = os.path.join(SRC, f"{cls}-column_stepping-{part}.cache")
            nc = netcdf_file(path, "r", mmap=False)
            for name, var in nc.variables.items():
                out[f"{cls}-column_stepping/{tag}/{name}"] = np.array(var.data if var.shape == () else var[:], dtype=np.float64)
            nc.close()
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_caches.npz")
    np.savez_compressed(dst, **out)
    print(f"{len(out)} arrays -> {dst} ({os.path.getsize(dst)/1e3:.1f} kB)")


if __name__ == "__main__":
    main()
