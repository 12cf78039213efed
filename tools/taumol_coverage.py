#!/usr/bin/env python
"""Which branches of the taumol evaluators does which kind of check reach?

The evaluators (`eval_band` in climt_b200/csrc/{lw,sw}_core.cuh -- the code the CUDA kernels run, compiled for the host by
tests/emul with -DCB_COVERAGE) count, per band and per lower / upper atmosphere, the data-dependent paths listed in cb::CovId
(cb_common.h).  This tool feeds them three sets of states and writes profiles/taumol_branch_coverage.{md,json}:

  R  the states behind the reference's golden caches (tests/cached_component_output/TestRRTMG*): the only vectors the reference
     itself holds for RRTMG.  All of them are DRY columns (specific_humidity defaults to 0, _core/initialization.py:755).
  X  the humid columns of tests/test_humid_crosscheck.py, where RRTMG is checked against CORK (bit-exact to the reference's numba
     kernels) -- the reference's own cross-check (tests/test_rrtmg_comparison.py) -- to < 1 % (LW up) / ~3 % (SW down).
  S  the synthetic states of the parity tests and of bench.py (humid, trace gases, high CO2 / N2O): CUDA == oracle to 1e-12, both
     written from the Fortran by the same reader.

Run: python tools/taumol_coverage.py         (CPU only, a few seconds)
"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import helpers as H  # noqa: E402
from climt_b200 import constants as C, rrtmg_tables as RT, state as S, synthetic as SY  # noqa: E402

_dp = ctypes.POINTER(ctypes.c_double)
IDS = ["region", "key>0", "s0<.125", "s0 mid", "s0>.875", "s1<.125", "s1 mid", "s1>.875", "self>0", "foreign>0", "minor0>0",
       "minor1>0", "minor2>0", "minor0 adj", "minor1 adj", "minor2 adj", "xsec0>0", "xsec1>0", "planck/source interp"]


def cov_lib(which):
    so = os.path.join(ROOT, "tests", "emul", f"libcb_cov_{which}.so")
    src = os.path.join(ROOT, "tests", "emul", f"{which}_emul.cpp")
    deps = [src] + [os.path.join(ROOT, "climt_b200", "csrc", f) for f in ("lw_core.cuh", "sw_core.cuh", "cb_common.h", "lw_tables.h", "sw_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-DCB_COVERAGE", "-o", so, src])
    return ctypes.CDLL(so)


def consts():
    k = C.rrtmg_constants()
    return np.array([k[n] for n in ("pi", "grav", "planck", "boltz", "clight", "avogad", "alosmt", "gascon", "sbcnst", "secdy", "cpdair")])


def run_lw(lib, st, flags=(1, 0, 2, 1, 1), mcica=(0, 1, 0)):
    nlay, ncol = st["play"].shape
    inp = (_dp * 23)(*[np.ascontiguousarray(st[f]).ctypes.data_as(_dp) for f in SY.LW_FIELDS])
    outs = [np.zeros((nlay + 1, ncol)) for _ in range(8)]
    outp = (_dp * 8)(*[o.ctypes.data_as(_dp) for o in outs])
    rc = lib.emul_lw_run(RT.lw_blob_path().encode(), consts().ctypes.data_as(_dp), (ctypes.c_int * 8)(*(tuple(flags) + tuple(mcica))),
                         ncol, nlay, inp, outp)
    assert rc == 0, rc


def run_sw(lib, st, iopt=(1, 0, 2, 1, 1, 0, 1), mcica=(0, 1, 0)):
    nlay, ncol = st["play"].shape
    scal = np.array([1.0, 1367.0, 0.0, 1.0, 1.0] + [1.0] * 14)
    inp = (_dp * 29)(*[np.ascontiguousarray(st[f]).ctypes.data_as(_dp) for f in SY.SW_FIELDS])
    outs = [np.zeros((nlay + 1, ncol)) for _ in range(6)]
    outp = (_dp * 6)(*[o.ctypes.data_as(_dp) for o in outs])
    rc = lib.emul_sw_run(RT.sw_blob_path().encode(), consts().ctypes.data_as(_dp), (ctypes.c_int * 10)(*(tuple(iopt) + tuple(mcica))),
                         scal.ctypes.data_as(_dp), ncol, nlay, inp, outp)
    assert rc == 0, rc


def counters(lib):
    n = lib.emul_cov_n()
    a = np.zeros((32, 2, n), dtype=np.int64)
    lib.emul_cov_get(a.ctypes.data_as(ctypes.POINTER(ctypes.c_long)))
    return a


def humid_states():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_humid_crosscheck as T
    lw, sw, _, _ = T._states()
    return lw, sw


def collect():
    out = {}
    for which, run, bands in (("lw", run_lw, range(1, 17)), ("sw", run_sw, range(16, 30))):
        lib = cov_lib(which)
        sets = {}
        # R: the golden families (tests/test_components.py:435-531): column (30 levels), 3d (28), clouds, external Tint, McICA
        lib.emul_cov_reset()
        if which == "lw":
            run(lib, H.default_lw_abi_state(30, 1))
            run(lib, H.default_lw_abi_state(28, 2))
            run(lib, H.default_lw_abi_state(30, 1, external_tint=True))
            st = H.default_lw_abi_state(28, 4)
            st["cldfr"][16:19] = 0.5
            st["cicewp"][16:19] = 0.3e3
            run(lib, st)
            run(lib, st, mcica=(1, 1, 7))
        else:
            run(lib, H.default_sw_abi_state(30, 1))
            run(lib, H.default_sw_abi_state(28, 2))
            st = H.default_sw_abi_state(15, 6)
            st["cldfr"][10:12] = 0.5
            st["cicewp"][10:12] = 0.3e3
            run(lib, st, mcica=(1, 1, 7))
        sets["R"] = counters(lib)
        lib.emul_cov_reset()
        lw, sw = humid_states()
        run(lib, lw if which == "lw" else sw, *(((0, 0, 2, 1, 1),) if which == "lw" else ((0, 0, 2, 1, 1, 0, 0),)))
        sets["X"] = counters(lib)
        lib.emul_cov_reset()
        if which == "lw":
            run(lib, SY.make_lw_state(48, 60, seed=20260925, clouds=True, aerosol=True))
            run(lib, SY.make_lw_state(24, 72, seed=3, clouds=False))
            hi = SY.make_lw_state(8, 40, seed=4)
            hi["co2"][:] = 4e-3   # tests/test_lw_emul.py::test_high_co2_and_n2o_trigger_adjusted_columns
            hi["n2o"][:] = 4e-6
            run(lib, hi)
        else:
            run(lib, SY.make_sw_state(48, 60, seed=20260925, clouds=True))
            run(lib, SY.make_sw_state(24, 72, seed=3))
        sets["S"] = counters(lib)
        res = {}
        for b in bands:
            for lower in (1, 0):
                row = {}
                for i, name in enumerate(IDS):
                    tags = "".join(t for t in "RXS" if sets[t][b, lower, i] > 0)
                    if tags:
                        row[name] = tags
                res[f"band {b} {'lower' if lower else 'upper'}"] = row
        out[which] = res
    return out


def markdown(cov):
    lines = ["# taumol branch coverage: which check reaches which path of the evaluators", "",
             "Generated by `python tools/taumol_coverage.py` (host build of `eval_band`, the code the CUDA kernels run, with",
             "`-DCB_COVERAGE`).  Letters: **R** = states behind the reference's golden caches (all dry), **X** = humid columns checked",
             "against CORK as the reference's own `tests/test_rrtmg_comparison.py` does (CORK pinned bit-exactly to the reference),",
             "**S** = synthetic parity / benchmark states (CUDA vs the C++ restatement only).  A path not listed for a band does not",
             "exist there.  `>0` = the term contributes a non-zero optical depth (with q = 0 the water-vapour terms are evaluated but",
             "multiply by zero).", ""]
    for which, title in (("lw", "Longwave (rrtmg_lw_taumol.f90, bands 1-16)"), ("sw", "Shortwave (rrtmg_sw_taumol.f90, bands 16-29)")):
        lines += [f"## {title}", "", "| band / region | " + " | ".join(IDS) + " |", "|---|" + "---|" * len(IDS)]
        for key, row in cov[which].items():
            if not row:
                continue
            lines.append(f"| {key} | " + " | ".join(row.get(n, "") for n in IDS) + " |")
        lines.append("")
    # summary: what only S reaches
    only_s = [(w, k, n) for w in cov for k, row in cov[w].items() for n, t in row.items() if t == "S"]
    not_r = [(w, k, n) for w in cov for k, row in cov[w].items() for n, t in row.items() if "R" not in t]
    lines += ["## Summary", "",
              f"* paths reached by some check: {sum(len(r) for w in cov for r in cov[w].values())}; "
              f"not reached by a reference golden (R): {len(not_r)}; of those pinned by the CORK cross-check (X): "
              f"{len(not_r) - len(only_s)}; reached by synthetic states only (S): {len(only_s)}.",
              "* S-only paths (agreement of two independently structured implementations is all that pins them): "
              + "; ".join(f"{w.upper()} {k}: {n}" for w, k, n in only_s) + ".", ""]
    return "\n".join(lines)


def main():
    cov = collect()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(cov, open(os.path.join(ROOT, "profiles", "taumol_branch_coverage.json"), "w"), indent=1)
    md = markdown(cov)
    open(os.path.join(ROOT, "profiles", "taumol_branch_coverage.md"), "w").write(md)
    print(md[-1500:])


if __name__ == "__main__":
    main()
