#!/usr/bin/env python
"""Extract the numerical DATA tables of AER RRTMG (as shipped in climt) into .npz.

Run once in the build container (needs /root/reference; the GPU box does not
have it).  Only numbers are extracted -- k-distribution coefficients, Planck
integrals, reference profiles, cloud-optics and aerosol tables -- never code.

What is parsed (all `name(slice) = (/ v, v, ... /)` array-constructor
statements, the regular form documented in SURVEY.md section 7.1 step 0):

  LW  climt/_lib/rrtmg_lw/rrtmg_lw_k_g.f90      (16 x lw_kgbNN, 16-g originals)
      climt/_lib/rrtmg_lw/rrtmg_lw_setcoef.f90  :418-1990 (pref, preflog, tref,
                                                 chi_mls, totplnk, totplk16, derivs)
      climt/_lib/rrtmg_lw/rrtmg_lw_init.f90     :2018-2656 (lwcldpr)
  SW  climt/_lib/rrtmg_sw/rrtmg_sw_k_g.f90      (14 x sw_kgbNN)
      climt/_lib/rrtmg_sw/rrtmg_sw_setcoef.f90  :308-362 (swatmref)
      climt/_lib/rrtmg_sw/rrtmg_sw_init.f90     :389-489 (swaerpr), :1692-3514 (swcldpr)

Array shapes and lower bounds come from the module declaration files
(rrlw_kgNN.f90, rrlw_ref.f90, rrlw_wvn.f90, rrlw_cld.f90 and the rrsw_*
twins).  Arrays are stored with the Fortran index order preserved
(`a[i-lb0, j-lb1, ...]` == Fortran `a(i, j, ...)`), float64, C-contiguous numpy;
lower bounds different from 1 are stored in `<key>__lb`.

Keys are `<module>.<name>`, e.g. `rrlw_kg03.kbo` (shape (5,5,47,16), lb (1,1,13,1)).
Scalars assigned with a plain `name = <number>_rb` inside the k_g subroutines
(SW `rayl`) are stored as 0-d arrays.
"""
import argparse
import glob
import os
import re
import sys

import numpy as np

REF = "/root/reference/climt/_lib"


def strip_comment(line):
    # no string literals with '!' in the data sections we parse
    i = line.find("!")
    return line if i < 0 else line[:i]


def logical_lines(path):
    """Join Fortran free-form continuation lines (trailing &, optional leading &)."""
    out, cur, start = [], "", 0
    with open(path, "r", errors="replace") as f:
        for ln, raw in enumerate(f, 1):
            s = strip_comment(raw.rstrip("\n")).strip()
            if not s:
                continue
            if not cur:
                start = ln
            if s.startswith("&"):
                s = s[1:].lstrip()
            if s.endswith("&"):
                cur += s[:-1] + " "
                continue
            cur += s
            out.append((start, cur))
            cur = ""
    return out


def parse_modules(paths):
    """-> params {name: int}, decls {module: {name: (shape, lbounds)}}"""
    params, raw_decl = {}, {}
    for p in paths:
        mod = None
        for _, s in logical_lines(p):
            low = s.lower()
            m = re.match(r"module\s+(\w+)", low)
            if m and not low.startswith("module procedure"):
                mod = m.group(1)
                raw_decl.setdefault(mod, [])
                continue
            if "::" not in low:
                continue
            head, tail = low.split("::", 1)
            if "parameter" in head:
                for pm in re.finditer(r"(\w+)\s*=\s*([-+0-9.e_a-z*/ ()]+?)(?:,|$)", tail):
                    name, val = pm.group(1), pm.group(2).strip()
                    val = re.sub(r"_rb|_im", "", val)
                    try:
                        params[name] = eval(val, {}, dict(params))
                    except Exception:
                        pass
                continue
            if not (head.strip().startswith("real") or head.strip().startswith("integer")):
                continue
            dm = re.search(r"dimension\s*\(([^)]*)\)", head)
            if dm:
                for nm in tail.split(","):
                    nm = nm.strip()
                    if nm:
                        raw_decl[mod].append((nm, dm.group(1)))
            else:
                for vm in re.finditer(r"(\w+)\s*(?:\(([^)]*)\))?\s*(?:,|$)", tail):
                    if vm.group(1):
                        raw_decl[mod].append((vm.group(1), vm.group(2)))
    decls = {}
    for mod, lst in raw_decl.items():
        d = {}
        for name, dims in lst:
            if dims is None:
                d[name] = ((), ())
                continue
            shape, lbs = [], []
            for part in dims.split(","):
                part = part.strip()
                if ":" in part:
                    lo, hi = part.split(":")
                    lo, hi = eval(lo, {}, dict(params)), eval(hi, {}, dict(params))
                else:
                    lo, hi = 1, eval(part, {}, dict(params))
                shape.append(hi - lo + 1)
                lbs.append(lo)
            d[name] = (tuple(shape), tuple(lbs))
        decls[mod] = d
    return params, decls


NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][-+]?\d+)?(?:_rb|_im)?")


def parse_values(txt):
    vals = []
    for tok in NUM.findall(txt):
        tok = tok.replace("_rb", "").replace("_im", "").replace("d", "e").replace("D", "e")
        vals.append(float(tok))
    return np.array(vals, dtype=np.float64)


def extract(source_paths, decls, out):
    assign = re.compile(r"^(\w+)\s*\(([^)]*)\)\s*=\s*\(/(.*)/\)\s*$")
    scalar = re.compile(r"^(\w+)\s*=\s*([-+0-9.eE_rb /()]+)$")
    for p in source_paths:
        base = {}      # module-level `use` (before `contains`)
        uses = base    # name -> module visible in the current scope
        for ln, s in logical_lines(p):
            low = s.lower()
            if re.match(r"(end\s+)?subroutine\b", low):
                if not low.startswith("end"):
                    uses = dict(base)
                continue
            m = re.match(r"use\s+(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", low)
            if m:
                mod = m.group(1)
                if mod in decls:
                    if m.group(2):
                        for item in m.group(2).split(","):
                            item = item.strip()
                            if "=>" in item:
                                item = item.split("=>")[1].strip()
                            if item:
                                uses[item] = mod
                    else:
                        for nm in decls[mod]:
                            uses[nm] = mod
                continue
            m = assign.match(low)
            if m:
                name, idx, body = m.group(1), m.group(2), m.group(3)
                if name not in uses:
                    continue
                mod = uses[name]
                shape, lbs = decls[mod][name]
                key = f"{mod}.{name}"
                if key not in out:
                    out[key] = np.full(shape, np.nan)
                    if any(lb != 1 for lb in lbs):
                        out[key + "__lb"] = np.array(lbs, dtype=np.int64)
                arr = out[key]
                sl = []
                for part, lb, n in zip(idx.split(","), lbs, shape):
                    part = part.strip()
                    if part == ":":
                        sl.append(slice(0, n))
                    elif ":" in part:
                        lo, hi = part.split(":")
                        sl.append(slice(int(lo) - lb, int(hi) - lb + 1))
                    else:
                        sl.append(int(part) - lb)
                vals = parse_values(body)
                target = arr[tuple(sl)]
                if target.size != vals.size:
                    raise ValueError(f"{p}:{ln}: {name}({idx}) expects {target.size} values, got {vals.size}")
                # Fortran array constructor fills in column-major order of the section
                arr[tuple(sl)] = vals.reshape(target.shape, order="F")
                continue
            m = scalar.match(low)
            if m and m.group(1) in uses:
                name = m.group(1)
                mod = uses[name]
                if decls[mod][name][0] == ():
                    expr = m.group(2).replace("_rb", "")
                    try:
                        out[f"{mod}.{name}"] = np.array(float(eval(expr)))
                    except Exception:
                        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=REF)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "climt_b200", "data"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for tag, sub, mods, srcs in (
        ("lw", "rrtmg_lw", ["parrrtm.f90", "rrlw_kg*.f90", "rrlw_ref.f90", "rrlw_wvn.f90", "rrlw_cld.f90"],
         ["rrtmg_lw_k_g.f90", "rrtmg_lw_setcoef.f90", "rrtmg_lw_init.f90"]),
        ("sw", "rrtmg_sw", ["parrrsw.f90", "rrsw_kg*.f90", "rrsw_ref.f90", "rrsw_wvn.f90", "rrsw_cld.f90", "rrsw_aer.f90"],
         ["rrtmg_sw_k_g.f90", "rrtmg_sw_setcoef.f90", "rrtmg_sw_init.f90"]),
    ):
        d = os.path.join(a.ref, sub)
        mpaths = []
        for pat in mods:
            mpaths += sorted(glob.glob(os.path.join(d, pat)))
        params, decls = parse_modules(mpaths)
        out = {}
        extract([os.path.join(d, s) for s in srcs], decls, out)
        if tag == "sw":
            # rrtmg_sw_k_g.f90:62442,62460-62461 -- the only non-literal data statement: band 29's irradnceo is
            # scaled in place by `irradscl`, declared default `real` (float32), so the factor is rounded to
            # float32 before the multiplication.
            irradscl = np.float64(np.float32(13.221 / (13.221 - 0.455)))
            out["rrsw_kg29.irradnceo"] = irradscl * out["rrsw_kg29.irradnceo"]
        bad = [k for k, v in out.items() if v.dtype == np.float64 and np.isnan(v).any()]
        # arrays that are declared but only partly data-initialised are dropped
        # (e.g. reduced-g arrays filled by cmbgb at run time never appear here)
        for k in bad:
            print(f"[{tag}] WARNING: {k} has unassigned entries ({np.isnan(out[k]).sum()} of {out[k].size})",
                  file=sys.stderr)
        path = os.path.join(a.out, f"rrtmg_{tag}_raw.npz")
        np.savez_compressed(path, **out)
        tot = sum(v.size for v in out.values())
        print(f"[{tag}] {len(out)} arrays, {tot} values -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


if __name__ == "__main__":
    main()
