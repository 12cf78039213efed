"""Lists every C symbol the reference's two RRTMG Cython shims declare `extern` (the names a Cython extension linked against
libclimt_b200.so must resolve at import, INTEGRATION.md option A) -> tests/golden/pyx_externs.json.

Run in the build container (reads /root/reference): python tests/golden/make_pyx_externs.py
"""
import json
import os
import re

REF = os.environ.get("CLIMT_REFERENCE", "/root/reference")
FILES = {
    "lw": "climt/_components/rrtmg/lw/_rrtmg_lw.pyx",
    "sw": "climt/_components/rrtmg/sw/_rrtmg_sw.pyx",
}


def externs(path):
    """names declared inside `cdef extern:` blocks: lines of the form `void name(` / `void name (`"""
    names, inside = [], False
    for line in open(path):
        if re.match(r"\s*cdef\s+extern\b", line):
            inside = True
            continue
        if inside:
            if line.strip() and not line.startswith((" ", "\t")):
                inside = False
                continue
            m = re.match(r"\s+(?:void|int|double)\s+(\w+)\s*\(", line)
            if m:
                names.append(m.group(1))
    return names


def main():
    out = {k: {"file": v, "symbols": externs(os.path.join(REF, v))} for k, v in FILES.items()}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyx_externs.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
