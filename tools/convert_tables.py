#!/usr/bin/env python
"""k-table converter and manifest tool (format: climt_b200/table_store.py).

  python tools/convert_tables.py convert <table.nc|table.npz|shipped name> [-o out.cb2k]   # reference format -> engine container
  python tools/convert_tables.py info <file.cb2k>                                          # entries of a container
  python tools/convert_tables.py manifest --write | --verify                               # climt_b200/data/MANIFEST.json
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from climt_b200 import table_store as TS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("convert")
    c.add_argument("src")
    c.add_argument("-o", "--out")
    i = sub.add_parser("info")
    i.add_argument("path")
    m = sub.add_parser("manifest")
    g = m.add_mutually_exclusive_group(required=True)
    g.add_argument("--write", action="store_true")
    g.add_argument("--verify", action="store_true")
    a = ap.parse_args()
    if a.cmd == "convert":
        dst = TS.convert_k_table(a.src, a.out)
        print(dst, os.path.getsize(dst), "bytes, sha256", TS.file_sha256(dst))
    elif a.cmd == "info":
        for k, v in TS.read_container(a.path).items():
            print(f"{k:28s}", repr(v) if isinstance(v, str) else f"{v.dtype.name} {tuple(v.shape)}")
    elif a.write:
        print(TS.write_manifest())
    else:
        bad = TS.verify_manifest()
        print("\n".join(bad) if bad else "manifest verified")
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
