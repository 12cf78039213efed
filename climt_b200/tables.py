"""Table plumbing for the RRTMG engines: raw (16-g) data -> packed blobs.

`rrtmg_{lw,sw}_raw.npz` are produced once by tools/extract_rrtmg_tables.py from the
reference's data statements.  A *blob* is the flat little-endian container the native
code reads (oracle/ftn.hpp and climt_b200/csrc/tables.h share the format):

    char magic[8] = "CB2TBL01"; int64 n;
    n x { char name[56]; int64 ndim; int64 shape[6]; int64 offset; int64 count }
    double data[]

`order="F"` stores every array in Fortran element order (what the oracle wants, so its
1-based column-major views index exactly like the Fortran); `order="C"` stores the numpy
layout as is (what the CUDA engine wants for its re-laid-out, g-fastest tables).
"""
import os
import struct

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_raw(tag):
    """tag in {"lw", "sw"} -> dict name -> float64 ndarray (Fortran index order, see extractor)."""
    with np.load(os.path.join(DATA_DIR, f"rrtmg_{tag}_raw.npz")) as z:
        return {k: np.asarray(z[k], dtype=np.float64) for k in z.files}


def write_blob(arrays, path, order="C"):
    names = sorted(arrays)
    entries, chunks, off = [], [], 0
    for k in names:
        a = np.asarray(arrays[k], dtype=np.float64)
        if a.ndim > 6:
            raise ValueError(f"{k}: more than 6 dims")
        flat = a.ravel(order=order)
        shape = list(a.shape) + [0] * (6 - a.ndim)
        nm = k.encode()
        if len(nm) > 55:
            raise ValueError(f"name too long: {k}")
        entries.append(struct.pack("<56sq6qqq", nm, a.ndim, *shape, off, flat.size))
        chunks.append(flat.tobytes())
        off += flat.size
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(b"CB2TBL01")
        f.write(struct.pack("<q", len(names)))
        for e in entries:
            f.write(e)
        for c in chunks:
            f.write(c)
    os.replace(tmp, path)
    return path


def raw_blob_path(tag, cache_dir=None):
    """Write (once) and return the Fortran-order raw blob used by the oracle."""
    cache_dir = cache_dir or os.path.join(DATA_DIR, "_cache")
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"rrtmg_{tag}_raw_F.blob")
    src = os.path.join(DATA_DIR, f"rrtmg_{tag}_raw.npz")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        raw = {k: v for k, v in load_raw(tag).items() if not k.endswith("__lb")}
        write_blob(raw, path, order="F")
    return path
