"""On-disk side of the k-tables: one typed container the native engine reads without Python, a converter from
the reference's formats, and a sha-pinned manifest of every table this package ships or derives.

Reference formats (read by `cork.load_k_table`, same contract as cork/optics/correlated_k.py:120-218): NetCDF-3 `.nc` as written
by scripts/cork_table_builder/netcdf_writer.py, and the `.npz` fixtures (scripts/convert_ck_table_to_npz.py); provenance and
sha256 of the reference's own files: climt/_data/cork/correlated_k/MANIFEST.md.

Container (`.cb2k`, little endian; every payload 64-byte aligned so a table can be memory-mapped or read straight into pinned
memory and uploaded):

    char magic[8] = "CB2KTB01"; int64 n;
    n x { char name[48]; int32 dtype; int32 ndim; int64 shape[7]; int64 offset; int64 nbytes }      (128 bytes per entry)
    payloads
    dtype: 0 = float64, 1 = float32, 2 = int32, 3 = UTF-8 text (shape = [nbytes])

A k-table keeps the reference's array names, shapes and dtypes (a float32 table stays float32: it is half the HBM traffic of the
dominant CORK kernel and the reference promotes on use as well), plus the classification the reference constructors derive
(cork/lw/component.py:46-59) resolved once at conversion time: int32 scalars `_premixed`, `_co2_logk`, `_overlap_additive`.
The g-point-contiguous re-layout for the GPU (`csrc/cork_tables.h`) depends on the engine's unit width and is done at create;
`cb200_cork_create_from_file` (include/climt_b200.h) is the whole load path for a host without numpy/scipy.
"""
import hashlib
import json
import os
import struct

import numpy as np

MAGIC = b"CB2KTB01"
_DT = {0: np.dtype("<f8"), 1: np.dtype("<f4"), 2: np.dtype("<i4")}
_CODE = {np.dtype("float64"): 0, np.dtype("float32"): 1, np.dtype("int32"): 2}
_ENTRY = "<48sii7qqq"
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
MANIFEST = os.path.join(DATA_DIR, "MANIFEST.json")
K_TABLE_ARRAYS = ("k_coefficients", "gpoint_weights", "temperature_grid", "pressure_grid_log", "h2o_vmr_grid", "co2_vmr_grid",
                  "band_wavenumber_limits", "planck_fraction", "solar_source_per_gpoint", "rayleigh_coefficient", "continuum_kappa")
K_TABLE_TEXT = ("overlap_method", "resolution", "background_is_premixed")


def write_container(path, arrays):
    """arrays: name -> ndarray (float64 / float32 / int32; other integer and float types are converted) or str."""
    entries, payloads = [], []
    off = 16 + 128 * len(arrays)
    for name in sorted(arrays):
        v = arrays[name]
        nm = name.encode()
        if len(nm) > 47:
            raise ValueError(f"name too long: {name}")
        if isinstance(v, str):
            raw, code, shape = v.encode("utf-8"), 3, [len(v.encode("utf-8"))]
        else:
            a = np.asarray(v)
            if a.dtype.kind in "US":
                raise TypeError(f"{name}: pass text as str")
            dt = a.dtype if a.dtype in _CODE else (np.dtype("int32") if a.dtype.kind in "iub" else np.dtype("float64"))
            a = np.ascontiguousarray(a, dtype=dt)
            if a.ndim > 7:
                raise ValueError(f"{name}: more than 7 dims")
            raw, code, shape = a.astype(_DT[_CODE[dt]], copy=False).tobytes(), _CODE[dt], list(a.shape)
        off = (off + 63) // 64 * 64
        entries.append(struct.pack(_ENTRY, nm, code, len(shape), *(shape + [0] * (7 - len(shape))), off, len(raw)))
        payloads.append((off, raw))
        off += len(raw)
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(MAGIC + struct.pack("<q", len(entries)))
        for e in entries:
            f.write(e)
        for o, raw in payloads:
            f.write(b"\0" * (o - f.tell()))
            f.write(raw)
    os.replace(tmp, path)
    return path


def read_container(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != MAGIC:
        raise ValueError(f"{path}: not a {MAGIC.decode()} container")
    (n,) = struct.unpack_from("<q", raw, 8)
    out = {}
    for i in range(n):
        nm, code, ndim, *rest = struct.unpack_from(_ENTRY, raw, 16 + 128 * i)
        shape, off, nbytes = rest[:7][:ndim], rest[7], rest[8]
        name = nm.rstrip(b"\0").decode()
        if off % 64 or off + nbytes > len(raw):
            raise ValueError(f"{path}: bad payload extent of {name}")
        if code == 3:
            out[name] = raw[off:off + nbytes].decode("utf-8")
        else:
            out[name] = np.frombuffer(raw, dtype=_DT[code], count=nbytes // _DT[code].itemsize, offset=off).reshape(shape).copy()
    return out


def ktable_to_container_arrays(table):
    """dict as returned by `cork.load_k_table` -> entries of a .cb2k file (reference names; classification resolved)"""
    from .cork import CO2_INTERP_LOGK, table_flags
    out = {}
    for name in K_TABLE_ARRAYS:
        if name in table and table[name] is not None:
            a = np.asarray(table[name])
            # dtypes are kept: the reference's arithmetic depends on them (float32 solar_source_per_gpoint * Python float is a
            # float32 product, cork/sw/component.py:371-372); the library promotes what it needs in fp64 exactly
            out[name] = a if a.dtype in (np.float32, np.float64) else a.astype(np.float64)
    gas_names, _, _, fully_premixed, premixed_bg = table_flags(table)
    out["gas_names"] = ",".join(gas_names)
    for name in K_TABLE_TEXT:
        if name in table:
            out[name] = str(np.asarray(table[name]))
    out["_premixed"] = np.array([1 if (fully_premixed or premixed_bg) else 0], dtype=np.int32)
    out["_co2_logk"] = np.array([1 if CO2_INTERP_LOGK else 0], dtype=np.int32)
    out["_overlap_additive"] = np.array([0 if str(np.asarray(table.get("overlap_method", "additive"))) == "esft" else 1], dtype=np.int32)
    return out


def container_to_ktable(arrays):
    """inverse of ktable_to_container_arrays: the dict `cork.load_k_table` returns for the source file"""
    out = {}
    for k, v in arrays.items():
        if k.startswith("_"):
            continue
        if k == "gas_names":
            out[k] = np.asarray(v.split(","))
        elif isinstance(v, str):
            out[k] = np.asarray(v)
        else:
            out[k] = v
    return out


def convert_k_table(src, dst=None):
    """`.nc` / `.npz` k-table (path or shipped name) -> `.cb2k` next to it (or at dst); returns the path written"""
    from .cork import load_k_table, resolve_k_table_path
    path = resolve_k_table_path(src)
    dst = dst or os.path.splitext(path)[0] + ".cb2k"
    return write_container(dst, ktable_to_container_arrays(load_k_table(path)))


def file_sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def content_sha256(arrays):
    """digest of the arrays' names, dtypes, shapes and bytes -- independent of the container (zip timestamps, entry order)"""
    h = hashlib.sha256()
    for name in sorted(arrays):
        v = arrays[name]
        if isinstance(v, str) or (isinstance(v, np.ndarray) and v.dtype.kind in "US"):
            txt = v if isinstance(v, str) else ",".join(str(x) for x in np.atleast_1d(v))
            h.update(f"{name}|text|{txt}|".encode())
        else:
            a = np.ascontiguousarray(v)
            h.update(f"{name}|{a.dtype.str}|{a.shape}|".encode())
            h.update(a.tobytes())
    return h.hexdigest()


def _shipped_files():
    out = []
    for dp, dn, files in os.walk(DATA_DIR):
        dn[:] = [d for d in dn if d != "_cache"]
        for f in sorted(files):
            if f != "MANIFEST.json":
                out.append(os.path.relpath(os.path.join(dp, f), DATA_DIR))
    return sorted(out)


def build_manifest():
    """sha256 of every shipped data file, content digest + geometry of every k-table, content digest of the derived RRTMG tables
    (the 140 / 112 g-point reduction the engines upload, `rrtmg_tables.reduce_{lw,sw}`)"""
    from . import rrtmg_tables as RT
    from .cork import load_k_table
    man = {"format": 1, "files": {}, "k_tables": {}, "derived": {}}
    for rel in _shipped_files():
        man["files"][rel] = {"sha256": file_sha256(os.path.join(DATA_DIR, rel)), "bytes": os.path.getsize(os.path.join(DATA_DIR, rel))}
    cork_dir = os.path.join(DATA_DIR, "cork")
    for f in sorted(os.listdir(cork_dir)):
        if f.endswith((".npz", ".nc", ".cb2k")):
            t = load_k_table(os.path.join(cork_dir, f))
            k = np.asarray(t["k_coefficients"])
            man["k_tables"][f] = {"content_sha256": content_sha256(ktable_to_container_arrays(t)), "k_shape": list(k.shape),
                                  "k_dtype": k.dtype.name, "gases": [str(g) for g in t.get("gas_names", ["effective"])]}
    for tag, fn in (("rrtmg_lw_reduced", RT.reduce_lw), ("rrtmg_sw_reduced", RT.reduce_sw)):
        red = fn()
        man["derived"][tag] = {"content_sha256": content_sha256(red), "arrays": len(red),
                               "doubles": int(sum(np.asarray(v).size for v in red.values()))}
    return man


def write_manifest():
    with open(MANIFEST, "w") as f:
        json.dump(build_manifest(), f, indent=1, sort_keys=True)
        f.write("\n")
    return MANIFEST


def verify_manifest():
    """-> list of discrepancies between the shipped data and MANIFEST.json (empty = verified)"""
    with open(MANIFEST) as f:
        want = json.load(f)
    have, bad = build_manifest(), []
    for sec in ("files", "k_tables", "derived"):
        for k in sorted(set(want[sec]) | set(have[sec])):
            if want[sec].get(k) != have[sec].get(k):
                bad.append(f"{sec}/{k}: manifest {want[sec].get(k)} != data {have[sec].get(k)}")
    return bad
