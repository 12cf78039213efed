"""On-disk side of the k-tables: the engine's typed container, the converter from the reference's .npz / .nc
formats, the sha-pinned manifest.  CPU: format round trips and the manifest; GPU: an engine created by the library from a
converted file (no numpy on that path) gives bit-identical fluxes to one created from the reference-format table."""
import os
import struct

import numpy as np
import pytest

import helpers as H
from climt_b200 import cork, table_store as TS

SHIPPED = ("earth_low_res_lw", "earth_low_res_sw", "single_band_gray_lw", "single_band_unit_lw", "test_2band_lw", "test_2band_sw")


def test_container_round_trip_and_alignment(tmp_path):
    arrs = {"a": np.arange(6.0).reshape(2, 3), "b32": np.linspace(0, 1, 5, dtype=np.float32), "i": np.array([3, -1], dtype=np.int64),
            "text": "additive", "empty": np.zeros((0, 4))}
    p = TS.write_container(str(tmp_path / "x.cb2k"), arrs)
    back = TS.read_container(p)
    assert set(back) == set(arrs) and back["text"] == "additive"
    assert back["a"].dtype == np.float64 and back["b32"].dtype == np.float32 and back["i"].dtype == np.int32
    for k in ("a", "b32", "i", "empty"):
        np.testing.assert_array_equal(back[k], arrs[k])
    raw = open(p, "rb").read()
    assert raw[:8] == b"CB2KTB01"
    (n,) = struct.unpack_from("<q", raw, 8)
    for i in range(n):
        off = struct.unpack_from("<48sii7qqq", raw, 16 + 128 * i)[-2]
        assert off % 64 == 0
    with pytest.raises(ValueError):
        open(tmp_path / "bad.cb2k", "wb").write(b"NOTATABLE" + raw[9:])
        TS.read_container(str(tmp_path / "bad.cb2k"))


def _write_reference_formats(table, stem):
    """the table as the reference would hold it on disk: `.npz` (scripts/convert_ck_table_to_npz.py) and NetCDF-3 `.nc`
    (scripts/cork_table_builder/netcdf_writer.py: one dimension per axis, text attributes, gas names as an attribute)"""
    from scipy.io import netcdf_file
    np.savez(stem + ".npz", **table)
    with netcdf_file(stem + ".nc", "w") as nc:
        for name in TS.K_TABLE_ARRAYS:
            if name in table and table[name] is not None:
                a = np.asarray(table[name])
                dims = []
                for i, n in enumerate(a.shape):
                    d = f"{name}_d{i}"
                    nc.createDimension(d, n)
                    dims.append(d)
                v = nc.createVariable(name, a.dtype.char, tuple(dims))
                v[:] = a
        nc.gas_names = ",".join(str(g) for g in table["gas_names"]) if "gas_names" in table else "effective"
        for name in TS.K_TABLE_TEXT:
            if name in table:
                setattr(nc, name, str(np.asarray(table[name])))
    return stem + ".npz", stem + ".nc"


@pytest.mark.parametrize("name", SHIPPED)
def test_converted_table_equals_the_reference_format_table(tmp_path, name):
    """reference formats -> container: arrays, dtypes and the classification survive, whichever format the table came in"""
    shipped = cork.load_k_table(name)                                # the shipped tables are containers themselves
    for src_path in _write_reference_formats(shipped, str(tmp_path / name)):
        src = cork.load_k_table(src_path)
        dst = TS.convert_k_table(src_path, str(tmp_path / (name + "_out.cb2k")))
        back = cork.load_k_table(dst)                                # load_k_table reads the container too
        assert np.asarray(back["k_coefficients"]).dtype == np.asarray(src["k_coefficients"]).dtype   # float32 tables stay float32
        for k in TS.K_TABLE_ARRAYS:
            assert (k in src and src[k] is not None) == (k in back), k
            if k in back:
                np.testing.assert_array_equal(back[k], src[k])
                np.testing.assert_array_equal(back[k], shipped[k])
                if np.asarray(src[k]).dtype.kind == "f":
                    assert np.asarray(back[k]).dtype == np.asarray(src[k]).dtype, k   # the reference's float32 products depend on it
        assert cork.table_flags(back) == cork.table_flags(src) == cork.table_flags(shipped)
        raw = TS.read_container(dst)
        _, _, _, fully, bg = cork.table_flags(src)
        assert int(raw["_premixed"][0]) == int(fully or bg) and int(raw["_co2_logk"][0]) == 1 and int(raw["_overlap_additive"][0]) == 1
        # the content digest does not depend on the container the table came from
        assert TS.content_sha256(TS.ktable_to_container_arrays(back)) == TS.content_sha256(TS.ktable_to_container_arrays(shipped))


def test_manifest_pins_every_shipped_and_derived_table():
    assert TS.verify_manifest() == []
    man = TS.build_manifest()
    assert set(f"{n}.cb2k" for n in SHIPPED) <= set(man["k_tables"])
    assert {"mars_lw.cb2k", "titan_sw.cb2k", "trappist1e_hab1_lw.cb2k"} <= set(man["k_tables"])
    assert man["k_tables"]["earth_low_res_lw.cb2k"]["k_shape"] == [1, 14, 8, 12, 8, 7, 10]
    assert man["k_tables"]["earth_low_res_lw.cb2k"]["k_dtype"] == "float32"
    assert {"rrtmg_lw_raw.npz", "rrtmg_sw_raw.npz", "ozone_profile.npy", "berger1978.npz"} <= set(man["files"])
    assert man["derived"]["rrtmg_lw_reduced"]["arrays"] > 100


def test_manifest_detects_a_changed_table(tmp_path, monkeypatch):
    import shutil
    d = tmp_path / "data"
    shutil.copytree(TS.DATA_DIR, d, ignore=shutil.ignore_patterns("_cache"))
    t = TS.read_container(str(d / "cork" / "test_2band_lw.cb2k"))
    t["k_coefficients"] = (t["k_coefficients"] * 1.0000001).astype(t["k_coefficients"].dtype)
    TS.write_container(str(d / "cork" / "test_2band_lw.cb2k"), t)
    monkeypatch.setattr(TS, "DATA_DIR", str(d))
    monkeypatch.setattr(TS, "MANIFEST", str(d / "MANIFEST.json"))
    monkeypatch.setattr(cork, "_DATA", str(d / "cork"))
    bad = TS.verify_manifest()
    assert any("k_tables/test_2band_lw.cb2k" in b for b in bad) and any("files/cork/test_2band_lw.cb2k" in b for b in bad)
    assert not any("earth_low_res" in b for b in bad)


@pytest.mark.gpu
@pytest.mark.parametrize("name,which", [("earth_low_res_lw", "lw"), ("earth_low_res_sw", "sw"), ("test_2band_lw", "lw")])
def test_engine_created_from_file_matches_engine_created_from_table(tmp_path, name, which):
    from climt_b200 import synthetic as SY
    dst = TS.convert_k_table(name, str(tmp_path / f"{name}.cb2k"))
    ncol, nlev = 640, 40
    rng = np.random.default_rng(11)
    lw = SY.make_lw_state(ncol, nlev, seed=3)
    a, b = cork.CorkEngine(name, device=0), cork.CorkEngine(dst, device=0)
    nb = a.nband
    s = {"T": lw["tlay"], "p": lw["play"] * 100.0, "p_int": lw["plev"] * 100.0, "T_surf": lw["tsfc"], "q": lw["h2o"] * 0.622,
         "co2": np.full((nlev, ncol), 4e-4), "emissivity": rng.uniform(0.9, 1.0, (nb, ncol)), "tau_cloud_lw": np.zeros((nlev, ncol, nb)),
         "zenith": np.deg2rad(rng.uniform(0, 88, ncol)), "albedo": rng.uniform(0.05, 0.3, ncol), "earth_sun_factor": np.full(ncol, 1.01),
         "tau_cloud_sw": np.zeros((nlev, ncol, nb)), "ssa_cloud": np.zeros((nlev, ncol, nb)), "g_cloud": np.zeros((nlev, ncol, nb))}
    arr = H.cork_arrays(s, which)
    if a.ctable.premixed == 0:
        arr["gas_q"] = rng.uniform(1e-6, 1e-3, (a.ngas, nlev, ncol))
    if which == "lw":
        ra, rb = a.lw_host(ncol, nlev, arr), b.lw_host(ncol, nlev, arr)
    else:
        ra, rb = a.sw_host(ncol, nlev, arr, earth_sun_factor=1.01), b.sw_host(ncol, nlev, arr, earth_sun_factor=1.01)
    assert float(np.abs(ra["up_broad"]).max()) > 0
    for k in ra:
        np.testing.assert_array_equal(ra[k], rb[k])
    a.close()
    b.close()


@pytest.mark.gpu
def test_create_from_file_rejects_bad_files(tmp_path):
    import ctypes
    from climt_b200 import _native
    L = _native.lib()
    L.cb200_cork_create_from_file.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p] + [ctypes.c_double] * 3 + [ctypes.c_int]
    h = ctypes.c_void_p()
    assert L.cb200_cork_create_from_file(ctypes.byref(h), str(tmp_path / "missing.cb2k").encode(), 9.8, 1004.0, 5.67e-8, 0) != 0
    assert b"cannot open" in L.cb200_global_error()
    good = TS.convert_k_table("test_2band_lw", str(tmp_path / "t.cb2k"))
    raw = open(good, "rb").read()
    open(tmp_path / "trunc.cb2k", "wb").write(raw[: len(raw) // 2])
    assert L.cb200_cork_create_from_file(ctypes.byref(h), str(tmp_path / "trunc.cb2k").encode(), 9.8, 1004.0, 5.67e-8, 0) != 0
    assert b"bad entry" in L.cb200_global_error() or b"truncated" in L.cb200_global_error()


def test_shipped_reduced_blobs_are_what_the_reduction_gives(tmp_path, monkeypatch):
    """The reduced-table blobs shipped under climt_b200/data/ (for binders that never run the Python side: the reference-named
    init symbols fall back to them) equal a fresh reduction written to a user cache directory (CLIMT_B200_CACHE)."""
    import os
    from climt_b200 import rrtmg_tables as RT, tables as TB
    monkeypatch.setenv("CLIMT_B200_CACHE", str(tmp_path))
    for which in ("lw", "sw"):
        fresh = getattr(RT, which + "_blob_path")(rebuild=True)
        assert os.path.dirname(fresh) == str(tmp_path)
        shipped = os.path.join(TB.DATA_DIR, f"rrtmg_{which}_reduced.blob")
        assert open(fresh, "rb").read() == open(shipped, "rb").read()
