"""Init-time table preparation for the CUDA RRTMG engines (host side, numpy).

Does what `rrtmg_lw_ini` / `rrtmg_sw_ini` do once per process in the reference
(climt/_lib/rrtmg_lw/rrtmg_lw_init.f90:28-175 and cmbgb1..16 :366-2015;
climt/_lib/rrtmg_sw/rrtmg_sw_init.f90:47-173 and cmbgb16s..29 :492-1689):
reduce the 16-g-point k-distributions to the 140 (LW) / 112 (SW) g-points used by the GCM
version, with the `rwgt` weights -- but vectorised over the table axes, and re-laid out
for the GPU: **g-point fastest** (one thread walks consecutive g's of a band with
unit-stride, 16-byte-vectorisable loads) and every 1-based Fortran index turned into a
0-based row.  The result is written as one blob (tables.write_blob, C order) that the
native engine uploads to HBM once per engine instance.

This is product code and deliberately independent of oracle/ (which does its own
reduction in C++, following the Fortran loop by loop); tests compare the two.
"""
import os

import numpy as np

from . import tables as _t

# lwcmbdat (rrtmg_lw_init.f90:300-345)
LW_NGC = np.array([10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2])
LW_NGN = np.array(
    [1, 1, 2, 2, 2, 2, 2, 2, 1, 1] + [1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2] + [1] * 16 + [1] * 13 + [3] + [1] * 16
    + [2] * 8 + [2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2] + [2] * 8 + [1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2]
    + [2, 2, 2, 2, 4, 4] + [1, 1, 2, 2, 2, 2, 3, 3] + [1, 1, 1, 1, 2, 2, 4, 4] + [3, 3, 4, 6] + [8, 8] + [8, 8] + [4, 12])
WT = np.array([0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893,
               0.0832767040, 0.0626720116, 0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086,
               0.0022199750, 0.0014140010, 0.0005330000, 0.0000750000])
LW_DELWAVE = np.array([340., 150., 130., 70., 120., 160., 100., 100., 210., 90., 320., 280., 170., 130., 220., 650.])
LW_NSPA = [1, 1, 9, 9, 9, 1, 9, 1, 9, 1, 1, 9, 9, 1, 9, 9]
LW_NSPB = [1, 1, 5, 5, 5, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0, 0]


def _groups(ngn_band):
    """[(start, stop)] of original g-points combined into each reduced g-point."""
    edges = np.concatenate([[0], np.cumsum(ngn_band)])
    assert edges[-1] == 16
    return [(int(edges[i]), int(edges[i + 1])) for i in range(len(ngn_band))]


def _reduce(arr_g_last, groups, weights=None):
    """Sum (optionally weighted by the normalised Gaussian weights) over groups of the last axis,
    accumulating left to right like the Fortran `sumk = sumk + k*rwgt` loops."""
    out = np.zeros(arr_g_last.shape[:-1] + (len(groups),))
    for i, (a, b) in enumerate(groups):
        acc = np.zeros(arr_g_last.shape[:-1])
        for j in range(a, b):
            acc = acc + (arr_g_last[..., j] * weights[j] if weights is not None else arr_g_last[..., j])
        out[..., i] = acc
    return out


def _band_weights(groups):
    """rwgt for one band (rrtmg_lw_init.f90:130-153): wt(ig) / sum of wt over its group; 1 if not reduced."""
    if len(groups) == 16:
        return np.ones(16)
    w = np.zeros(16)
    for a, b in groups:
        s = 0.0
        for j in range(a, b):
            s = s + WT[j]
        w[a:b] = WT[a:b] / s
    return w


def _rows_g(a):
    """(i1, i2, .., g) numpy array (Fortran index order) -> 2-D (rows, g) with Fortran column-major row
    flattening, i.e. row = (i1-1) + n1*((i2-1) + n2*(...)) -- the `absa(ind, ig)` equivalence."""
    ng = a.shape[-1]
    lead = a.shape[:-1]
    if len(lead) <= 1:
        return np.ascontiguousarray(a.reshape(-1, ng))
    perm = tuple(range(len(lead) - 1, -1, -1)) + (len(lead),)
    return np.ascontiguousarray(a.transpose(perm).reshape(-1, ng))


def reduce_lw(raw=None):
    """-> dict of float64 arrays, g fastest.  Names: `b<NN>.<table>`; see DESIGN.md for the layout."""
    raw = raw or _t.load_raw("lw")
    out = {}
    gstart = 0
    off = 0
    meta = np.zeros((16, 4))
    for ib in range(1, 17):
        ngc = int(LW_NGC[ib - 1])
        groups = _groups(LW_NGN[off:off + ngc])
        off += ngc
        w = _band_weights(groups)
        pre = f"rrlw_kg{ib:02d}."
        for key in sorted(k for k in raw if k.startswith(pre) and not k.endswith("__lb")):
            name = key[len(pre):]
            a = raw[key]
            if name.startswith("fracref"):
                red = "fracref" + name[7]             # fracrefao -> fracrefa
                if a.ndim == 1:
                    r = _reduce(a[None, :], groups)    # (1, ng)
                else:                                  # (16, np) -> (np, ng)
                    r = _reduce(np.ascontiguousarray(a.T), groups)
                out[f"b{ib:02d}.{red}"] = np.ascontiguousarray(r)
                continue
            if name.startswith("kao"):
                red = "ka" + name[3:]
            elif name.startswith("kbo"):
                red = "kb" + name[3:]
            else:
                red = name[:-1]                        # selfrefo, forrefo, ccl4o, cfc11adjo, ...
            if a.ndim == 1:
                a = a[None, :]
            r = _reduce(a, groups, w)
            if red == "ka":
                red = "absa"
            elif red == "kb":
                red = "absb"
            out[f"b{ib:02d}.{red}"] = _rows_g(r)
        meta[ib - 1] = (ngc, gstart, LW_NSPA[ib - 1], LW_NSPB[ib - 1])
        gstart += ngc
    # reference profiles and derived ratio tables (setcoef.f90:319-332, 373-377): rat(jp) = chi(a,jp)/chi(b,jp)
    chi = raw["rrlw_ref.chi_mls"]                      # (7, 59), chi[imol-1, jp-1]
    out["chi_mls"] = np.ascontiguousarray(chi)
    out["preflog"] = raw["rrlw_ref.preflog"]
    out["tref"] = raw["rrlw_ref.tref"]
    pairs = [(1, 2), (1, 3), (1, 4), (1, 6), (4, 2), (3, 2)]   # h2oco2, h2oo3, h2on2o, h2och4, n2oco2, o3co2
    out["rat"] = np.stack([chi[a - 1] / chi[b - 1] for a, b in pairs])      # (6, 59)
    out["totplnk"] = np.ascontiguousarray(raw["rrlw_wvn.totplnk"].T)        # (16, 181): band-major
    out["totplnkderiv"] = np.ascontiguousarray(raw["rrlw_wvn.totplnkderiv"].T)
    out["delwave"] = LW_DELWAVE
    out["band_meta"] = meta                            # ng, gstart, nspa, nspb
    # cloud optics (lwcldpr, rrtmg_lw_init.f90:2018-2656)
    out["cld.abscld1"] = np.array([raw["rrlw_cld.abscld1"]]).reshape(1)
    out["cld.absliq0"] = np.array([raw["rrlw_cld.absliq0"]]).reshape(1)
    out["cld.absice0"] = raw["rrlw_cld.absice0"]
    out["cld.absice1"] = np.ascontiguousarray(raw["rrlw_cld.absice1"].T)    # (5, 2)  [ib][k]
    out["cld.absice2"] = np.ascontiguousarray(raw["rrlw_cld.absice2"])      # (43, 16) [index][ib]
    out["cld.absice3"] = np.ascontiguousarray(raw["rrlw_cld.absice3"])      # (46, 16)
    out["cld.absliq1"] = np.ascontiguousarray(raw["rrlw_cld.absliq1"])      # (58, 16)
    return out


# swcmbdat (rrtmg_sw_init.f90:286-340), swdatinit (:195-203)
SW_NGC = np.array([6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12])
SW_NGN = np.array([2, 2, 2, 2, 4, 4] + [1, 1, 1, 1, 1, 2, 1, 2, 1, 2, 1, 2] + [1, 1, 1, 1, 2, 2, 4, 4] * 2
                  + [1, 1, 1, 1, 1, 1, 1, 1, 2, 6] * 2 + [8, 8] + [2, 2, 1, 1, 1, 1, 1, 1, 2, 4] + [2] * 8
                  + [1, 1, 2, 2, 4, 6] * 2 + [1, 1, 1, 1, 1, 1, 4, 6] + [1, 1, 2, 2, 4, 6]
                  + [1, 1, 1, 1, 2, 2, 2, 2, 1, 1, 1, 1])
SW_NSPA = [9, 9, 9, 9, 1, 9, 9, 1, 9, 1, 0, 1, 9, 1]
SW_NSPB = [1, 5, 1, 1, 1, 5, 1, 0, 1, 0, 0, 1, 5, 1]
SW_WAVENUM2 = np.array([3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000., 38000.,
                        50000., 2600.])
_SW_PLAIN = ("sfluxrefo", "irradnceo", "facbrghto", "snsptdrko")   # plain sums (cmbgb16s :579-591)


def reduce_sw(raw=None):
    """-> dict of float64 arrays, g fastest; names `b<NN>.<table>` with NN = 16..29."""
    raw = raw or _t.load_raw("sw")
    out = {}
    off = 0
    for ib in range(1, 15):
        ngc = int(SW_NGC[ib - 1])
        groups = _groups(SW_NGN[off:off + ngc])
        off += ngc
        w = _band_weights(groups)
        pre = f"rrsw_kg{ib + 15:02d}."
        for key in sorted(k for k in raw if k.startswith(pre) and not k.endswith("__lb")):
            name = key[len(pre):]
            a = raw[key]
            if name == "rayl":
                out[f"b{ib + 15:02d}.rayl"] = np.array([float(a)])
                continue
            red = {"kao": "absa", "kbo": "absb"}.get(name, name[:-1])
            plain = name in _SW_PLAIN
            if a.ndim == 1:
                r = _reduce(a[None, :], groups, None if plain else w)          # (1, ng)
            elif plain or name == "raylao":                                    # (16, np): g first -> (np, ng)
                r = _reduce(np.ascontiguousarray(a.T), groups, None if plain else w)
            else:
                r = _rows_g(_reduce(a, groups, w))
            out[f"b{ib + 15:02d}.{red}"] = np.ascontiguousarray(r)
    out["preflog"] = raw["rrsw_ref.preflog"]
    out["tref"] = raw["rrsw_ref.tref"]
    # cloud optics (swcldpr, rrtmg_sw_init.f90:1692-3514): (index, band) tables
    for nm in ("extliq1", "ssaliq1", "asyliq1", "extice2", "ssaice2", "asyice2", "extice3", "ssaice3", "asyice3",
               "fdlice3"):
        out["cld." + nm] = np.ascontiguousarray(raw["rrsw_cld." + nm])       # (n, 14)
    for nm in ("abari", "bbari", "cbari", "dbari", "ebari", "fbari"):
        out["cld." + nm] = raw["rrsw_cld." + nm]
    # ECMWF aerosol optics (swaerpr :389-489): (band, type)
    for nm in ("rsrtaua", "rsrpiza", "rsrasya"):
        out["aer." + nm] = np.ascontiguousarray(raw["rrsw_aer." + nm])       # (14, 6)
    return out


def _blob_path(name, src_name, reduce_fn, rebuild):
    """The reduced tables as the engine's blob.  Derived data: regenerated when older than the raw tables or this file, into
    $CLIMT_B200_CACHE if set, else `data/_cache` next to the package; when neither can be written (a read-only install) the copy
    shipped under `data/` is used as it is (its digest is pinned by data/MANIFEST.json).  The library's reference-named init
    symbols search the same places (csrc/engine_common.h: find_table_blob)."""
    src = os.path.join(_t.DATA_DIR, src_name)
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.abspath(__file__)))
    shipped = os.path.join(_t.DATA_DIR, name)
    for cache in ([os.environ["CLIMT_B200_CACHE"]] if os.environ.get("CLIMT_B200_CACHE") else []) + [os.path.join(_t.DATA_DIR, "_cache")]:
        path = os.path.join(cache, name)
        if not rebuild and os.path.exists(path) and os.path.getmtime(path) >= newest:
            return path
        try:
            os.makedirs(cache, exist_ok=True)
            tmp = f"{path}.{os.getpid()}.tmp"
            _t.write_blob(reduce_fn(), tmp, order="C")
            os.replace(tmp, path)
            return path
        except OSError:
            continue
    if os.path.exists(shipped):
        return shipped
    raise OSError(f"cannot write {name} (set CLIMT_B200_CACHE to a writable directory)")


def sw_blob_path(rebuild=False):
    return _blob_path("rrtmg_sw_reduced.blob", "rrtmg_sw_raw.npz", reduce_sw, rebuild)


def lw_blob_path(rebuild=False):
    return _blob_path("rrtmg_lw_reduced.blob", "rrtmg_lw_raw.npz", reduce_lw, rebuild)


if __name__ == "__main__":
    print(lw_blob_path(rebuild=True))
    print(sw_blob_path(rebuild=True))
