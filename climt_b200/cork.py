"""CORK radiation -- drop-ins for climt.CorkLongwaveRadiation / climt.CorkShortwaveRadiation.

Mirrors climt/_components/cork/lw/component.py:20-466 and cork/sw/component.py:20-532 for ``optics="correlated_k"``
(additive overlap) and ``optics="parmentier"`` (picket-fence analytic optics, the constructors' default): same
constructor arguments, property dictionaries, aliases, units and ``array_call`` return values.  The optical depths
(k-table interpolation or Freedman/Parmentier fits), Planck sources, transport sweeps, flux sums and heating rates
run in the CUDA engine (csrc/cork_engine.cu) behind the C ABI of include/climt_b200.h; there is no CPU
implementation here.

ESFT-overlap tables are evaluated as additive tables on the combined g-points (expand_esft_table); ``diagnostics_level >= 1``
returns the reference's per-band g-point averages of the kernel diagnostics (cb200_cork_set_diagnostics).  (Round 1 rejected both.)
(per-g-point diagnostic dumps).
"""
import ctypes
import os

import numpy as np

from . import _native
from .constants import get_constant
from .sympl_shim import TendencyComponent

_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "cork")
DIFFUSIVITY_FACTOR = 1.66  # cork/lw/kernels.py:6
CO2_INTERP_LOGK = True     # cork/optics/correlated_k.py:27

_NETCDF_VARS = ("k_coefficients", "gpoint_weights", "temperature_grid", "pressure_grid_log", "h2o_vmr_grid", "co2_vmr_grid",
                "band_wavenumber_limits", "planck_fraction", "solar_source_per_gpoint", "rayleigh_coefficient", "continuum_kappa")


def _decode(x):
    if isinstance(x, bytes):
        return x.decode("utf-8")
    if isinstance(x, np.ndarray) and x.dtype.kind == "S":
        return x.tobytes().decode("utf-8").rstrip("\x00")
    return str(x)


def _load_netcdf_table(path):
    from scipy.io import netcdf_file
    out = {}
    with netcdf_file(path, "r", mmap=False) as nc:
        for name in _NETCDF_VARS:
            if name in nc.variables:
                arr = np.asarray(nc.variables[name][:]).copy()
                if arr.dtype.byteorder not in ("=", "|"):
                    arr = arr.astype(arr.dtype.newbyteorder("="))
                out[name] = arr
        if "gas_names" in nc.variables:
            out["gas_names"] = np.asarray([_decode(x) for x in np.atleast_1d(nc.variables["gas_names"][:])])
        elif getattr(nc, "gas_names", None) is not None:
            out["gas_names"] = np.asarray([s.strip() for s in _decode(nc.gas_names).split(",") if s.strip()])
        for attr in ("overlap_method", "resolution", "background_is_premixed"):
            val = getattr(nc, attr, None)
            if val is not None:
                out[attr] = np.asarray(_decode(val))
    return out


def resolve_k_table_path(name_or_path):
    """a path, or the name of a table shipped under ``climt_b200/data/cork/`` (.npz, then .nc like the reference, then .cb2k)"""
    path = os.fspath(name_or_path)
    if os.path.isfile(path):
        return path
    for ext in (".npz", ".nc", ".cb2k"):
        cand = os.path.join(_DATA, f"{name_or_path}{ext}")
        if os.path.isfile(cand):
            return cand
    raise FileNotFoundError(f"No k-table named {name_or_path!r} (.npz or .nc)")


def load_k_table(name_or_path):
    """Same contract as the reference's load_k_table (cork/optics/correlated_k.py:186-218): a path to a ``.npz`` /
    ``.nc`` file, or the name of a table shipped under ``climt_b200/data/cork/``; returns a dict of arrays.
    Also reads the engine's own container (``.cb2k``, climt_b200/table_store.py)."""
    path = resolve_k_table_path(name_or_path)
    if path.endswith(".cb2k"):
        from . import table_store
        return table_store.container_to_ktable(table_store.read_container(path))
    if path.endswith(".nc"):
        return _load_netcdf_table(path)
    with np.load(path, allow_pickle=True) as z:
        return {k: z[k] for k in z.files}


class CorkTable(ctypes.Structure):
    """cb200_cork_table (include/climt_b200.h)."""
    _fields_ = [(n, ctypes.c_int) for n in ("ngas", "nband", "ngpt", "nT", "nP", "nX", "nC")] + [
        ("k_coefficients_f32", _fp), ("k_coefficients_f64", _dp), ("temperature_grid", _dp), ("pressure_grid_log", _dp),
        ("h2o_vmr_grid", _dp), ("co2_vmr_grid", _dp), ("gpoint_weights", _dp), ("planck_fraction", _dp),
        ("nband_pf", ctypes.c_int), ("ngpt_pf", ctypes.c_int), ("continuum_kappa", _dp), ("solar_source_per_gpoint", _dp),
        ("rayleigh_coefficient", _dp), ("co2_logk", ctypes.c_int), ("premixed", ctypes.c_int)]


CORK_IN = ("T", "p", "p_int", "T_surf", "q_h2o", "co2_vmr", "gas_q", "emissivity", "tau_cloud", "zenith", "albedo", "ssa_cloud", "g_cloud",
           "T_irr", "T_int", "bond_albedo")
CORK_OUT = ("up_broad", "down_broad", "heating_rate", "up_band", "down_band", "tau_band", "trans_band", "hr_band")


class CorkInputs(ctypes.Structure):
    _fields_ = [(n, _dp) for n in CORK_IN]


class CorkOutputs(ctypes.Structure):
    _fields_ = [(n, _dp) for n in CORK_OUT]


PICKET_MAX_REGIONS = 8


class CorkDiagnostics(ctypes.Structure):
    """cb200_cork_diagnostics (include/climt_b200.h)."""
    _fields_ = [("level", ctypes.c_int), ("field", _dp * 10), ("weight_sum", _dp)]


# diagnostics_level >= 1: field index in cb200_cork_diagnostics -> (component diagnostic, interface-level field?, minimum level)
LW_DIAG = {0: ("lw_layer_transmittance", False, 1), 1: ("lw_up_per_gpoint", True, 1), 2: ("lw_down_per_gpoint", True, 1)}
SW_DIAG = {0: ("sw_layer_diffuse_reflectance", False, 1), 1: ("sw_layer_diffuse_transmittance", False, 1),
           2: ("sw_layer_direct_transmittance", False, 1), 3: ("sw_direct_beam_profile", True, 1),
           4: ("sw_layer_direct_reflectance", False, 2), 5: ("sw_layer_direct_source_transmittance", False, 2),
           6: ("sw_delta_scaled_optical_depth", False, 2), 7: ("sw_delta_scaled_ssa", False, 2),
           8: ("sw_delta_scaled_asymmetry", False, 2), 9: ("sw_combined_albedo", True, 2)}


class PicketCoeffs(ctypes.Structure):
    """cb200_picket_coeffs (include/climt_b200.h)."""
    _ab = ctypes.c_double * 2 * PICKET_MAX_REGIONS
    _fields_ = [("nregion", ctypes.c_int), ("T_eff_boundaries", ctypes.c_double * (PICKET_MAX_REGIONS + 1)),
                ("log10_gamma_v1_ab", _ab), ("log10_gamma_v2_ab", _ab), ("log10_gamma_v3_ab", _ab), ("beta_ab", _ab),
                ("log10_gamma_P_quad", ctypes.c_double * 3)] + [
        (n, ctypes.c_double) for n in ("T_boundary", "a_hi", "b_hi", "c_hi", "a_lo", "b_lo", "c_lo")]


def _load_npz(kind, name_or_path, package_dir):
    path = name_or_path if os.path.isfile(name_or_path) else os.path.join(_DATA, package_dir, f"{name_or_path}.npz")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"No {kind} named {name_or_path!r}")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def load_parmentier_coefficients(name_or_path):
    """cork/optics/parmentier.py:77-97: "solar_composition" (shipped) or the path of an .npz file."""
    return _load_npz("Parmentier coefficient table", name_or_path, "parmentier")


def load_freedman2014_coefficients():
    """cork/optics/parmentier.py:36-52."""
    return _load_npz("Rosseland-mean fit", "freedman2014", "parmentier")


def load_stellar_spectrum(name_or_path):
    """cork/optics/stellar.py:7-29: "sun", "trappist1" or a path -> wavenumber [cm-1], irradiance [W m-2 / cm-1]."""
    z = _load_npz("stellar spectrum", name_or_path, "stellar_spectra")
    return {"wavenumber": np.array(z["wavenumber"]), "irradiance": np.array(z["irradiance"])}


def integrate_spectrum_over_bands(spectrum, band_wavenumber_limits):
    """cork/optics/stellar.py:32-62 (trapezoid over the band, end points interpolated); host-side, constructor only."""
    wn, irr = spectrum["wavenumber"], spectrum["irradiance"]
    flux = np.zeros(band_wavenumber_limits.shape[0])
    for b, (wn_lo, wn_hi) in enumerate(band_wavenumber_limits):
        mask = (wn > wn_lo) & (wn < wn_hi)
        wn_band = np.concatenate(([wn_lo], wn[mask], [wn_hi]))
        irr_band = np.concatenate(([np.interp(wn_lo, wn, irr)], irr[mask], [np.interp(wn_hi, wn, irr)]))
        flux[b] = np.trapezoid(irr_band, wn_band)
    return flux


def make_picket_coeffs(coefficients, freedman):
    c = PicketCoeffs()
    bounds = np.asarray(coefficients["T_eff_boundaries"], dtype=np.float64)
    nreg = len(bounds) - 1
    if not 1 <= nreg <= PICKET_MAX_REGIONS:
        raise ValueError(f"Parmentier coefficient table: 1..{PICKET_MAX_REGIONS} T_eff regions supported, got {nreg}")
    c.nregion = nreg
    for i, v in enumerate(bounds):
        c.T_eff_boundaries[i] = float(v)
    for name in ("log10_gamma_v1_ab", "log10_gamma_v2_ab", "log10_gamma_v3_ab", "beta_ab"):
        ab = np.asarray(coefficients[name], dtype=np.float64)
        if ab.shape != (nreg, 2):
            raise ValueError(f"{name}: expected shape {(nreg, 2)}, got {ab.shape}")
        dst = getattr(c, name)
        for i in range(nreg):
            dst[i][0], dst[i][1] = float(ab[i, 0]), float(ab[i, 1])
    for i in range(3):
        c.log10_gamma_P_quad[i] = float(np.asarray(coefficients["log10_gamma_P_quad"])[i])
    for n in ("T_boundary", "a_hi", "b_hi", "c_hi", "a_lo", "b_lo", "c_lo"):
        setattr(c, n, float(freedman[n]))
    return c


def table_flags(table):
    """The table-classification logic of the reference constructors (cork/lw/component.py:46-59)."""
    gas_names = [str(g) for g in table["gas_names"]] if "gas_names" in table else ["effective"]
    has_h2o = "h2o_vmr_grid" in table
    has_co2 = "co2_vmr_grid" in table
    fully_premixed = gas_names == ["effective"] and not has_h2o
    premixed_bg = (gas_names == ["effective"] and has_h2o) or str(table.get("background_is_premixed", np.array(""))).lower() == "true"
    return gas_names, has_h2o, has_co2, fully_premixed, premixed_bg


def _bind(L):
    vp = ctypes.c_void_p
    L.cb200_cork_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(CorkTable), ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_int]
    L.cb200_cork_create_picket.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(PicketCoeffs), ctypes.c_int, ctypes.c_double,
                                           ctypes.c_double, ctypes.c_double, ctypes.c_int]
    L.cb200_cork_destroy.argtypes = [vp]
    L.cb200_cork_destroy.restype = None
    L.cb200_cork_last_error.argtypes = [vp]
    L.cb200_cork_last_error.restype = ctypes.c_char_p
    L.cb200_cork_last_launches.argtypes = [vp]
    L.cb200_cork_enable_timing.argtypes = [vp, ctypes.c_int]
    L.cb200_cork_last_unit_kernel_ms.argtypes = [vp]
    L.cb200_cork_last_unit_kernel_ms.restype = ctypes.c_double
    pi, po = ctypes.POINTER(CorkInputs), ctypes.POINTER(CorkOutputs)
    L.cb200_cork_lw_run_device.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, pi, po, vp]
    L.cb200_cork_sw_run_device.argtypes = [vp, ctypes.c_int, ctypes.c_int, _dp, pi, po, vp]
    L.cb200_cork_lw_run_host.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, pi, po]
    L.cb200_cork_sw_run_host.argtypes = [vp, ctypes.c_int, ctypes.c_int, _dp, pi, po]
    L.cb200_cork_set_diagnostics.argtypes = [vp, ctypes.POINTER(CorkDiagnostics)]


def esft_weights(gpoint_weights, ngas):
    """Combined g-point weights of the ESFT overlap (cork/optics/correlated_k.py:343-372): for N gases with G g-points each,
    G**N weights per band, the products of the per-gas weights multiplied in gas order (digit `gas` of the combined index in
    base G selects that gas's g-point) -- the same multiplication order as the reference, so the same bits."""
    w = np.asarray(gpoint_weights, dtype=np.float64)
    nband, ngpt = w.shape
    idx = np.arange(ngpt ** ngas)
    out = np.ones((nband, idx.size))
    rem = idx.copy()
    for _ in range(ngas):
        out = out * w[:, rem % ngpt]
        rem //= ngpt
    return out


def is_esft(table):
    return str(table.get("overlap_method", np.array("additive"))) == "esft"


def expand_esft_table(table):
    """An ESFT-overlap table (cork/optics/correlated_k.py:564-594: every combination of one g-point per gas is a g-point of the
    band, optical depths add over the gases) restated as an additive-overlap table on the G**N combined g-points, which is what
    the engine evaluates:
      * k[gas, band, idx] = k[gas, band, digit_gas(idx)] (base-G digits, gas 0 the lowest) -- the engine's sum over gases, in gas
        order, is then term for term the reference's `tau[ib, idx] += k_interp[ig, ib, g_idx] * amount[ig]`;
      * weights = the products of the per-gas weights (`compute_esft_weights`, :343-372);
      * Planck fractions and the solar source are looked up at `idx % G` (cork/lw/kernels.py:43,65; cork/sw/component.py:373-385);
      * the band-grey continuum is not part of the ESFT path (:564-594 never reads it) and is dropped.
    With one gas the expansion is the identity (apart from the continuum)."""
    k = np.asarray(table["k_coefficients"])
    ngas, nband, ngpt = k.shape[:3]
    ncomb = ngpt ** ngas
    idx = np.arange(ncomb)
    out = dict(table)
    kx = np.empty((ngas, nband, ncomb) + k.shape[3:], dtype=k.dtype)
    for ig in range(ngas):
        kx[ig] = k[ig][:, (idx // ngpt ** ig) % ngpt]
    out["k_coefficients"] = kx
    out["gpoint_weights"] = esft_weights(table["gpoint_weights"], ngas)
    for name in ("planck_fraction", "solar_source_per_gpoint"):
        if name in table and table[name] is not None:
            a = np.asarray(table[name])
            out[name] = np.ascontiguousarray(a[:, idx % a.shape[1]])
    out.pop("continuum_kappa", None)
    out["overlap_method"] = np.array("additive")
    out["_esft_expanded_from"] = np.array([ngas, ngpt])
    return out


def make_ctable(table, co2_logk=CO2_INTERP_LOGK):
    """dict of arrays -> (cb200_cork_table, keep-alive list).  float32 k stays float32; grids are promoted to float64 exactly."""
    if is_esft(table):
        table = expand_esft_table(table)
    k = np.asarray(table["k_coefficients"])
    if k.ndim not in (5, 6, 7):
        raise ValueError(f"k_coefficients must have 5, 6 or 7 dimensions, got {k.ndim}")
    gas_names, has_h2o, has_co2, fully_premixed, premixed_bg = table_flags(table)
    keep = []

    def d64(name):
        if name not in table or table[name] is None:
            return None
        a = np.ascontiguousarray(table[name], dtype=np.float64)
        keep.append(a)
        return a.ctypes.data_as(_dp)

    t = CorkTable()
    t.ngas, t.nband, t.ngpt, t.nT, t.nP = k.shape[:5]
    t.nX = k.shape[5] if k.ndim >= 6 else 0
    t.nC = k.shape[6] if k.ndim == 7 else 0
    if k.dtype == np.float32:
        kk = np.ascontiguousarray(k)
        t.k_coefficients_f32 = kk.ctypes.data_as(_fp)
    else:
        kk = np.ascontiguousarray(k, dtype=np.float64)
        t.k_coefficients_f64 = kk.ctypes.data_as(_dp)
    keep.append(kk)
    t.temperature_grid = d64("temperature_grid")
    t.pressure_grid_log = d64("pressure_grid_log")
    t.h2o_vmr_grid = d64("h2o_vmr_grid") if t.nX else None
    t.co2_vmr_grid = d64("co2_vmr_grid") if t.nC else None
    t.gpoint_weights = d64("gpoint_weights")
    if "planck_fraction" in table:
        pf = np.asarray(table["planck_fraction"])
        t.planck_fraction = d64("planck_fraction")
        t.nband_pf, t.ngpt_pf = pf.shape[0], pf.shape[1]
    cont = table.get("continuum_kappa")
    if cont is not None and np.asarray(cont).ndim == 4 and t.nX:
        t.continuum_kappa = d64("continuum_kappa")
    if "solar_source_per_gpoint" in table:
        t.solar_source_per_gpoint = d64("solar_source_per_gpoint")
    if table.get("rayleigh_coefficient") is not None:
        t.rayleigh_coefficient = d64("rayleigh_coefficient")
    t.co2_logk = 1 if co2_logk else 0
    t.premixed = 1 if (fully_premixed or premixed_bg) else 0
    return t, keep


class CorkEngine:
    """Handle of one cb200_cork_engine (one k-table resident in HBM)."""

    def __init__(self, table, g=None, cpd=None, sigma=None, device=0, picket=None):
        """table: a k-table (name, path or dict); or None with picket=("lw" | "sw", coefficients, freedman) for the
        picket-fence engines (cb200_cork_create_picket)."""
        self._L = _native.lib()
        _bind(self._L)
        g = get_constant("gravitational_acceleration", "m/s^2") if g is None else g
        cpd = get_constant("heat_capacity_of_dry_air_at_constant_pressure", "J/kg/K") if cpd is None else cpd
        sigma = get_constant("stefan_boltzmann_constant", "W/m^2/K^4") if sigma is None else sigma
        self.sigma = sigma
        self._h = ctypes.c_void_p()
        if picket is not None:
            which, coefficients, freedman = picket
            self.table = None
            self.picket = make_picket_coeffs(coefficients, freedman)
            if self._L.cb200_cork_create_picket(ctypes.byref(self._h), ctypes.byref(self.picket), 1 if which == "lw" else 0,
                                                g, cpd, sigma, device):
                raise RuntimeError(self._L.cb200_global_error().decode())
            self.nband, self.ngpt, self.ngas = (2 if which == "lw" else 3), 1, 1
            return
        self.table = load_k_table(table) if isinstance(table, (str, os.PathLike)) else table
        esft = is_esft(self.table)
        if esft:  # evaluated as an additive table on the combined g-points (expand_esft_table)
            self.table = expand_esft_table(self.table)
        self.ctable, self._keep = make_ctable(self.table)
        if not esft and isinstance(table, (str, os.PathLike)) and os.fspath(table).endswith(".cb2k"):
            # the engine's own container: the library reads, classifies and re-lays out the file itself (no numpy on this path)
            self._L.cb200_cork_create_from_file.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p] + [ctypes.c_double] * 3 + [ctypes.c_int]
            rc = self._L.cb200_cork_create_from_file(ctypes.byref(self._h), os.fspath(table).encode(), g, cpd, sigma, device)
        else:
            rc = self._L.cb200_cork_create(ctypes.byref(self._h), ctypes.byref(self.ctable), g, cpd, sigma, device)
        if rc:
            raise RuntimeError(self._L.cb200_global_error().decode())
        self.nband, self.ngpt, self.ngas = self.ctable.nband, self.ctable.ngpt, self.ctable.ngas

    def close(self):
        if getattr(self, "_h", None):
            self._L.cb200_cork_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return self._L.cb200_cork_last_error(self._h).decode()

    def shapes(self, ncol, nlev):
        L, n, nb = nlev, ncol, self.nband
        ins = {"T": (L, n), "p": (L, n), "p_int": (L + 1, n), "T_surf": (n,), "q_h2o": (L, n), "co2_vmr": (L, n),
               "gas_q": (self.ngas, L, n), "emissivity": (nb, n), "tau_cloud": (L, n, nb), "zenith": (n,), "albedo": (n,),
               "ssa_cloud": (L, n, nb), "g_cloud": (L, n, nb), "T_irr": (n,), "T_int": (n,), "bond_albedo": (n,)}
        outs = {"up_broad": (L + 1, n), "down_broad": (L + 1, n), "heating_rate": (L, n), "up_band": (nb, L + 1, n),
                "down_band": (nb, L + 1, n), "tau_band": (nb, L, n), "trans_band": (nb, L, n), "hr_band": (nb, L, n)}
        return ins, outs

    def _pack_host(self, ncol, nlev, arrays, out, which, bands):
        ins, outs = self.shapes(ncol, nlev)
        keep, pin = [], CorkInputs()
        for k in CORK_IN:
            a = arrays.get(k)
            if a is None:
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != ins[k]:
                raise ValueError(f"{k}: expected shape {ins[k]}, got {a.shape}")
            keep.append(a)
            setattr(pin, k, a.ctypes.data_as(_dp))
        names = ["up_broad", "down_broad", "heating_rate"]
        if bands:
            names += ["up_band", "down_band", "tau_band", "hr_band"] + (["trans_band"] if which == "lw" else [])
        out = out if out is not None else {k: np.empty(outs[k]) for k in names}
        pout = CorkOutputs()
        for k, a in out.items():
            if a.shape != outs[k] or a.dtype != np.float64 or not a.flags.c_contiguous:
                raise ValueError(f"output {k}: need C-contiguous float64 {outs[k]}")
            setattr(pout, k, a.ctypes.data_as(_dp))
        return pin, pout, out, keep

    def _host_diagnostics(self, which, level, ncol, nlev):
        """diagnostics_level >= 1: numpy fields for the next host call -> {component diagnostic name: (nband, nlev[+1], ncol)}"""
        d, fields = CorkDiagnostics(), {}
        d.level = int(level)
        for j, (name, iface, minlevel) in (LW_DIAG if which == "lw" else SW_DIAG).items():
            if level >= minlevel:
                a = np.zeros((self.nband, nlev + 1 if iface else nlev, ncol))
                fields[name] = a
                d.field[j] = a.ctypes.data_as(_dp)
        if level > 0 and self.table is not None:  # `weights.sum(axis=1)` in the table's own dtype (cork/lw/component.py:342)
            wsum = np.ascontiguousarray(np.asarray(self.table["gpoint_weights"]).sum(axis=1), dtype=np.float64)
            d.weight_sum = wsum.ctypes.data_as(_dp)
        if self._L.cb200_cork_set_diagnostics(self._h, ctypes.byref(d) if level > 0 else None):
            raise RuntimeError(self._err())
        return fields

    def lw_host(self, ncol, nlev, arrays, out=None, diffusivity_factor=DIFFUSIVITY_FACTOR, bands=True, diagnostics_level=0):
        """diagnostics_level >= 1: returns (out, {diagnostic name: (nband, nlev[+1], ncol)})"""
        pin, pout, out, keep = self._pack_host(ncol, nlev, arrays, out, "lw", bands)
        fields = self._host_diagnostics("lw", diagnostics_level, ncol, nlev)
        rc = self._L.cb200_cork_lw_run_host(self._h, ncol, nlev, float(diffusivity_factor), ctypes.byref(pin), ctypes.byref(pout))
        if diagnostics_level:
            self._L.cb200_cork_set_diagnostics(self._h, None)
        if rc:
            raise (ValueError if rc == -3 else RuntimeError)(self._err())
        return (out, fields) if diagnostics_level else out

    def solar_flux(self, earth_sun_factor):
        """solar_source_per_gpoint * earth_sun_factor with numpy's dtype rules, as the reference evaluates it
        (cork/sw/component.py:371-372: a float32 table gives a float32 product), handed to the engine as float64."""
        if self.table is None or "solar_source_per_gpoint" not in self.table:
            raise ValueError("cork: this table has no solar_source_per_gpoint (not a shortwave table)")
        return np.ascontiguousarray(np.asarray(self.table["solar_source_per_gpoint"]) * float(earth_sun_factor), dtype=np.float64)

    def sw_host(self, ncol, nlev, arrays, out=None, earth_sun_factor=1.0, bands=True, solar_flux=None, diagnostics_level=0):
        """solar_flux: (nband, ngpt) W m-2 already scaled (picket-fence engines: mandatory); else the table's * earth_sun_factor.
        diagnostics_level >= 1: returns (out, {diagnostic name: (nband, nlev[+1], ncol)})"""
        pin, pout, out, keep = self._pack_host(ncol, nlev, arrays, out, "sw", bands)
        sf = self.solar_flux(earth_sun_factor) if solar_flux is None else np.ascontiguousarray(solar_flux, dtype=np.float64)
        if sf.shape != (self.nband, self.ngpt):
            raise ValueError(f"solar_flux: expected shape {(self.nband, self.ngpt)}, got {sf.shape}")
        fields = self._host_diagnostics("sw", diagnostics_level, ncol, nlev)
        rc = self._L.cb200_cork_sw_run_host(self._h, ncol, nlev, sf.ctypes.data_as(_dp), ctypes.byref(pin), ctypes.byref(pout))
        if diagnostics_level:
            self._L.cb200_cork_set_diagnostics(self._h, None)
        if rc:
            raise (ValueError if rc == -3 else RuntimeError)(self._err())
        return (out, fields) if diagnostics_level else out

    def _run_device(self, fn, ncol, nlev, scalar, tensors, out, stream):
        """scalar: float (lw diffusivity) or a float64 numpy (nband, ngpt) solar flux (sw)"""
        import torch
        ins, outs = self.shapes(ncol, nlev)
        pin, pout = CorkInputs(), CorkOutputs()
        for k, t in tensors.items():
            if t is None:
                continue
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == ins[k]):
                raise ValueError(f"{k}: need contiguous float64 CUDA tensor of shape {ins[k]}")
            setattr(pin, k, ctypes.cast(t.data_ptr(), _dp))
        for k, t in out.items():
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == outs[k]):
                raise ValueError(f"output {k}: need contiguous float64 CUDA tensor of shape {outs[k]}")
            setattr(pout, k, ctypes.cast(t.data_ptr(), _dp))
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        arg = scalar.ctypes.data_as(_dp) if isinstance(scalar, np.ndarray) else float(scalar)
        rc = fn(self._h, ncol, nlev, arg, ctypes.byref(pin), ctypes.byref(pout), ctypes.c_void_p(s))
        if rc:
            raise (ValueError if rc == -3 else RuntimeError)(self._err())

    def lw_device(self, ncol, nlev, tensors, out, diffusivity_factor=DIFFUSIVITY_FACTOR, stream=None):
        self._run_device(self._L.cb200_cork_lw_run_device, ncol, nlev, diffusivity_factor, tensors, out, stream)

    def sw_device(self, ncol, nlev, tensors, out, earth_sun_factor=1.0, stream=None, solar_flux=None):
        sf = self.solar_flux(earth_sun_factor) if solar_flux is None else np.ascontiguousarray(solar_flux, dtype=np.float64)
        self._run_device(self._L.cb200_cork_sw_run_device, ncol, nlev, sf, tensors, out, stream)

    def enable_timing(self, on=True):
        self._L.cb200_cork_enable_timing(self._h, 1 if on else 0)

    @property
    def last_unit_kernel_ms(self):
        return self._L.cb200_cork_last_unit_kernel_ms(self._h)

    @property
    def last_launches(self):
        return self._L.cb200_cork_last_launches(self._h)


# ---------------------------------------------------------------------------------------------------------------------
MOLAR_MASS_DRY_AIR = 28.970  # cork/common.py:9-17
MOLAR_MASS = {"h2o": 18.015, "co2": 44.010, "o3": 47.998, "ch4": 16.043, "n2o": 44.013, "o2": 31.998}
_GAS_CF_NAME = {"h2o": "specific_humidity", "co2": "mole_fraction_of_carbon_dioxide_in_air"}
_num_bands = {"lw": None, "sw": None}  # set_num_{long,short}wave_bands (climt/_core/initialization.py:108-122)


def _p(dims, units, alias=None):
    d = {"dims": dims, "units": units}
    if alias:
        d["alias"] = alias
    return d


class _CorkBase(TendencyComponent):
    _which = "lw"

    def _setup(self, optics, table, kwargs, device, coefficients="solar_composition"):
        if optics not in ("parmentier", "correlated_k"):
            raise ValueError(f"Unknown optics mode: {optics}")
        self._optics_mode = optics
        self._diagnostics_level = kwargs.pop("diagnostics_level", 0)
        if optics == "parmentier":  # cork/lw/component.py:41-44, cork/sw/component.py:32-35
            self._coefficients = load_parmentier_coefficients(coefficients)
            self._freedman_coeffs = load_freedman2014_coefficients()
            self._num_bands = 2 if self._which == "lw" else 3
            self._has_co2_axis = False
            self._engine = CorkEngine(None, device=device, picket=(self._which, self._coefficients, self._freedman_coeffs))
            _num_bands[self._which] = self._num_bands
            return
        self._table = load_k_table(table) if isinstance(table, (str, os.PathLike)) else table
        if is_esft(self._table):
            self._table = expand_esft_table(self._table)
        k = self._table["k_coefficients"]
        self._num_bands, self._num_gpts = k.shape[1], k.shape[2]
        (self._gas_names, _has_h2o, self._has_co2_axis, self._fully_premixed, self._premixed_bg) = table_flags(self._table)
        self._engine = CorkEngine(self._table, device=device)
        _num_bands[self._which] = self._num_bands

    def _gas_props(self, props, with_co2):
        if self._optics_mode == "parmentier":  # cork/lw/component.py:96-106, cork/sw/component.py:115-125
            props["irradiation_temperature"] = _p(["*"], "degK", "T_irr")
            props["internal_temperature"] = _p(["*"], "degK", "T_int")
            return
        if self._premixed_bg:
            props["specific_humidity"] = _p(["mid_levels", "*"], "kg/kg", "h2o")
            if with_co2 and self._has_co2_axis:
                props["mole_fraction_of_carbon_dioxide_in_air"] = _p(["mid_levels", "*"], "mole/mole", "co2")
        elif not self._fully_premixed:
            for gas in self._gas_names:
                props[_GAS_CF_NAME.get(gas, f"mole_fraction_of_{gas}_in_air")] = _p(
                    ["mid_levels", "*"], "kg/kg" if gas == "h2o" else "mole/mole", gas)

    def _gas_arrays(self, state, nlev, arrays):
        """the gas part of array_call (cork/lw/component.py:243-287): what goes to the engine, in its units"""
        if self._optics_mode == "parmentier":
            arrays["T_irr"] = state["T_irr"].reshape(-1)
            arrays["T_int"] = state["T_int"].reshape(-1)
            return
        if self._fully_premixed:
            return
        if self._premixed_bg:
            arrays["q_h2o"] = state["h2o"].reshape(nlev, -1)
            if self._has_co2_axis and self._which == "lw":
                arrays["co2_vmr"] = state["co2"].reshape(nlev, -1)
            return
        gq = []
        for gas in self._gas_names:
            q = state[gas].reshape(nlev, -1)
            if gas != "h2o":
                q = q * (MOLAR_MASS.get(gas, MOLAR_MASS_DRY_AIR) / MOLAR_MASS_DRY_AIR)
            gq.append(q)
        arrays["gas_q"] = np.stack(gq)

    @staticmethod
    def _band_last(a, shape):
        # (nband, nlev, ncol) -> (nlev, *horizontal, nband); the reference returns the same values through moveaxis+reshape
        return np.moveaxis(a, 0, -1).reshape(shape + (a.shape[0],))


class CorkLongwaveRadiation(_CorkBase):
    """Drop-in for climt.CorkLongwaveRadiation (cork/lw/component.py:20-466)."""
    _which = "lw"

    def __init__(self, optics="parmentier", table=None, coefficients="solar_composition", rosseland_mean_fit="freedman2014",
                 diffusivity_factor=DIFFUSIVITY_FACTOR, device=0, **kwargs):
        self._diffusivity_factor = diffusivity_factor
        self._setup(optics, table, kwargs, device, coefficients)
        super().__init__(**kwargs)

    @property
    def input_properties(self):
        props = {
            "air_temperature": _p(["mid_levels", "*"], "degK", "T"),
            "air_pressure": _p(["mid_levels", "*"], "Pa", "p"),
            "air_pressure_on_interface_levels": _p(["interface_levels", "*"], "Pa", "p_int"),
            "surface_temperature": _p(["*"], "degK", "T_surf"),
            "surface_longwave_emissivity": _p(["num_longwave_bands", "*"], "dimensionless", "emissivity"),
        }
        self._gas_props(props, with_co2=True)
        props["longwave_optical_thickness_due_to_cloud"] = _p(["mid_levels", "*", "num_longwave_bands"], "dimensionless", "tau_cloud_lw")
        return props

    @property
    def tendency_properties(self):
        return {"air_temperature": {"units": "degK s^-1"}}

    @property
    def diagnostic_properties(self):
        band_i = ["interface_levels", "*", "num_longwave_bands"]
        band_m = ["mid_levels", "*", "num_longwave_bands"]
        return {
            "upwelling_longwave_flux_in_air": _p(["interface_levels", "*"], "W m^-2"),
            "downwelling_longwave_flux_in_air": _p(["interface_levels", "*"], "W m^-2"),
            "upwelling_longwave_flux_in_air_per_band": _p(band_i, "W m^-2"),
            "downwelling_longwave_flux_in_air_per_band": _p(band_i, "W m^-2"),
            "air_temperature_tendency_from_longwave": _p(["mid_levels", "*"], "degK day^-1"),
            "longwave_optical_depth_per_band": _p(band_m, "dimensionless"),
            "longwave_transmittance_per_band": _p(band_m, "dimensionless"),
            "air_temperature_tendency_from_longwave_per_band": _p(band_m, "degK day^-1"),
            **({"lw_layer_transmittance": _p(band_m, "dimensionless"), "lw_up_per_gpoint": _p(band_i, "W m^-2"),
                "lw_down_per_gpoint": _p(band_i, "W m^-2")} if self._diagnostics_level >= 1 else {}),  # cork/lw/component.py:189-202
        }

    @property
    def num_longwave_bands(self):
        return self._num_bands

    def array_call(self, state):
        T, p_int = _alias(state, "T", "air_temperature"), _alias(state, "p_int", "air_pressure_on_interface_levels")
        st = _AliasView(state, self.input_properties)
        shape_T, shape_pint = T.shape, p_int.shape
        nlev = T.shape[0]
        arrays = {"T": T.reshape(nlev, -1), "p": st["p"].reshape(nlev, -1), "p_int": p_int.reshape(nlev + 1, -1),
                  "T_surf": st["T_surf"].reshape(-1)}
        ncol = arrays["T"].shape[1]
        self._gas_arrays(st, nlev, arrays)
        arrays["emissivity"] = st["emissivity"].reshape(self._num_bands, ncol)
        arrays["tau_cloud"] = st["tau_cloud_lw"].reshape(nlev, ncol, self._num_bands)
        o = self._engine.lw_host(ncol, nlev, arrays, diffusivity_factor=self._diffusivity_factor,
                                 diagnostics_level=self._diagnostics_level)
        fields = {}
        if self._diagnostics_level:
            o, fields = o
        hr = o["heating_rate"].reshape(shape_T)
        diagnostics = {
            "upwelling_longwave_flux_in_air": o["up_broad"].reshape(shape_pint),
            "downwelling_longwave_flux_in_air": o["down_broad"].reshape(shape_pint),
            "upwelling_longwave_flux_in_air_per_band": self._band_last(o["up_band"], shape_pint),
            "downwelling_longwave_flux_in_air_per_band": self._band_last(o["down_band"], shape_pint),
            "air_temperature_tendency_from_longwave": hr * 86400.0,
            "longwave_optical_depth_per_band": self._band_last(o["tau_band"], shape_T),
            "longwave_transmittance_per_band": self._band_last(o["trans_band"], shape_T),
            "air_temperature_tendency_from_longwave_per_band": self._band_last(o["hr_band"], shape_T),
        }
        for name, a in fields.items():  # cork/lw/component.py:341-358
            diagnostics[name] = self._band_last(a, shape_pint if a.shape[1] == nlev + 1 else shape_T)
        return {_tend_key(state): hr}, diagnostics


class CorkShortwaveRadiation(_CorkBase):
    """Drop-in for climt.CorkShortwaveRadiation (cork/sw/component.py:20-532)."""
    _which = "sw"

    def __init__(self, optics="parmentier", table=None, coefficients="solar_composition", stellar_spectrum="sun",
                 rosseland_mean_fit="freedman2014", device=0, **kwargs):
        self._bond_albedo_feedback = kwargs.pop("bond_albedo_feedback", False)
        self._setup(optics, table, kwargs, device, coefficients)
        if optics == "parmentier":
            # fallback solar flux for un-irradiated columns: the stellar spectrum over three equal-width bands
            # (cork/sw/component.py:36-62)
            try:
                spec = load_stellar_spectrum(stellar_spectrum)
                wn = spec["wavenumber"]
                wn_lo, wn_hi = wn.min(), wn.max()
                bw = (wn_hi - wn_lo) / 3.0
                limits = np.array([[wn_lo, wn_lo + bw], [wn_lo + bw, wn_lo + 2 * bw], [wn_lo + 2 * bw, wn_hi]])
                self._solar_flux_per_band = integrate_spectrum_over_bands(spec, limits)
            except (FileNotFoundError, KeyError):
                self._solar_flux_per_band = np.array([1361.0 / 3.0] * 3)
        else:
            self._solar_source = self._table["solar_source_per_gpoint"]
            self._rayleigh = self._table.get("rayleigh_coefficient", None)
        super().__init__(**kwargs)

    @property
    def input_properties(self):
        props = {
            "air_temperature": _p(["mid_levels", "*"], "degK", "T"),
            "air_pressure": _p(["mid_levels", "*"], "Pa", "p"),
            "air_pressure_on_interface_levels": _p(["interface_levels", "*"], "Pa", "p_int"),
            "surface_temperature": _p(["*"], "degK", "T_surf"),
            "zenith_angle": _p(["*"], "radians", "zenith"),
            "surface_albedo_for_direct_shortwave": _p(["*"], "dimensionless", "albedo"),
            "flux_adjustment_for_earth_sun_distance": _p(["*"], "dimensionless", "earth_sun_factor"),
        }
        self._gas_props(props, with_co2=False)
        band_m = ["mid_levels", "*", "num_shortwave_bands"]
        props["shortwave_optical_thickness_due_to_cloud"] = _p(band_m, "dimensionless", "tau_cloud_sw")
        props["single_scattering_albedo_due_to_cloud"] = _p(band_m, "dimensionless", "ssa_cloud")
        props["cloud_asymmetry_parameter"] = _p(band_m, "dimensionless", "g_cloud")
        return props

    @property
    def tendency_properties(self):
        return {"air_temperature": {"units": "degK s^-1"}}

    @property
    def diagnostic_properties(self):
        band_i = ["interface_levels", "*", "num_shortwave_bands"]
        band_m = ["mid_levels", "*", "num_shortwave_bands"]
        return {
            "upwelling_shortwave_flux_in_air": _p(["interface_levels", "*"], "W m^-2"),
            "downwelling_shortwave_flux_in_air": _p(["interface_levels", "*"], "W m^-2"),
            "upwelling_shortwave_flux_in_air_per_band": _p(band_i, "W m^-2"),
            "downwelling_shortwave_flux_in_air_per_band": _p(band_i, "W m^-2"),
            "air_temperature_tendency_from_shortwave": _p(["mid_levels", "*"], "degK day^-1"),
            "shortwave_optical_depth_per_band": _p(band_m, "dimensionless"),
            "air_temperature_tendency_from_shortwave_per_band": _p(band_m, "degK day^-1"),
            # cork/sw/component.py:201-216
            **({name: _p(band_i if iface else band_m, "W m^-2" if name == "sw_direct_beam_profile" else "dimensionless")
                for name, iface, minlevel in SW_DIAG.values() if self._diagnostics_level >= minlevel}),
        }

    @property
    def num_shortwave_bands(self):
        return self._num_bands

    def _parmentier_call(self, ncol, nlev, arrays, earth_sun_factor):
        """The picket-fence branch of CorkShortwaveRadiation.array_call (cork/sw/component.py:242-306): stellar flux from
        the hottest irradiation temperature, one engine pass, and with bond_albedo_feedback a second pass whose T_eff uses
        the Bond albedo of the first (bond_albedo_from_fluxes, cork/optics/parmentier.py:156-163)."""
        sigma = self._engine.sigma
        T_irr_max = arrays["T_irr"].max()
        if T_irr_max > 0:
            F0 = sigma * T_irr_max**4
            per_band = np.array([F0 / 3.0] * 3)
        else:
            per_band = self._solar_flux_per_band
        solar_flux = per_band.reshape(3, 1) * np.ones((3, 1)) * earth_sun_factor
        o = None
        for it in range(2 if self._bond_albedo_feedback else 1):
            if it:
                oo = o[0] if isinstance(o, tuple) else o
                up_toa, down_toa = oo["up_broad"][-1, :], oo["down_broad"][-1, :]
                with np.errstate(divide="ignore", invalid="ignore"):
                    arrays["bond_albedo"] = np.clip(np.where(down_toa > 0, up_toa / down_toa, 0.0), 0.0, 1.0)
            o = self._engine.sw_host(ncol, nlev, arrays, solar_flux=solar_flux, out=o[0] if isinstance(o, tuple) else o,
                                     diagnostics_level=self._diagnostics_level)
        return o

    def array_call(self, state):
        T, p_int = _alias(state, "T", "air_temperature"), _alias(state, "p_int", "air_pressure_on_interface_levels")
        st = _AliasView(state, self.input_properties)
        shape_T, shape_pint = T.shape, p_int.shape
        nlev = T.shape[0]
        arrays = {"T": T.reshape(nlev, -1), "p": st["p"].reshape(nlev, -1), "p_int": p_int.reshape(nlev + 1, -1),
                  "zenith": st["zenith"].reshape(-1), "albedo": st["albedo"].reshape(-1)}
        ncol = arrays["T"].shape[1]
        self._gas_arrays(st, nlev, arrays)
        nb = self._num_bands
        arrays["tau_cloud"] = st["tau_cloud_sw"].reshape(nlev, ncol, nb)
        arrays["ssa_cloud"] = st["ssa_cloud"].reshape(nlev, ncol, nb)
        arrays["g_cloud"] = st["g_cloud"].reshape(nlev, ncol, nb)
        esf = float(np.asarray(st["earth_sun_factor"]).reshape(-1)[0])  # cork/sw/component.py:370
        if self._optics_mode == "parmentier":
            o = self._parmentier_call(ncol, nlev, arrays, esf)
        else:
            o = self._engine.sw_host(ncol, nlev, arrays, earth_sun_factor=esf, diagnostics_level=self._diagnostics_level)
        fields = {}
        if self._diagnostics_level:
            o, fields = o
        hr = o["heating_rate"].reshape(shape_T)
        diagnostics = {
            "upwelling_shortwave_flux_in_air": o["up_broad"].reshape(shape_pint),
            "downwelling_shortwave_flux_in_air": o["down_broad"].reshape(shape_pint),
            "upwelling_shortwave_flux_in_air_per_band": self._band_last(o["up_band"], shape_pint),
            "downwelling_shortwave_flux_in_air_per_band": self._band_last(o["down_band"], shape_pint),
            "air_temperature_tendency_from_shortwave": hr * 86400.0,
            "shortwave_optical_depth_per_band": self._band_last(o["tau_band"], shape_T),
            "air_temperature_tendency_from_shortwave_per_band": self._band_last(o["hr_band"], shape_T),
        }
        for name, a in fields.items():  # cork/sw/component.py:455-492
            diagnostics[name] = self._band_last(a, shape_pint if a.shape[1] == nlev + 1 else shape_T)
        return {_tend_key(state): hr}, diagnostics


# sympl hands array_call a dict keyed by alias when aliases are declared (the reference indexes state["T"]); the local shim
# keys by quantity name.  Accept both.
def _alias(state, alias, name):
    return state[alias] if alias in state else state[name]


def _tend_key(state):
    return "T" if "T" in state else "air_temperature"


class _AliasView:
    def __init__(self, state, props):
        self._s = state
        self._by_alias = {v["alias"]: k for k, v in props.items() if "alias" in v}

    def __getitem__(self, alias):
        return self._s[alias] if alias in self._s else self._s[self._by_alias[alias]]
