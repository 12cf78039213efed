"""Python face of the native engines: thin, typed wrappers over the C ABI (include/climt_b200.h).

`LWEngine.run_host` takes numpy arrays (what sympl hands to `array_call`), `LWEngine.run_device`
takes torch CUDA tensors and is asynchronous on the current torch stream.
"""
import ctypes

import numpy as np

from . import _native
from .constants import rrtmg_constants
from .rrtmg_tables import lw_blob_path, sw_blob_path

_dp = ctypes.POINTER(ctypes.c_double)
LW_IN = [f[0] for f in _native.LwInputs._fields_]
LW_OUT = [f[0] for f in _native.LwOutputs._fields_]
_CONST_ORDER = ("pi", "grav", "planck", "boltz", "clight", "avogad", "alosmt", "gascon", "sbcnst", "secdy", "cpdair")


def lw_shapes(ncol, nlay):
    L, n = nlay, ncol
    ins = {k: (L, n) for k in LW_IN}
    ins.update(plev=(L + 1, n), tlev=(L + 1, n), tsfc=(n,), emis=(16, n), taucld=(L, n, 16), tauaer=(16, L, n))
    outs = {"uflx": (L + 1, n), "dflx": (L + 1, n), "uflxc": (L + 1, n), "dflxc": (L + 1, n), "hr": (L, n), "hrc": (L, n)}
    return ins, outs


class LWEngine:
    """One RRTMG-LW engine instance: tables resident in HBM, options and constants per instance
    (the reference keeps them in Cython/Fortran module globals, _rrtmg_lw.pyx:8-14)."""

    def __init__(self, constants=None, device=0, icld=1, idrv=0, inflag=2, iceflag=1, liqflag=1, mcica=False, irng=1,
                 permuteseed=0):
        self._L = _native.lib()
        k = constants or rrtmg_constants()
        c = np.array([k[n] for n in _CONST_ORDER], dtype=np.float64)
        h = ctypes.c_void_p()
        rc = self._L.cb200_lw_create(ctypes.byref(h), lw_blob_path().encode(), c.ctypes.data_as(_dp), int(device))
        if rc != 0 or not h:
            raise RuntimeError("cb200_lw_create failed: " + self._L.cb200_global_error().decode())
        self._h = h
        self.device = device
        self.set_mcica(mcica, irng, permuteseed)
        self.set_options(icld, idrv, inflag, iceflag, liqflag)

    def set_host_marshal(self, from_specific_humidity=False, compute_tlev=False):
        """run_host only: the engine converts specific humidity (handed over as `h2ovmr`) and / or computes `tlev` (may then be
        None) on the device, chunk by chunk -- the components' numpy marshal arithmetic (cb200_*_set_host_marshal)."""
        self._host_marshal = (1 if from_specific_humidity else 0) | (2 if compute_tlev else 0)
        fn = getattr(self._L, "cb200_lw_set_host_marshal" if isinstance(self, LWEngine) else "cb200_sw_set_host_marshal")
        fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
        fn(self._h, self._host_marshal)

    def set_mcica(self, enabled, irng=1, permuteseed=0):
        """McICA on/off, RNG (0 kissvec on the device, 1 Mersenne twister on the host for bit parity), seed."""
        self._L.cb200_lw_set_mcica(self._h, 1 if enabled else 0, int(irng), int(permuteseed))

    def set_options(self, icld, idrv, inflag, iceflag, liqflag):
        if self._L.cb200_lw_set_options(self._h, icld, idrv, inflag, iceflag, liqflag):
            raise ValueError(self._err())
        self.idrv = int(idrv)

    def _err(self):
        return self._L.cb200_lw_last_error(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            self._L.cb200_lw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host buffers (numpy) ------------------------------------------------------------------
    def run_host(self, ncol, nlay, arrays, out=None, wait=True):
        """wait=False: return as soon as the call is enqueued (overlap with other engines); finish with .wait()."""
        ins, outs = lw_shapes(ncol, nlay)
        keep = []
        pin = _native.LwInputs()
        for k in LW_IN:
            if k == "tlev" and getattr(self, "_host_marshal", 0) & 2:
                continue  # computed by the engine on the device (set_host_marshal); the pointer stays NULL
            a = np.ascontiguousarray(arrays[k], dtype=np.float64)
            if a.shape != ins[k]:
                raise ValueError(f"{k}: expected shape {ins[k]}, got {a.shape}")
            keep.append(a)
            setattr(pin, k, a.ctypes.data_as(_dp))
        out = out if out is not None else {k: np.empty(outs[k]) for k in LW_OUT}
        pout = _native.LwOutputs()
        for k in LW_OUT:
            a = out[k]
            if a.shape != outs[k] or a.dtype != np.float64 or not a.flags.c_contiguous:
                raise ValueError(f"output {k}: need C-contiguous float64 {outs[k]}")
            setattr(pout, k, a.ctypes.data_as(_dp))
        if self.idrv == 1:  # calculate_change_up_flux: duflx_dt / duflxc_dt of rrtmg_lw_c_binder.f90:176-256
            for k in ("duflx_dt", "duflxc_dt"):
                a = out.setdefault(k, np.empty(outs["uflx"]))
                if a.shape != outs["uflx"] or a.dtype != np.float64 or not a.flags.c_contiguous:
                    raise ValueError(f"output {k}: need C-contiguous float64 {outs['uflx']}")
            self._L.cb200_lw_set_derivative_outputs(self._h, out["duflx_dt"].ctypes.data_as(_dp), out["duflxc_dt"].ctypes.data_as(_dp))
        fn = self._L.cb200_lw_run_host_async if wait is False else self._L.cb200_lw_run_host
        rc = fn(self._h, ncol, nlay, ctypes.byref(pin), ctypes.byref(pout))
        if wait is False and rc == 0:
            self._pending = (keep, out)  # keep the (possibly converted) host buffers alive until wait()
        if rc == -3:
            raise ValueError(self._err())
        if rc < 0:
            raise RuntimeError(self._err())
        if rc > 0:
            raise ValueError(self._err())
        return out

    def wait(self):
        """Complete a run_host(..., wait=False) call: the outputs are filled on return."""
        rc = self._L.cb200_lw_wait(self._h)
        pend, self._pending = getattr(self, "_pending", None), None
        if rc == -3:
            raise ValueError(self._err())
        if rc < 0:
            raise RuntimeError(self._err())
        if rc > 0:
            raise ValueError(self._err())
        return pend[1] if pend else None

    @property
    def last_transfer_bytes(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        self._L.cb200_lw_last_transfer_bytes(self._h, ctypes.byref(a), ctypes.byref(b))
        return int(a.value), int(b.value)

    @property
    def zero_scan_state(self):
        """1 scanning, 0 disabled by the environment, -1 disabled by the engine's own guard (see include/climt_b200.h)"""
        return int(self._L.cb200_lw_zero_scan_state(self._h))

    # -- device buffers (torch CUDA tensors), asynchronous ------------------------------------------
    def run_device(self, ncol, nlay, tensors, out, stream=None):
        import torch
        ins, outs = lw_shapes(ncol, nlay)
        pin = _native.LwInputs()
        for k in LW_IN:
            t = tensors[k]
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == ins[k]):
                raise ValueError(f"{k}: need contiguous float64 CUDA tensor of shape {ins[k]}")
            setattr(pin, k, ctypes.cast(t.data_ptr(), _dp))
        pout = _native.LwOutputs()
        for k in LW_OUT:
            t = out[k]
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == outs[k]):
                raise ValueError(f"output {k}: need contiguous float64 CUDA tensor of shape {outs[k]}")
            setattr(pout, k, ctypes.cast(t.data_ptr(), _dp))
        if self.idrv == 1:
            for k in ("duflx_dt", "duflxc_dt"):
                t = out[k]
                if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == outs["uflx"]):
                    raise ValueError(f"output {k}: need contiguous float64 CUDA tensor of shape {outs['uflx']}")
            self._L.cb200_lw_set_derivative_outputs(self._h, ctypes.cast(out["duflx_dt"].data_ptr(), _dp),
                                                    ctypes.cast(out["duflxc_dt"].data_ptr(), _dp))
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        rc = self._L.cb200_lw_run_device(self._h, ncol, nlay, ctypes.byref(pin), ctypes.byref(pout), ctypes.c_void_p(s))
        if rc:
            raise RuntimeError(self._err())

    def check(self):
        rc = self._L.cb200_lw_check(self._h)
        if rc:
            raise ValueError(self._err())

    def enable_timing(self, on=True):
        self._L.cb200_lw_enable_timing(self._h, 1 if on else 0)

    @property
    def last_unit_kernel_ms(self):
        return self._L.cb200_lw_last_unit_kernel_ms(self._h)

    @property
    def last_taumol_kernel_ms(self):
        return self._L.cb200_lw_last_taumol_kernel_ms(self._h)

    @property
    def last_launches(self):
        return self._L.cb200_lw_last_launches(self._h)


SW_IN = [f[0] for f in _native.SwInputs._fields_]


def sw_shapes(ncol, nlay):
    L, n = nlay, ncol
    ins = {k: (L, n) for k in SW_IN}
    ins.update(plev=(L + 1, n), tlev=(L + 1, n), tsfc=(n,), asdir=(n,), asdif=(n,), aldir=(n,), aldif=(n,), coszen=(n,),
               taucld=(L, n, 14), ssacld=(L, n, 14), asmcld=(L, n, 14), fsfcld=(L, n, 14),
               tauaer=(14, L, n), ssaaer=(14, L, n), asmaer=(14, L, n), ecaer=(6, L, n))
    outs = {"uflx": (L + 1, n), "dflx": (L + 1, n), "uflxc": (L + 1, n), "dflxc": (L + 1, n), "hr": (L, n), "hrc": (L, n)}
    return ins, outs


class SWEngine:
    """One RRTMG-SW engine instance (non-McICA driver)."""

    def __init__(self, constants=None, device=0, icld=1, iaer=0, inflag=2, iceflag=1, liqflag=1, isolvar=0,
                 scon=1367.0, indsolvar=(1.0, 1.0), bndsolvar=None, mcica=False, irng=1, permuteseed=0):
        self._L = _native.lib()
        k = constants or rrtmg_constants()
        c = np.array([k[n] for n in _CONST_ORDER], dtype=np.float64)
        h = ctypes.c_void_p()
        rc = self._L.cb200_sw_create(ctypes.byref(h), sw_blob_path().encode(), c.ctypes.data_as(_dp), int(device))
        if rc != 0 or not h:
            raise RuntimeError("cb200_sw_create failed: " + self._L.cb200_global_error().decode())
        self._h = h
        self.device = device
        self.set_mcica(mcica, irng, permuteseed)
        self._L.cb200_sw_set_options(self._h, icld, iaer, inflag, iceflag, liqflag)
        ind = np.array(indsolvar, dtype=np.float64)
        bnd = np.ones(14) if bndsolvar is None else np.ascontiguousarray(np.asarray(bndsolvar, dtype=np.float64)[:14])
        self._L.cb200_sw_set_solar(self._h, int(isolvar), float(scon), ind.ctypes.data_as(_dp), bnd.ctypes.data_as(_dp))

    def set_host_marshal(self, from_specific_humidity=False, compute_tlev=False):
        """run_host only: the engine converts specific humidity (handed over as `h2ovmr`) and / or computes `tlev` (may then be
        None) on the device, chunk by chunk -- the components' numpy marshal arithmetic (cb200_*_set_host_marshal)."""
        self._host_marshal = (1 if from_specific_humidity else 0) | (2 if compute_tlev else 0)
        fn = getattr(self._L, "cb200_lw_set_host_marshal" if isinstance(self, LWEngine) else "cb200_sw_set_host_marshal")
        fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
        fn(self._h, self._host_marshal)

    def set_mcica(self, enabled, irng=1, permuteseed=0):
        self._L.cb200_sw_set_mcica(self._h, 1 if enabled else 0, int(irng), int(permuteseed))

    def _err(self):
        return self._L.cb200_sw_last_error(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            self._L.cb200_sw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_host(self, ncol, nlay, arrays, out=None, adjes=1.0, dyofyr=0, solcycfrac=0.0, wait=True):
        ins, outs = sw_shapes(ncol, nlay)
        keep = []
        pin = _native.SwInputs()
        for k in SW_IN:
            if k == "tlev" and getattr(self, "_host_marshal", 0) & 2:
                continue  # computed by the engine on the device (set_host_marshal); the pointer stays NULL
            a = np.ascontiguousarray(arrays[k], dtype=np.float64)
            if a.shape != ins[k]:
                raise ValueError(f"{k}: expected shape {ins[k]}, got {a.shape}")
            keep.append(a)
            setattr(pin, k, a.ctypes.data_as(_dp))
        out = out if out is not None else {k: np.empty(outs[k]) for k in LW_OUT}
        pout = _native.LwOutputs()
        for k in LW_OUT:
            a = out[k]
            if a.shape != outs[k] or a.dtype != np.float64 or not a.flags.c_contiguous:
                raise ValueError(f"output {k}: need C-contiguous float64 {outs[k]}")
            setattr(pout, k, a.ctypes.data_as(_dp))
        fn = self._L.cb200_sw_run_host_async if wait is False else self._L.cb200_sw_run_host
        rc = fn(self._h, ncol, nlay, float(adjes), int(dyofyr), float(solcycfrac), ctypes.byref(pin), ctypes.byref(pout))
        if wait is False and rc == 0:
            self._pending = (keep, out)
        if rc == -3:
            raise ValueError(self._err())
        if rc < 0:
            raise RuntimeError(self._err())
        if rc > 0:
            raise ValueError(self._err())
        return out

    def wait(self):
        """Complete a run_host(..., wait=False) call: the outputs are filled on return."""
        rc = self._L.cb200_sw_wait(self._h)
        pend, self._pending = getattr(self, "_pending", None), None
        if rc == -3:
            raise ValueError(self._err())
        if rc < 0:
            raise RuntimeError(self._err())
        if rc > 0:
            raise ValueError(self._err())
        return pend[1] if pend else None

    @property
    def last_transfer_bytes(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        self._L.cb200_sw_last_transfer_bytes(self._h, ctypes.byref(a), ctypes.byref(b))
        return int(a.value), int(b.value)

    def run_device(self, ncol, nlay, tensors, out, adjes=1.0, dyofyr=0, solcycfrac=0.0, stream=None):
        import torch
        ins, outs = sw_shapes(ncol, nlay)
        pin = _native.SwInputs()
        for k in SW_IN:
            t = tensors[k]
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == ins[k]):
                raise ValueError(f"{k}: need contiguous float64 CUDA tensor of shape {ins[k]}")
            setattr(pin, k, ctypes.cast(t.data_ptr(), _dp))
        pout = _native.LwOutputs()
        for k in LW_OUT:
            t = out[k]
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == outs[k]):
                raise ValueError(f"output {k}: need contiguous float64 CUDA tensor of shape {outs[k]}")
            setattr(pout, k, ctypes.cast(t.data_ptr(), _dp))
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        rc = self._L.cb200_sw_run_device(self._h, ncol, nlay, float(adjes), int(dyofyr), float(solcycfrac),
                                         ctypes.byref(pin), ctypes.byref(pout), ctypes.c_void_p(s))
        if rc:
            raise RuntimeError(self._err())

    def check(self):
        rc = self._L.cb200_sw_check(self._h)
        if rc:
            raise ValueError(self._err())

    def enable_timing(self, on=True):
        self._L.cb200_sw_enable_timing(self._h, 1 if on else 0)

    @property
    def last_unit_kernel_ms(self):
        return self._L.cb200_sw_last_unit_kernel_ms(self._h)

    @property
    def last_taumol_kernel_ms(self):
        return self._L.cb200_sw_last_taumol_kernel_ms(self._h)

    @property
    def last_launches(self):
        return self._L.cb200_sw_last_launches(self._h)
