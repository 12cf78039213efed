"""Loader / builder of the native CUDA engine (libclimt_b200.so, C ABI in include/climt_b200.h).

The product path has no CPU fallback: if the shared library (or a CUDA device) is missing the
components raise.  `build()` is what __graft_entry__.build() calls; nvcc cross-compiles sm_100a
without a GPU.
"""
import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CLIMT_B200_SO") or os.path.join(_HERE, "libclimt_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("lw_engine.cu", "sw_engine.cu", "gray_engine.cu", "cork_engine.cu", "marshal.cu", "emanuel_engine.cu", "adjacent_engine.cu", "simple_physics.cu")]
HEADERS = [os.path.join(_HERE, "csrc", f) for f in ("cb_common.h", "engine_common.h", "lw_core.cuh", "lw_tables.h", "sw_core.cuh", "sw_tables.h",
                                                          "mcica_core.cuh", "mcica_host.h", "mcica_compat.h", "cork_core.cuh", "cork_tables.h", "emanuel_core.cuh", "simple_physics_core.cuh")] + [
    os.path.join(_HERE, "..", "include", "climt_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared", "--fmad=true"]

_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


class NativeUnavailable(RuntimeError):
    pass


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise NativeUnavailable("nvcc not found; cannot build libclimt_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", SO_PATH + ".tmp"] + SOURCES + ["-lcudart", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    os.replace(SO_PATH + ".tmp", SO_PATH)
    return SO_PATH


class LwInputs(ctypes.Structure):
    _fields_ = [(n, _dp) for n in (
        "play", "plev", "tlay", "tlev", "tsfc", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
        "cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr", "emis", "cldfr", "taucld", "cicewp", "cliqwp",
        "reice", "reliq", "tauaer")]


class LwOutputs(ctypes.Structure):
    _fields_ = [(n, _dp) for n in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")]


class SwInputs(ctypes.Structure):
    _fields_ = [(n, _dp) for n in (
        "play", "plev", "tlay", "tlev", "tsfc", "h2ovmr", "o3vmr", "co2vmr", "ch4vmr", "n2ovmr", "o2vmr",
        "asdir", "asdif", "aldir", "aldif", "coszen", "cldfr", "taucld", "ssacld", "asmcld", "fsfcld",
        "cicewp", "cliqwp", "reice", "reliq", "tauaer", "ssaaer", "asmaer", "ecaer")]


EXPORTS = ["cb200_lw_zero_scan_state", "mcica_subcol_lw_wrapper", "rrtmg_lw_mcica_wrapper", "mcica_subcol_sw_wrapper", "rrtmg_sw_mcica_wrapper",
           "cb200_lw_set_derivative_outputs", "cb200_berger_scalars", "cb200_berger_run_device", "cb200_berger_run_host", "cb200_simple_physics_run_device", "cb200_simple_physics_run_host", "set_fortran_constants", "simple_physics",
           "cb200_cork_create_from_file", "cb200_instellation_orbit", "cb200_instellation_run_device", "cb200_instellation_run_host", "cb200_slab_surface_run_device",
           "cb200_slab_surface_run_host", "cb200_emanuel_create", "cb200_emanuel_destroy", "cb200_emanuel_last_error", "cb200_emanuel_last_launches", "cb200_emanuel_enable_timing",
           "cb200_emanuel_last_kernel_ms", "cb200_emanuel_run_device", "cb200_emanuel_run_host", "init_emanuel_convection_fortran", "emanuel_convection",
           "cb200_cork_create_picket", "cb200_marshal_device", "cb200_lw_last_taumol_kernel_ms", "cb200_sw_last_taumol_kernel_ms", "cb200_lw_run_host_async", "cb200_lw_wait", "cb200_lw_last_transfer_bytes", "cb200_sw_run_host_async", "cb200_sw_wait",
           "cb200_sw_last_transfer_bytes", "cb200_cork_create", "cb200_cork_destroy", "cb200_cork_last_error", "cb200_cork_last_launches", "cb200_cork_enable_timing",
           "cb200_cork_last_unit_kernel_ms", "cb200_cork_lw_run_device", "cb200_cork_sw_run_device", "cb200_cork_lw_run_host",
           "cb200_cork_sw_run_host", "cb200_cork_set_diagnostics", "cb200_lw_set_host_marshal", "cb200_sw_set_host_marshal",
           "cb200_gray_lw_run_device", "cb200_gray_lw_run_host", "cb200_sw_create", "cb200_sw_destroy", "cb200_sw_set_options", "cb200_sw_set_mcica", "cb200_sw_set_solar", "cb200_sw_run_device",
           "cb200_sw_run_host", "cb200_sw_check", "cb200_sw_last_error", "cb200_sw_last_launches",
           "cb200_sw_enable_timing", "cb200_sw_last_unit_kernel_ms", "rrtmg_sw_set_constants", "rrtmg_sw_ini_wrapper",
           "rrtmg_sw_nomcica_wrapper",
           "cb200_lw_create", "cb200_lw_destroy", "cb200_lw_set_options", "cb200_lw_set_mcica", "cb200_lw_run_device",
           "cb200_lw_run_host", "cb200_lw_check", "cb200_lw_last_error", "cb200_global_error",
           "cb200_lw_last_launches", "cb200_lw_enable_timing", "cb200_lw_last_unit_kernel_ms",
           "rrtmg_set_constants", "rrtmg_lw_ini_wrapper", "rrtmg_lw_nomcica_wrapper"]


def lib():
    """ctypes handle of libclimt_b200.so; raises NativeUnavailable (never falls back) if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise NativeUnavailable(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA engine is the only implementation; there is no CPU fallback)")
    L = ctypes.CDLL(SO_PATH)
    vp = ctypes.c_void_p
    L.cb200_lw_create.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, _dp, ctypes.c_int]
    L.cb200_lw_create.restype = ctypes.c_int
    L.cb200_lw_destroy.argtypes = [vp]
    L.cb200_lw_destroy.restype = None
    L.cb200_lw_set_options.argtypes = [vp] + [ctypes.c_int] * 5
    L.cb200_lw_set_mcica.argtypes = [vp] + [ctypes.c_int] * 3
    L.cb200_lw_set_derivative_outputs.argtypes = [vp, _dp, _dp]
    L.cb200_lw_zero_scan_state.argtypes = [vp]
    L.cb200_lw_run_device.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(LwInputs),
                                      ctypes.POINTER(LwOutputs), vp]
    L.cb200_lw_run_host.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(LwInputs),
                                    ctypes.POINTER(LwOutputs)]
    L.cb200_lw_run_host_async.argtypes = L.cb200_lw_run_host.argtypes
    L.cb200_lw_wait.argtypes = [vp]
    L.cb200_lw_last_transfer_bytes.argtypes = [vp, _dp, _dp]
    L.cb200_lw_last_transfer_bytes.restype = None
    L.cb200_lw_check.argtypes = [vp]
    L.cb200_lw_last_error.argtypes = [vp]
    L.cb200_lw_last_error.restype = ctypes.c_char_p
    L.cb200_global_error.restype = ctypes.c_char_p
    L.cb200_lw_last_launches.argtypes = [vp]
    L.cb200_lw_enable_timing.argtypes = [vp, ctypes.c_int]
    L.cb200_lw_last_unit_kernel_ms.argtypes = [vp]
    L.cb200_lw_last_unit_kernel_ms.restype = ctypes.c_double
    L.cb200_lw_last_taumol_kernel_ms.argtypes = [vp]
    L.cb200_lw_last_taumol_kernel_ms.restype = ctypes.c_double
    L.cb200_sw_last_taumol_kernel_ms.argtypes = [vp]
    L.cb200_sw_last_taumol_kernel_ms.restype = ctypes.c_double
    L.cb200_sw_create.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, _dp, ctypes.c_int]
    L.cb200_sw_destroy.argtypes = [vp]
    L.cb200_sw_destroy.restype = None
    L.cb200_sw_set_options.argtypes = [vp] + [ctypes.c_int] * 5
    L.cb200_sw_set_mcica.argtypes = [vp] + [ctypes.c_int] * 3
    L.cb200_sw_set_solar.argtypes = [vp, ctypes.c_int, ctypes.c_double, _dp, _dp]
    L.cb200_sw_run_device.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                      ctypes.POINTER(SwInputs), ctypes.POINTER(LwOutputs), vp]
    L.cb200_sw_run_host.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                    ctypes.POINTER(SwInputs), ctypes.POINTER(LwOutputs)]
    L.cb200_sw_run_host_async.argtypes = L.cb200_sw_run_host.argtypes
    L.cb200_sw_wait.argtypes = [vp]
    L.cb200_sw_last_transfer_bytes.argtypes = [vp, _dp, _dp]
    L.cb200_sw_last_transfer_bytes.restype = None
    L.cb200_sw_check.argtypes = [vp]
    L.cb200_sw_last_error.argtypes = [vp]
    L.cb200_sw_last_error.restype = ctypes.c_char_p
    L.cb200_sw_last_launches.argtypes = [vp]
    L.cb200_sw_enable_timing.argtypes = [vp, ctypes.c_int]
    L.cb200_sw_last_unit_kernel_ms.argtypes = [vp]
    L.cb200_sw_last_unit_kernel_ms.restype = ctypes.c_double
    _lib = L
    return L
