"""Device-resident model state: when the arrays handed to a component's ``array_call`` are torch CUDA tensors
the component stays on the GPU -- the numpy marshal arithmetic of the reference's array_call (specific humidity -> volume mixing
ratio, interface temperatures, cos(zenith)) runs in ``cb200_marshal_device``, the engines take the tensors' device pointers
(zero copy, asynchronous on the current stream) and the outputs are torch CUDA tensors.  No PCIe traffic at all."""
import ctypes

from . import _native

_vp = ctypes.c_void_p


def is_device_state(state):
    """True when the state's arrays are torch CUDA tensors (checked on the temperature field)."""
    t = state.get("air_temperature")
    return t is not None and type(t).__module__.startswith("torch") and getattr(t, "is_cuda", False)


def dense(t):
    """contiguous float64 CUDA tensor (no copy when it already is one)"""
    import torch
    return t.to(dtype=torch.float64).contiguous()


def _ptr(t):
    return _vp(t.data_ptr()) if t is not None else None


def marshal(q, t, tsfc, p, p_int, zenith=None, want_tlev=True):
    """-> (h2ovmr, tlev or None, coszen or None), asynchronous on the current stream."""
    import torch
    L = _native.lib()
    L.cb200_marshal_device.argtypes = [ctypes.c_int] * 3 + [_vp] * 10
    nlay, ncol = t.shape
    h2o = torch.empty_like(t)
    tlev = torch.empty((nlay + 1, ncol), dtype=torch.float64, device=t.device) if want_tlev else None
    cz = torch.empty_like(zenith) if zenith is not None else None
    rc = L.cb200_marshal_device(t.device.index or 0, ncol, nlay, _ptr(q), _ptr(t), _ptr(tsfc), _ptr(p), _ptr(p_int), _ptr(zenith),
                                _ptr(h2o), _ptr(tlev), _ptr(cz), _vp(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise RuntimeError(L.cb200_global_error().decode())
    return h2o, tlev, cz


def host_marshal_on_device():
    """Host-array calls of the RRTMG components: let the engine do the q -> vmr conversion and the ln-p interface temperatures on
    the device (cb200_*_set_host_marshal; default) or evaluate the reference's numpy expressions on the host
    (CLIMT_B200_HOST_MARSHAL=numpy).  The two agree to the last bits of log()."""
    import os
    return os.environ.get("CLIMT_B200_HOST_MARSHAL", "device").lower() != "numpy"


def finish(engine, asynchronous):
    """Synchronous mode (default): wait for the call and raise the input-validation errors the Fortran turns into `stop`;
    asynchronous mode: return at once, the caller synchronises and calls ``engine.check()`` when it wants to."""
    if not asynchronous:
        import torch
        torch.cuda.current_stream().synchronize()
        engine.check()


def keep_alive_on(stream, tensors):
    """A call was enqueued on an explicit raw `stream` handle: tensors allocated inside that call (on torch's current stream)
    must not be handed back to the caching allocator for reuse before that stream has consumed them -> record_stream."""
    if stream is None:
        return
    import torch
    s = torch.cuda.ExternalStream(int(stream))
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            t.record_stream(s)
