"""The host-pointer call streams column chunks through a 3-stream pipeline and skips arrays the flags make dead.
It must give bit-identical fluxes to one device-pointer call over the whole grid, for chunk counts that exercise
slot reuse (>= 3 chunks) and a ragged last chunk."""
import numpy as np
import pytest

import helpers as H
from climt_b200 import synthetic as SY

pytestmark = pytest.mark.gpu


def _device_lw(eng, ncol, nlay, abi):
    import torch
    from climt_b200.engine import LW_IN, lw_shapes
    dev_in = {k: torch.from_numpy(abi[k]).cuda() for k in LW_IN}
    _, outs = lw_shapes(ncol, nlay)
    dev_out = {k: torch.empty(s, dtype=torch.float64, device="cuda") for k, s in outs.items()}
    eng.run_device(ncol, nlay, dev_in, dev_out)
    torch.cuda.synchronize()
    eng.check()
    return {k: v.cpu().numpy() for k, v in dev_out.items()}


def _device_sw(eng, ncol, nlay, abi, **kw):
    import torch
    from climt_b200.engine import SW_IN, lw_shapes
    dev_in = {k: torch.from_numpy(abi[k]).cuda() for k in SW_IN}
    _, outs = lw_shapes(ncol, nlay)
    dev_out = {k: torch.empty(s, dtype=torch.float64, device="cuda") for k, s in outs.items()}
    eng.run_device(ncol, nlay, dev_in, dev_out, **kw)
    torch.cuda.synchronize()
    eng.check()
    return {k: v.cpu().numpy() for k, v in dev_out.items()}


@pytest.mark.parametrize("kw", [dict(), dict(icld=0), dict(inflag=0), dict(icld=1, mcica=True, irng=1, permuteseed=7),
                                dict(icld=2, mcica=True, irng=0, permuteseed=7)])
def test_lw_chunked_host_call_equals_device_call(monkeypatch, kw):
    from climt_b200.engine import LWEngine
    monkeypatch.setenv("CLIMT_B200_HOST_CHUNK", "256")
    ncol, nlay = 1000, 40
    st = SY.make_lw_state(ncol, nlay, seed=17, clouds=True, aerosol=True)
    if kw.get("inflag") == 0:
        st["taucld"][:] = np.random.default_rng(1).uniform(0.0, 2.0, st["taucld"].shape) * (st["cldfr"] > 0)[..., None]
    abi = H.to_abi(st)
    eng = LWEngine(**kw)
    host = eng.run_host(ncol, nlay, abi)
    dev = _device_lw(eng, ncol, nlay, abi)
    eng.close()
    for k in dev:
        np.testing.assert_array_equal(host[k], dev[k], err_msg=k)
    assert np.isfinite(host["uflx"]).all() and host["uflx"].min() > 0


@pytest.mark.parametrize("kw,mode", [(dict(), "clouds"), (dict(icld=0), "clear"), (dict(iaer=10), "aerosol"), (dict(iaer=6), "ecmwf"),
                                     (dict(icld=1, mcica=True, irng=1, permuteseed=7), "mcica"),
                                     (dict(icld=3, mcica=True, irng=0, permuteseed=7), "mcica")])
def test_sw_chunked_host_call_equals_device_call(monkeypatch, kw, mode):
    from climt_b200.engine import SWEngine
    monkeypatch.setenv("CLIMT_B200_HOST_CHUNK", "256")
    ncol, nlay = 1000, 40
    st = SY.make_sw_state(ncol, nlay, seed=23, clouds=mode in ("clouds", "mcica"), aerosol=mode == "aerosol",
                          ecmwf=mode == "ecmwf", **({"overcast_only": False} if mode == "mcica" else {}))
    abi = H.to_abi_sw(st)
    eng = SWEngine(**kw)
    host = eng.run_host(ncol, nlay, abi, dyofyr=100)
    dev = _device_sw(eng, ncol, nlay, abi, dyofyr=100)
    eng.close()
    for k in dev:
        np.testing.assert_array_equal(host[k], dev[k], err_msg=k)
    assert np.isfinite(host["dflx"]).all()


def test_async_host_calls_overlap_and_match_sync(monkeypatch):
    """run_host(wait=False) on the LW and SW engines, then wait(): same bits as the synchronous calls; transfer-byte
    accounting reflects the skipped arrays."""
    from climt_b200.engine import LWEngine, SWEngine
    monkeypatch.setenv("CLIMT_B200_HOST_CHUNK", "512")
    ncol, nlay = 1500, 30
    abi = H.to_abi(SY.make_lw_state(ncol, nlay, seed=3, clouds=True))
    abis = H.to_abi_sw(SY.make_sw_state(ncol, nlay, seed=3, clouds=True))
    lw, sw = LWEngine(), SWEngine()
    ref_lw = lw.run_host(ncol, nlay, abi)
    ref_sw = sw.run_host(ncol, nlay, abis, dyofyr=40)
    out_lw = lw.run_host(ncol, nlay, abi, wait=False)
    out_sw = sw.run_host(ncol, nlay, abis, dyofyr=40, wait=False)
    with pytest.raises(ValueError, match="has not been waited for"):
        lw.run_host(ncol, nlay, abi, wait=False)
    lw.wait()
    sw.wait()
    for k in ref_lw:
        np.testing.assert_array_equal(out_lw[k], ref_lw[k], err_msg=k)
        np.testing.assert_array_equal(out_sw[k], ref_sw[k], err_msg=k)
    h2d, d2h = lw.last_transfer_bytes
    L, n = nlay, ncol
    # 12 p/T/gas + 5 cloud-physics layer fields, emis; no taucld (inflag=2); the all-zero tauaer is set by a memset
    assert h2d == 8 * n * (17 * L + 2 * (L + 1) + 1 + 16)
    assert d2h == 8 * n * (4 * (L + 1) + 2 * L)
    h2d_sw, _ = sw.last_transfer_bytes
    assert h2d_sw == 8 * n * (13 * L + 2 * (L + 1) + 6)             # no direct cloud optics, no aerosol arrays (iaer=0)
    lw.close()
    sw.close()


def test_all_zero_inputs_do_not_cross_pcie_and_results_are_unchanged(monkeypatch):
    """Halocarbons, aerosol optical depth and cloud arrays that are zero everywhere are replaced by a device memset; a single
    non-zero element anywhere makes the array travel again.  Bit-identical to the device-pointer call either way."""
    from climt_b200.engine import LWEngine, SWEngine
    monkeypatch.setenv("CLIMT_B200_HOST_CHUNK", "256")
    ncol, nlay = 1000, 40
    L, n = nlay, ncol
    abi = H.to_abi(SY.make_lw_state(ncol, nlay, seed=5, clouds=False))
    abis = H.to_abi_sw(SY.make_sw_state(ncol, nlay, seed=5, clouds=False))
    for k in ("cfc11vmr", "cfc12vmr", "cfc22vmr", "ccl4vmr"):
        abi[k][:] = 0.0
    assert not abi["tauaer"].any() and not abi["cldfr"].any()
    lw, sw = LWEngine(), SWEngine()
    host = lw.run_host(ncol, nlay, abi)
    assert lw.last_transfer_bytes[0] == 8 * n * (8 * L + 2 * (L + 1) + 1 + 16)
    dev = _device_lw(lw, ncol, nlay, abi)
    hs = sw.run_host(ncol, nlay, abis, dyofyr=100)
    assert sw.last_transfer_bytes[0] == 8 * n * (8 * L + 2 * (L + 1) + 6)
    ds = _device_sw(sw, ncol, nlay, abis, dyofyr=100)
    for k in dev:
        np.testing.assert_array_equal(host[k], dev[k], err_msg=k)
        np.testing.assert_array_equal(hs[k], ds[k], err_msg=k)
    # the scan switched off: every live array travels, same bits
    monkeypatch.setenv("CLIMT_B200_SKIP_ZERO_INPUTS", "0")
    lw0 = LWEngine()
    host0 = lw0.run_host(ncol, nlay, abi)
    assert lw0.last_transfer_bytes[0] == 8 * n * (17 * L + 2 * (L + 1) + 1 + 16 + 16 * L)
    for k in dev:
        np.testing.assert_array_equal(host0[k], dev[k], err_msg=k)
    lw0.close()
    monkeypatch.delenv("CLIMT_B200_SKIP_ZERO_INPUTS")
    # one aerosol element and one CFC-12 element in the last chunk's last column
    abi2 = {k: v.copy() for k, v in abi.items()}
    abi2["tauaer"].reshape(-1)[-1] = 0.35
    abi2["tauaer"][:, 3, -1] = 0.2
    abi2["cfc12vmr"][2, -1] = 5e-10
    host2 = lw.run_host(ncol, nlay, abi2)
    base, last = 8 * n * (8 * L + 2 * (L + 1) + 1 + 16), n - 3 * 256   # the scan works chunk by chunk: only the last chunk's share travels
    assert lw.last_transfer_bytes[0] == base + 8 * last * (L + 16 * L)
    dev2 = _device_lw(lw, ncol, nlay, abi2)
    for k in dev2:
        np.testing.assert_array_equal(host2[k], dev2[k], err_msg=k)
    assert np.abs(host2["dflx"][:, -1] - host["dflx"][:, -1]).max() > 1e-6
    np.testing.assert_array_equal(host2["dflx"][:, :-1], host["dflx"][:, :-1])
    # -0.0 is not "zero": the array is transferred (and gives the same fluxes)
    abi3 = {k: v.copy() for k, v in abi.items()}
    abi3["cfc22vmr"][0, 0] = -0.0
    host3 = lw.run_host(ncol, nlay, abi3)
    assert lw.last_transfer_bytes[0] == base + 8 * 128 * L   # chunk 0 of the ramp (128, 128, 256, 256, 232 columns)
    for k in dev:
        np.testing.assert_array_equal(host3[k], dev[k], err_msg=k)
    lw.close()
    sw.close()


@pytest.mark.gpu
def test_device_side_marshal_of_the_host_call_matches_the_numpy_marshal(monkeypatch):
    """cb200_{lw,sw}_set_host_marshal: q -> vmr and the ln-p interface temperatures evaluated by the engine on the device, chunk by
    chunk (the components' default), against the reference's numpy expressions on the host (CLIMT_B200_HOST_MARSHAL=numpy)."""
    from climt_b200 import synthetic as SY
    from climt_b200.rrtmg_lw import RRTMGLongwave
    from climt_b200.rrtmg_sw import RRTMGShortwave
    lw_state, sw_state = SY.component_states(3000, 40, seed=5, clouds=True)
    out = {}
    for mode in ("device", "numpy"):
        monkeypatch.setenv("CLIMT_B200_HOST_MARSHAL", mode)
        out[mode] = (RRTMGLongwave().array_call(dict(lw_state)), RRTMGShortwave().array_call(dict(sw_state)))
    for which in (0, 1):
        for part in (0, 1):
            a, b = out["device"][which][part], out["numpy"][which][part]
            assert set(a) == set(b)
            for k in a:
                scale = max(float(np.abs(b[k]).max()), 1e-30)
                assert float(np.abs(a[k] - b[k]).max()) <= 1e-11 * scale, (which, k)
