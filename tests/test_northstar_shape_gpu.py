"""Parity at the shape BASELINE.json's north_star scales on (configs[2]: RRTMG LW+SW with McICA clouds, 72 levels): more than
20 000 columns, so that a host-pointer call crosses all three chunk sizes of the host pipeline (1 024, 2 048, 4 096 columns),
both of its slots and a ragged tail, and a device-pointer call crosses the 16 384 / 8 192-column workspace chunks -- with the
per-chunk offset into the whole-call Mersenne-twister mask (`Work::moff`).

* kissvec (per-column seeds): host path == device path bit for bit, both == the oracle on a strided sample of columns
  (the generator of a column depends on that column alone, so the oracle can run on the sample);
* Mersenne twister (ONE stream over the whole call, mcica_subcol_gen_lw.f90:360-368): host path == device path, both == the
  oracle run on the WHOLE grid (a column's draws depend on its global index and on ncol).
Tolerance 1e-6 relative (north_star); observed ~1e-13.
"""
import numpy as np
import pytest

import helpers as H
from climt_b200 import synthetic as SY

pytestmark = pytest.mark.gpu
NCOL, NLAY, SEED = 20500, 72, 112
RTOL = 1e-6
FLUX = ("uflx", "dflx", "uflxc", "dflxc")


def _take(st, idx, nb):
    return {k: np.ascontiguousarray(np.take(v, idx, axis=(1 if v.ndim == 3 and v.shape[-1] == nb else v.ndim - 1))) for k, v in st.items()}


def _device_call(eng, shapes_fn, names_in, arrays, ncol, **kw):
    import torch
    from climt_b200.engine import LW_OUT
    _, outs = shapes_fn(ncol, NLAY)
    d_in = {k: torch.from_numpy(np.ascontiguousarray(arrays[k])).cuda() for k in names_in}
    d_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in LW_OUT}
    eng.run_device(ncol, NLAY, d_in, d_out, **kw)
    torch.cuda.synchronize()
    eng.check()
    return {k: v.cpu().numpy() for k, v in d_out.items()}


@pytest.fixture(scope="module")
def lw_state():
    return SY.make_lw_state(NCOL, NLAY, seed=31, clouds=True, aerosol=True)


@pytest.fixture(scope="module")
def sw_state():
    return SY.make_sw_state(NCOL, NLAY, seed=31, clouds=True, overcast_only=False)


@pytest.mark.parametrize("icld", [1, 2, 3])
def test_lw_mcica_kissvec_host_device_oracle(lw_state, icld):
    from climt_b200.engine import LWEngine, LW_IN, lw_shapes
    from oracle.rrtmg import lw_mcica
    eng = LWEngine(icld=icld, mcica=True, irng=0, permuteseed=SEED)
    abi = H.to_abi(lw_state)
    host = eng.run_host(NCOL, NLAY, abi)
    dev = _device_call(eng, lw_shapes, LW_IN, abi, NCOL)
    eng.close()
    for k in FLUX + ("hr", "hrc"):
        np.testing.assert_array_equal(host[k], dev[k])          # same kernels, different chunking: same bits
    idx = np.arange(7, NCOL, 97)                                # 212 columns across every chunk of both paths
    ref = lw_mcica(H.lw_oracle(cloud_overlap=icld), _take(lw_state, idx, 16), SEED, irng=0)
    for k in FLUX:
        assert H.rel_err(host[k][:, idx], ref[k]) < RTOL, (k, H.rel_err(host[k][:, idx], ref[k]))
    np.testing.assert_allclose(host["hr"][:, idx], ref["hr"], rtol=1e-5, atol=1e-7)
    assert np.abs(host["uflx"] - host["uflxc"]).max() > 1.0     # clouds present


@pytest.mark.parametrize("icld", [1, 2, 3])
def test_sw_mcica_kissvec_host_device_oracle(sw_state, icld):
    from climt_b200.engine import SWEngine, SW_IN, sw_shapes
    from oracle.rrtmg import sw_mcica
    eng = SWEngine(icld=icld, mcica=True, irng=0, permuteseed=SEED)
    abi = H.to_abi_sw(sw_state)
    host = eng.run_host(NCOL, NLAY, abi, dyofyr=172)
    dev = _device_call(eng, sw_shapes, SW_IN, abi, NCOL, dyofyr=172)
    eng.close()
    for k in FLUX + ("hr", "hrc"):
        np.testing.assert_array_equal(host[k], dev[k])
    idx = np.arange(11, NCOL, 97)
    ref = sw_mcica(H.sw_oracle(cloud_overlap=icld), _take(sw_state, idx, 14), SEED, irng=0, dyofyr=172)
    for k in FLUX:
        assert H.rel_err(host[k][:, idx], ref[H.SW_KEYS[k]]) < RTOL, (k, H.rel_err(host[k][:, idx], ref[H.SW_KEYS[k]]))
    assert np.abs(host["dflx"] - host["dflxc"]).max() > 1.0


def test_lw_mcica_mersenne_twister_whole_grid(lw_state):
    from climt_b200.engine import LWEngine, LW_IN, lw_shapes
    from oracle.rrtmg import lw_mcica
    eng = LWEngine(icld=2, mcica=True, irng=1, permuteseed=SEED)
    abi = H.to_abi(lw_state)
    host = eng.run_host(NCOL, NLAY, abi)
    dev = _device_call(eng, lw_shapes, LW_IN, abi, NCOL)
    eng.close()
    for k in FLUX + ("hr", "hrc"):
        np.testing.assert_array_equal(host[k], dev[k])
    ref = lw_mcica(H.lw_oracle(cloud_overlap=2), lw_state, SEED, irng=1)
    for k in FLUX:
        assert H.rel_err(host[k], ref[k]) < RTOL, (k, H.rel_err(host[k], ref[k]))


def test_sw_mcica_mersenne_twister_whole_grid(sw_state):
    from climt_b200.engine import SWEngine, SW_IN, sw_shapes
    from oracle.rrtmg import sw_mcica
    eng = SWEngine(icld=2, mcica=True, irng=1, permuteseed=SEED)
    abi = H.to_abi_sw(sw_state)
    host = eng.run_host(NCOL, NLAY, abi, dyofyr=172)
    dev = _device_call(eng, sw_shapes, SW_IN, abi, NCOL, dyofyr=172)
    eng.close()
    for k in FLUX + ("hr", "hrc"):
        np.testing.assert_array_equal(host[k], dev[k])
    ref = sw_mcica(H.sw_oracle(cloud_overlap=2), sw_state, SEED, irng=1, dyofyr=172)
    for k in FLUX:
        assert H.rel_err(host[k], ref[H.SW_KEYS[k]]) < RTOL, (k, H.rel_err(host[k], ref[H.SW_KEYS[k]]))
