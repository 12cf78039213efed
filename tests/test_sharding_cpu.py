"""The N > 1 path on CPU: two gloo ranks shard a column grid, 'compute' a column-independent function on their block and
reassemble the global field with the single all-gather bench.py uses over NCCL."""
import os
import socket

import numpy as np
import pytest

from climt_b200 import sharding as SH


def test_shard_bounds_cover_grid_without_overlap():
    for ncol, world in ((8192, 8), (1000, 8), (7, 8), (131072, 3), (1, 2)):
        b = SH.shard_bounds(ncol, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == ncol
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(hi >= lo for lo, hi in b)
        assert max(hi - lo for lo, hi in b) == -(-ncol // world)


def test_shard_arrays_finds_the_column_axis():
    ncol, nlev = 10, 4
    arrays = {"t": np.arange(nlev * ncol, dtype=float).reshape(nlev, ncol), "ts": np.arange(ncol, dtype=float),
              "taucld": np.arange(nlev * ncol * 3, dtype=float).reshape(nlev, ncol, 3), "emis": np.ones((16, ncol)), "scalar": np.array(2.0)}
    parts = [SH.shard_arrays(arrays, ncol, r, 3) for r in range(3)]
    assert [p["t"].shape[1] for p in parts] == [4, 4, 2]
    np.testing.assert_array_equal(np.concatenate([p["t"] for p in parts], axis=1), arrays["t"])
    np.testing.assert_array_equal(np.concatenate([p["taucld"] for p in parts], axis=1), arrays["taucld"])
    np.testing.assert_array_equal(np.concatenate([p["ts"] for p in parts]), arrays["ts"])
    assert all(p["t"].flags.c_contiguous for p in parts) and float(parts[0]["scalar"]) == 2.0
    with pytest.raises(ValueError):
        SH.shard_arrays({"bad": np.ones((3, 5))}, ncol, 0, 2)


def _fake_radiation(t, ts):
    """column-independent stand-in for an engine call: fluxes on nlev+1 interfaces, heating on nlev layers"""
    up = np.concatenate([ts[None] ** 2, np.cumsum(t, axis=0) + ts[None]], axis=0)
    return {"uflx": up, "hr": np.diff(up, axis=0) * 0.5}


def _worker(rank, world, port, ncol, nlev, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    arrays = {"t": rng.normal(size=(nlev, ncol)), "ts": rng.normal(size=ncol)}
    mine = SH.shard_arrays(arrays, ncol, rank, world)
    out = _fake_radiation(mine["t"], mine["ts"])
    local = {k: torch.from_numpy(v) for k, v in out.items()}
    full = SH.all_gather_columns(local, ["uflx", "hr"], ncol)
    ref = _fake_radiation(arrays["t"], arrays["ts"])
    ok = all(np.array_equal(full[k].numpy(), ref[k]) for k in ref) and full["uflx"].shape == (nlev + 1, ncol)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ncol", [64, 37])  # even split and a ragged last block
def test_two_rank_gloo_all_gather_reassembles_global_field(ncol):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ncol, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _packed_worker(rank, world, port, ncol, nlev, q):
    """what bench.py --gpus N does per step: the 'engine' writes into row slices of the rank's packed buffer, one all-gather"""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    arrays = {"t": rng.normal(size=(nlev, ncol)), "ts": rng.normal(size=ncol)}
    mine = SH.shard_arrays(arrays, ncol, rank, world)
    out = _fake_radiation(mine["t"], mine["ts"])
    po = SH.PackedOutputs([("uflx", nlev + 1), ("hr", nlev)], ncol // world, world, device="cpu")
    for k, v in out.items():
        po.views[k].copy_(torch.from_numpy(v))      # the engines write here directly
    po.gather_async()
    ref = _fake_radiation(arrays["t"], arrays["ts"])
    ok = all(np.array_equal(po.global_field(k).numpy(), ref[k]) for k in ref)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_packed_outputs_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_packed_worker, args=(r, 2, port, 48, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
