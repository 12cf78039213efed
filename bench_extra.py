#!/usr/bin/env python
"""Benchmarks of the other BASELINE.json configurations (bench.py keeps the headline contract: configs[1], RRTMG LW+SW).

  python bench_extra.py --workload mcica   # configs[2] per-GPU share: RRTMG LW+SW with McICA clouds (KISS RNG), 16384 col x 72 lev
  python bench_extra.py --workload cork    # configs[3] per-GPU share: CORK correlated-k LW+SW, 65536 col x 60 lev
  python -m torch.distributed.run --nproc-per-node N ... bench_extra.py --workload cork --gpus N   # weak scaling, columns sharded

Same measurement rules as bench.py: W >= 3 warm-up steps, CUDA events around K steps, max over ranks, inputs resident in HBM
for `value`; `e2e` = pinned host buffers through the host-pointer C ABI.  One JSON line per run on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from bench import ClockSampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=["mcica", "cork"], required=True)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ncol", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import helpers as H
    from climt_b200 import synthetic as SY

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise SystemExit("bench_extra.py needs a CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K = max(args.warmup, 3), args.steps

    if args.workload == "mcica":
        from climt_b200.engine import LWEngine, SWEngine, LW_IN, LW_OUT, SW_IN, lw_shapes
        ncol, nlay = args.ncol or 16384, 72
        st = SY.make_lw_state(ncol, nlay, seed=20260925 + rank, clouds=True)
        sts = SY.make_sw_state(ncol, nlay, seed=20260925 + rank, clouds=True, overcast_only=False)
        abi, abis = H.to_abi(st), H.to_abi_sw(sts)
        kw = dict(icld=2, mcica=True, irng=0, permuteseed=112, device=local)
        lw, sw = LWEngine(**kw), SWEngine(**kw)
        _, outs = lw_shapes(ncol, nlay)
        d_in = {k: torch.from_numpy(abi[k]).cuda() for k in LW_IN}
        ds_in = {k: torch.from_numpy(abis[k]).cuda() for k in SW_IN}
        d_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in LW_OUT}
        ds_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in LW_OUT}
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
        h_in, hs_in = {k: pin(abi[k]) for k in LW_IN}, {k: pin(abis[k]) for k in SW_IN}
        h_out, hs_out = {k: pin(np.empty(outs[k])) for k in LW_OUT}, {k: pin(np.empty(outs[k])) for k in LW_OUT}

        def step_device():
            lw.run_device(ncol, nlay, d_in, d_out)
            sw.run_device(ncol, nlay, ds_in, ds_out, dyofyr=1)

        def step_host():
            lw.run_host(ncol, nlay, h_in, h_out, wait=False)
            sw.run_host(ncol, nlay, hs_in, hs_out, dyofyr=1, wait=False)
            lw.wait()
            sw.wait()
        name = "RRTMG LW+SW McICA columns/s (72 lev)"
        workload = f"RRTMG LW+SW with McICA clouds (maximum-random overlap, KISS generator), {ncol} columns x 72 levels per GPU (BASELINE.json configs[2] is 131072 columns)"
        launches = lambda: lw.last_launches + sw.last_launches  # noqa: E731
        xfer = lambda: tuple(a + b for a, b in zip(lw.last_transfer_bytes, sw.last_transfer_bytes))  # noqa: E731

        def cpu():
            from oracle.rrtmg import lw_mcica, sw_mcica
            n = 64
            sub = {k: (v[:, :n] if v.ndim == 2 else (v[:, :n, :] if v.ndim == 3 and v.shape[-1] in (14, 16) else v[..., :n])) for k, v in st.items()}
            subs = {k: (v[:, :n] if v.ndim == 2 else (v[:, :n, :] if v.ndim == 3 and v.shape[-1] in (14, 16) else v[..., :n])) for k, v in sts.items()}
            olw, osw = H.lw_oracle(cloud_overlap=2), H.sw_oracle(cloud_overlap=2)
            t0, reps = time.perf_counter(), 0
            while time.perf_counter() - t0 < 10.0:
                lw_mcica(olw, sub, 112, irng=0)
                sw_mcica(osw, subs, 112, irng=0, dyofyr=1)
                reps += 1
            dt = time.perf_counter() - t0
            return reps * n / dt, f"{n} of the workload's {ncol} columns x {reps} repetitions in {dt:.1f} s (C++ restatement, 1 core)"
    else:
        from climt_b200 import cork
        ncol, nlay = args.ncol or 65536, 60
        rng = np.random.default_rng(20260925 + rank)
        lws = SY.make_lw_state(ncol, nlay, seed=20260925 + rank)
        s = {"T": lws["tlay"], "p": lws["play"] * 100.0, "p_int": lws["plev"] * 100.0, "T_surf": lws["tsfc"], "q": lws["h2o"] * 0.622,
             "co2": np.full((nlay, ncol), 4e-4), "emissivity": np.ones((14, ncol)), "tau_cloud_lw": np.zeros((nlay, ncol, 14)),
             "zenith": np.deg2rad(rng.uniform(0, 85, ncol)), "albedo": rng.uniform(0.06, 0.3, ncol),
             "tau_cloud_sw": np.zeros((nlay, ncol, 3)), "ssa_cloud": np.zeros((nlay, ncol, 3)), "g_cloud": np.zeros((nlay, ncol, 3)),
             "earth_sun_factor": np.ones(ncol)}
        el, es = cork.CorkEngine("earth_low_res_lw", device=local), cork.CorkEngine("earth_low_res_sw", device=local)
        al, as_ = H.cork_arrays(s, "lw"), H.cork_arrays(s, "sw")
        names_l = ["up_broad", "down_broad", "heating_rate", "up_band", "down_band", "tau_band", "hr_band", "trans_band"]
        names_s = names_l[:-1]
        _, outs_l = el.shapes(ncol, nlay)
        _, outs_s = es.shapes(ncol, nlay)
        dl_in = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in al.items()}
        dsw_in = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in as_.items()}
        dl_out = {k: torch.empty(outs_l[k], dtype=torch.float64, device="cuda") for k in names_l}
        dsw_out = {k: torch.empty(outs_s[k], dtype=torch.float64, device="cuda") for k in names_s}
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory().numpy()  # noqa: E731
        hl_in, hs_in = {k: pin(v) for k, v in al.items()}, {k: pin(v) for k, v in as_.items()}
        hl_out = {k: pin(np.empty(outs_l[k])) for k in names_l}
        hs_out = {k: pin(np.empty(outs_s[k])) for k in names_s}

        def step_device():
            el.lw_device(ncol, nlay, dl_in, dl_out)
            es.sw_device(ncol, nlay, dsw_in, dsw_out)

        def step_host():
            el.lw_host(ncol, nlay, hl_in, out=hl_out)
            es.sw_host(ncol, nlay, hs_in, out=hs_out)
        name = "CORK LW+SW columns/s (60 lev)"
        workload = f"CORK correlated-k LW (14 bands x 8 g) + SW (3 x 2), earth_low_res tables, per-band diagnostics on, {ncol} columns x 60 levels per GPU (BASELINE.json configs[3] is 524288 columns on 8 GPUs)"
        launches = lambda: el.last_launches + es.last_launches  # noqa: E731
        h2d = sum(v.nbytes for v in hl_in.values()) + sum(v.nbytes for v in hs_in.values())
        d2h = sum(v.nbytes for v in hl_out.values()) + sum(v.nbytes for v in hs_out.values())
        xfer = lambda: (h2d, d2h)  # noqa: E731

        def cpu():
            from oracle import cork as OC
            n = 128
            sub = {k: (v[..., :n] if v.ndim < 3 else v[:, :n]) for k, v in s.items()}
            sub["emissivity"] = s["emissivity"][:, :n]
            t0, reps = time.perf_counter(), 0
            while time.perf_counter() - t0 < 10.0:
                OC.lw_call(el.table, sub, H.CORK_G, H.CORK_CPD, H.CORK_SIGMA)
                OC.sw_call(es.table, sub, H.CORK_G, H.CORK_CPD)
                reps += 1
            dt = time.perf_counter() - t0
            return reps * n / dt, f"{n} of the workload's {ncol} columns x {reps} repetitions in {dt:.1f} s (C++ restatement of the numba kernels, 1 core)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    nl = launches() * K
    for _ in range(W):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_host()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = [float(x) for x in t.tolist()]
    if rank == 0:
        sampler.stop.set()
        sampler.join(timeout=2)
        h2d_b, d2h_b = xfer()
        line = {"metric": name, "value": world * ncol * K / (ms * 1e-3), "unit": "columns/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "columns_per_gpu": ncol, "levels": nlay,
                           "cache": "per-g-point scratch of one step (> 10 GB) exceeds the 126 MB L2"},
                "e2e": {"value": world * ncol * K / (e2e_ms * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b},
                "gpu_launches": nl, "clocks": sampler.summary()}
        if not args.no_cpu_baseline and world == 1:
            v, sample = cpu()
            line["cpu_baseline"] = {"value": v, "unit": "columns/s", "cores": 1, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
