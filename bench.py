#!/usr/bin/env python
"""Benchmark of the column radiative-transfer hot path (contract: the task statement / DESIGN.md 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one full RRTMG LW + SW evaluation of the workload's columns (every kernel of both engines).

N = 1   workload = the grid of BASELINE.json configs[1]: 128x64 columns x 60 levels, clear sky, fp64, LW + SW.
        `value`  columns/s with the inputs resident in HBM (CUDA events);
        `e2e`    the same through the host-pointer C ABI on pinned host buffers (H2D + kernels + D2H inside the timed region);
        extra keys: `e2e_component` (through RRTMGLongwave / RRTMGShortwave.array_call on pageable numpy state -- the call a climt
        user makes), `config2_n1` (configs[2] -- McICA, 512x256 columns x 72 levels -- on this one GPU: the N=1 point of the
        strong-scaling series below), `roofline`, `cpu_baseline`.
N > 1   (torchrun, one rank per GPU) workload = BASELINE.json configs[2]: RRTMG LW+SW with McICA clouds (maximum-random overlap,
        kissvec generator), 512x256 = 131072 columns x 72 levels, STRONG scaling: the fixed global grid is cut into contiguous
        column blocks (climt_b200/sharding.py), no data-path collective, one all-gather per step reassembles the global flux /
        heating-rate field on every rank.  The global state does not depend on N (8 blocks of 16384 columns with fixed seeds).
        extra keys: `weak_clear_sky` (the configs[1] grid per GPU, weak scaling, last round's series), `n1_same_workload`
        (rank 0 alone on the whole grid, measured after the timed region: the strong-scaling denominator in the same run).

--impl reference: the reference's CPU implementation of the same workload on the box's host cores (see reference_arm()).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

NCOL, NLAY = 128 * 64, 60                      # configs[1]
WORKLOAD = "RRTMG LW+SW clear-sky, 128x64 columns x 60 levels, fp64 (grid of BASELINE.json configs[1])"
NCOL2, NLAY2, BLOCK2, SEED2 = 512 * 256, 72, 16384, 112   # configs[2]
WORKLOAD2 = ("RRTMG LW+SW with McICA clouds (maximum-random overlap, kissvec generator), 512x256 columns x 72 levels, fp64 "
             "(BASELINE.json configs[2])")
METRIC = "RRTMG LW+SW columns/s (60 lev)"
METRIC2 = "RRTMG LW+SW columns/s (72 lev, McICA)"
# reference ABI, every array the wrappers read/write
ALG_BYTES_LW = (49 * NLAY + 2 * (NLAY + 1) + 17 + 4 * (NLAY + 1) + 2 * NLAY) * 8
ALG_BYTES_SW = (117 * NLAY + 2 * (NLAY + 1) + 6 + 4 * (NLAY + 1) + 2 * NLAY) * 8
# measured DRAM bytes (read + write) of ONE launch of the dominant kernels on the configs[1] grid, from the `ncu --set full`
# captures under profiles/ (file names in the roofline note)
NCU = json.load(open(os.path.join(ROOT, "profiles", "ncu_dram_bytes.json")))
PUBLISHED_REFERENCE = ("reference's own published RRTMG-LW figure: 45-91 us/column at 30 levels, 100-1000 columns, one laptop core "
                       "(docs/radiative-transfer/performance.qmd:17,25-26; docs/superpowers/plans/2026-05-16-cork-co2-band-refinement.md:1617)")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------------------
# states
def clear_sky_states(ncol, nlay, seed):
    from climt_b200 import synthetic as SY
    return SY.make_lw_state(ncol, nlay, seed=seed), SY.make_sw_state(ncol, nlay, seed=seed)


def mcica_block_states(block):
    """block b (16384 columns) of the configs[2] grid: the global state is the concatenation of 8 such blocks, whatever N is."""
    from climt_b200 import synthetic as SY
    lw = SY.make_lw_state(BLOCK2, NLAY2, seed=20260925 + block, clouds=True)
    sw = SY.make_sw_state(BLOCK2, NLAY2, seed=20260925 + block, clouds=True, overcast_only=False)
    return lw, sw


def concat_columns(states):
    """[{name: array}] of 16384-column blocks -> one state, columns concatenated"""
    from climt_b200.sharding import column_axis
    if len(states) == 1:
        return states[0]
    return {k: np.concatenate([s[k] for s in states], axis=column_axis(states[0][k].shape, BLOCK2)) for k in states[0]}


# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    """LW + SW engines over one rank's columns: device-resident step (two streams) and host-pointer step (two pipelines)."""

    def __init__(self, lw_state, sw_state, nlay, local, world, engine_kw=None, dyofyr=1, host=True):
        import torch
        import helpers as H
        from climt_b200.engine import LWEngine, SWEngine, LW_IN, LW_OUT, SW_IN, lw_shapes
        from climt_b200.sharding import PackedOutputs
        self.torch = torch
        self.ncol, self.nlay, self.dyofyr, self.world = lw_state["play"].shape[1], nlay, dyofyr, world
        kw = dict(engine_kw or {}, device=local)
        self.lw, self.sw = LWEngine(**kw), SWEngine(**kw)
        abi, abis = H.to_abi(lw_state), H.to_abi_sw(sw_state)
        _, outs = lw_shapes(self.ncol, nlay)
        self.d_in = {k: torch.from_numpy(np.ascontiguousarray(abi[k])).cuda() for k in LW_IN}
        self.ds_in = {k: torch.from_numpy(np.ascontiguousarray(abis[k])).cuda() for k in SW_IN}
        # The 12 output fields of a rank live in ONE buffer the engines write into directly (sharding.PackedOutputs): one
        # all-gather, no packing copy.  Two buffers alternate so that the gather of step i overlaps the kernels of step i+1.
        fields = [("lw_" + k, outs[k][0]) for k in LW_OUT] + [("sw_" + k, outs[k][0]) for k in LW_OUT]
        self.packed = [PackedOutputs(fields, self.ncol, world) for _ in range(2)]
        self.d_out = [{k: p.views["lw_" + k] for k in LW_OUT} for p in self.packed]
        self.ds_out = [{k: p.views["sw_" + k] for k in LW_OUT} for p in self.packed]
        self.s_lw, self.s_sw = torch.cuda.Stream(), torch.cuda.Stream()
        self.istep = 0
        self.h = None
        if host:
            def pin(a):
                return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
            # arrays the option flags make dead (direct cloud optics under inflag = 2, SW aerosol arrays under iaer = 0) never cross
            # PCIe and are never read by the host calls: no pinned copy for them
            dead = {"taucld", "ssacld", "asmcld", "fsfcld", "ssaaer", "asmaer", "ecaer"}
            self.h = ({k: (abi[k] if k in dead else pin(abi[k])) for k in LW_IN}, {k: pin(np.empty(outs[k])) for k in LW_OUT},
                      {k: (abis[k] if k in dead | {"tauaer"} else pin(abis[k])) for k in SW_IN},
                      {k: pin(np.empty(outs[k])) for k in LW_OUT})

    def step_device(self, lw=True, sw=True, gather=True):
        torch = self.torch
        b = self.istep & 1
        self.istep += 1
        self.packed[b].wait()
        if lw and sw:
            # the two engines are independent: two streams, so one engine's kernels fill the tail waves of the other's
            cur = torch.cuda.current_stream()
            self.s_lw.wait_stream(cur)
            self.lw.run_device(self.ncol, self.nlay, self.d_in, self.d_out[b], stream=self.s_lw.cuda_stream)
            self.s_sw.wait_stream(cur)
            self.sw.run_device(self.ncol, self.nlay, self.ds_in, self.ds_out[b], dyofyr=self.dyofyr, stream=self.s_sw.cuda_stream)
            cur.wait_stream(self.s_lw)
            cur.wait_stream(self.s_sw)
        elif lw:
            self.lw.run_device(self.ncol, self.nlay, self.d_in, self.d_out[b])
        elif sw:
            self.sw.run_device(self.ncol, self.nlay, self.ds_in, self.ds_out[b], dyofyr=self.dyofyr)
        if gather:
            self.packed[b].gather_async()

    def drain(self):
        for p in self.packed:
            p.wait()

    def step_host(self):
        # both host-pointer calls are enqueued, then completed: the LW and SW pipelines (H2D | kernels | D2H) overlap
        hi, ho, hsi, hso = self.h
        self.lw.run_host(self.ncol, self.nlay, hi, ho, wait=False)
        self.sw.run_host(self.ncol, self.nlay, hsi, hso, dyofyr=self.dyofyr, wait=False)
        self.lw.wait()
        self.sw.wait()

    def launches(self):
        return self.lw.last_launches + self.sw.last_launches

    def transfer_bytes(self):
        (a, b), (c, d) = self.lw.last_transfer_bytes, self.sw.last_transfer_bytes
        return a + c, b + d

    def close(self):
        self.lw.close()
        self.sw.close()


def barrier(torch, dist, world, wl=None):
    if wl is not None:
        wl.drain()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def time_device(torch, dist, world, wl, W, K, **kw):
    """W untimed + K timed device-resident steps, barrier + synchronize on both sides, CUDA events; ms for the K steps"""
    for _ in range(W):
        wl.step_device(**kw)
    barrier(torch, dist, world, wl)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        wl.step_device(**kw)
    wl.drain()  # the last gathers are part of the timed region
    e1.record()
    barrier(torch, dist, world, wl)
    return e0.elapsed_time(e1)


def time_host(torch, dist, world, wl, W, K):
    for _ in range(W):
        wl.step_host()
    barrier(torch, dist, world)
    t0 = time.perf_counter()
    for _ in range(K):
        wl.step_host()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return (time.perf_counter() - t0) * 1e3


def max_over_ranks(torch, dist, world, vals):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs
def _take(d, lo, hi):
    return {k: np.ascontiguousarray(v[:, lo:hi, :] if (v.ndim == 3 and v.shape[-1] in (14, 16)) else v[..., lo:hi]) for k, v in d.items()}


def oracle_columns_per_s(st, threads, min_seconds=10.0, max_cols=None, mcica=False):
    """Time the CPU restatement (the reference's algorithm, column-serial like the Fortran) on a bounded sample."""
    import helpers as H
    from concurrent.futures import ThreadPoolExecutor
    lw, sw = st
    ncol = lw["play"].shape[1]
    n = min(ncol, max_cols or ncol)
    blocks = [(_take(lw, i * n // threads, (i + 1) * n // threads), _take(sw, i * n // threads, (i + 1) * n // threads))
              for i in range(threads)]
    if mcica:
        from oracle.rrtmg import lw_mcica, sw_mcica
        orc, orcs = H.lw_oracle(cloud_overlap=2), H.sw_oracle(cloud_overlap=2)

        def one(b):
            lw_mcica(orc, b[0], SEED2, irng=0)
            sw_mcica(orcs, b[1], SEED2, irng=0, dyofyr=1)
    else:
        orc, orcs = H.lw_oracle(cloud_overlap=1), H.sw_oracle()

        def one(b):
            H.run_lw_oracle(orc, b[0])
            orcs(b[1], adjes=1.0, dyofyr=1, solcycfrac=0.0)
    one(blocks[0])     # warm
    reps, t0 = 0, time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        while True:
            list(ex.map(one, blocks))
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds:
                break
    return reps * n / dt, f"{n} of the workload's {ncol} columns x {reps} repetitions in {dt:.1f} s"


def native_oracle():
    """The timing build of the C++ restatement: -O3 -march=native, FP contraction allowed (the parity build keeps -O2
    -ffp-contract=off).  -march=native code does not travel between hosts, so it is compiled on the box that times it
    (15 s, outside every timed region) and keyed by the CPU model."""
    cpu = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith(("model name", "flags")):
                cpu += line
                if line.startswith("flags"):
                    break
    except OSError:
        pass
    tag = hashlib.sha256(cpu.encode()).hexdigest()[:12]
    d = os.path.join(ROOT, "oracle", "_fast")
    so = os.path.join(d, f"liborc_rrtmg_native_{tag}.so")
    flags = ["-O3", "-march=native", "-std=c++17", "-fPIC", "-shared"]
    if not os.path.exists(so):
        os.makedirs(d, exist_ok=True)
        src = [os.path.join(ROOT, "oracle", f) for f in ("rrtmg_lw_oracle.cpp", "rrtmg_sw_oracle.cpp")]
        subprocess.check_call(["g++"] + flags + ["-o", so + ".tmp"] + src)
        os.replace(so + ".tmp", so)
    return so, "g++ " + " ".join(flags[:2])


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, on this arm's workload.  The Fortran cannot be
    compiled in this image (no Fortran compiler: SURVEY.md 8c), so this is the C++ restatement of it (oracle/), built for speed
    (native_oracle()), columns block-partitioned over every host core -- the reference itself runs the column loop on ONE core
    (rrtmg_lw_rad.nomcica.f90:453, no OpenMP).  Every step is the same fixed sample on every box (`same_config`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ctypes
    import oracle.rrtmg as ORC
    so, flags = native_oracle()
    ORC._LIB = ctypes.CDLL(so)
    ORC._LIB.orc_last_error.restype = ctypes.c_char_p
    threads = os.cpu_count() or 1
    if args.gpus == 1:
        st = clear_sky_states(NCOL, NLAY, 20260925)
        ncs, metric, workload, nlay, mc = NCOL, METRIC, WORKLOAD, NLAY, False
        sample = f"all {NCOL} columns of the workload per step"
    else:
        st = mcica_block_states(0)
        ncs, metric, workload, nlay, mc = 4096, METRIC2, WORKLOAD2, NLAY2, True
        sample = f"the first {ncs} of the workload's {NCOL2} columns per step (the CPU time is linear in columns)"
    vals = []
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, _ = oracle_columns_per_s(st, threads, min_seconds=0.0, max_cols=ncs, mcica=mc)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": "columns/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * ncs / value, "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "columns_per_step": ncs, "levels": nlay, "same_config": True},
        "cpu_baseline": {"value": value, "unit": "columns/s", "cores": threads, "kind": "port", "flags": flags,
                         "sample": f"{sample}, {args.steps} steps; C++ restatement of the reference Fortran (gfortran absent), "
                                   f"{flags}, columns block-partitioned over {threads} threads",
                         "published": PUBLISHED_REFERENCE},
        "e2e": {"value": value, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all}))


# ---------------------------------------------------------------------------------------------------------------------
def component_e2e(K, W, local):
    """One step the way a climt user makes it: RRTMGLongwave.array_call(state) then RRTMGShortwave.array_call(state) on pageable
    numpy arrays in the components' units (what sympl hands down) -- vmr conversion, interface temperatures, output allocation,
    H2D, kernels, D2H all inside the timed region (climt/_components/rrtmg/lw/component.py:373-393,482-522)."""
    from climt_b200 import synthetic as SY
    from climt_b200.rrtmg_lw import RRTMGLongwave
    from climt_b200.rrtmg_sw import RRTMGShortwave
    lw_raw, sw_raw = SY.component_states(NCOL, NLAY, seed=20260925)
    lw, sw = RRTMGLongwave(device=local), RRTMGShortwave(device=local)
    for _ in range(W):
        lw.array_call(lw_raw)
        sw.array_call(sw_raw)
    t0 = time.perf_counter()
    for _ in range(K):
        lw.array_call(lw_raw)
        sw.array_call(sw_raw)
    dt = time.perf_counter() - t0
    (a, b), (c, d) = lw._engine.last_transfer_bytes, sw._engine.last_transfer_bytes
    return {"value": NCOL * K / dt, "unit": "columns/s", "ms_per_step": 1e3 * dt / K, "h2d_bytes_per_step": a + c,
            "d2h_bytes_per_step": b + d,
            "note": "RRTMGLongwave.array_call + RRTMGShortwave.array_call back to back on pageable numpy state (no pinning, no "
                    "overlap between the two calls): numpy marshal + staged H2D + kernels + D2H"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys (component e2e, configs[2] at N=1, weak series)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K = max(args.warmup, 3), args.steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    mc_kw = dict(icld=2, mcica=True, irng=0, permuteseed=SEED2)
    line = None

    if world == 1:
        # ---- headline: configs[1] grid, clear sky ----------------------------------------------------------------
        st, sts = clear_sky_states(NCOL, NLAY, 20260925)
        wl = Workload(st, sts, NLAY, local, 1)
        ms = time_device(torch, dist, 1, wl, W, K)
        wl.lw.check()
        wl.sw.check()
        launches = wl.launches() * K
        part = {n: time_device(torch, dist, 1, wl, 1, K, **kw) / K
                for n, kw in (("lw", dict(lw=True, sw=False)), ("sw", dict(lw=False, sw=True)))}
        # the dominant kernels timed alone with CUDA events on their launch stream (the engines record them)
        wl.lw.enable_timing(True)
        wl.sw.enable_timing(True)
        km = {"lw_transfer": [], "sw_transfer": [], "lw_taumol": [], "sw_taumol": []}
        for _ in range(max(3, min(K, 10))):
            wl.step_device(lw=True, sw=False, gather=False)
            wl.step_device(lw=False, sw=True, gather=False)
            torch.cuda.synchronize()
            km["lw_transfer"].append(wl.lw.last_unit_kernel_ms)
            km["sw_transfer"].append(wl.sw.last_unit_kernel_ms)
            km["lw_taumol"].append(wl.lw.last_taumol_kernel_ms)
            km["sw_taumol"].append(wl.sw.last_taumol_kernel_ms)
        wl.lw.enable_timing(False)
        wl.sw.enable_timing(False)
        km = {k: float(np.mean(v)) for k, v in km.items()}
        e2e_ms = time_host(torch, dist, 1, wl, W, K)
        h2d, d2h = wl.transfer_bytes()
        scan = {1: "on", 0: "off (environment)", -1: "off (the engine measured the scan slower than the copy it saves)"}[wl.lw.zero_scan_state]
        peaks, which = measured_peaks()
        value = NCOL * K / (ms * 1e-3)
        sw_dom = km["sw_transfer"] >= km["lw_transfer"]
        dom = NCU["sw"]["kernel"] if sw_dom else NCU["lw"]["kernel"]
        dom_ms = max(km["sw_transfer"], km["lw_transfer"])
        dom_bytes = ALG_BYTES_SW if sw_dom else ALG_BYTES_LW
        achieved = dom_bytes * NCOL / (dom_ms * 1e-3) / 1e9
        traffic = NCU["sw" if sw_dom else "lw"]["dram_bytes"]
        line = {
            "metric": METRIC, "value": value, "unit": "columns/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "columns_per_gpu": NCOL, "levels": NLAY, "gpoints": "140 LW + 112 SW",
                       "lw_only_columns_per_s": NCOL / (part["lw"] * 1e-3), "sw_only_columns_per_s": NCOL / (part["sw"] * 1e-3),
                       "cache": "the engines' per-step workspace traffic (GBs, see roofline.traffic) exceeds the 126 MB L2: nothing "
                                "survives between timed iterations",
                       "parallelism": "1 GPU"},
            "e2e": {"value": NCOL * K / (e2e_ms * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "zero_input_scan": scan,
                    "note": "host-pointer C ABI, pinned buffers; bytes as counted by the engines: inputs that are zero everywhere "
                            "(this clear-sky state's aerosol optical depth and cloud arrays) are scanned on the host inside the "
                            "timed region and set by a device memset instead of being copied"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": which,
                         "kernel": dom, "kernel_ms": dom_ms, "alg_bytes_per_column": dom_bytes,
                         "lw_transfer_ms": km["lw_transfer"], "sw_transfer_ms": km["sw_transfer"],
                         "lw_taumol_ms": km["lw_taumol"], "sw_taumol_ms": km["sw_taumol"],
                         "traffic_gbs": traffic / (dom_ms * 1e-3) / 1e9,
                         "traffic_over_algorithmic": traffic / (dom_bytes * NCOL),
                         "lw_transfer": {"kernel": NCU["lw"]["kernel"], "kernel_ms": km["lw_transfer"],
                                         "achieved": ALG_BYTES_LW * NCOL / (km["lw_transfer"] * 1e-3) / 1e9,
                                         "traffic": NCU["lw"]["dram_bytes"],
                                         "traffic_over_algorithmic": NCU["lw"]["dram_bytes"] / (ALG_BYTES_LW * NCOL)},
                         "ncu": NCU,
                         "note": "achieved = algorithmic bytes of the engine call (reference ABI, SURVEY.md 8d) x columns / time of "
                                 "the dominant kernel (CUDA events on its launch stream). `traffic` = DRAM bytes one launch of "
                                 "that kernel really moves (ncu capture named in roofline.ncu); traffic_gbs = that over the same "
                                 "time. The path is ~100 flop per algorithmic byte: it cannot approach the HBM roof by "
                                 "algorithmic bytes (DESIGN.md 3)"},
        }
        if not args.no_extras:
            wl.h = None
            line["e2e_component"] = component_e2e(max(3, K // 2), 2, local)
        wl.close()
        del wl
        torch.cuda.empty_cache()
        if not args.no_extras:
            # ---- configs[2] on this one GPU: the N = 1 point of the strong-scaling series -----------------------------
            K2 = max(3, K // 4)
            blocks = [mcica_block_states(b) for b in range(NCOL2 // BLOCK2)]
            lw2, sw2 = concat_columns([b[0] for b in blocks]), concat_columns([b[1] for b in blocks])
            del blocks
            wl2 = Workload(lw2, sw2, NLAY2, local, 1, engine_kw=mc_kw)
            ms2 = time_device(torch, dist, 1, wl2, 2, K2)
            wl2.lw.check()
            wl2.sw.check()
            e2e2 = time_host(torch, dist, 1, wl2, 1, K2)
            h2, d2 = wl2.transfer_bytes()
            line["config2_n1"] = {"metric": METRIC2, "workload": WORKLOAD2, "value": NCOL2 * K2 / (ms2 * 1e-3), "unit": "columns/s",
                                  "ms_per_step": ms2 / K2, "steps": K2,
                                  "e2e": {"value": NCOL2 * K2 / (e2e2 * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": h2,
                                          "d2h_bytes_per_step": d2}}
            wl2.close()
            del wl2
        line["clocks"] = sampler_summary(sampler)
        if not args.no_cpu_baseline:
            v, sample = oracle_columns_per_s((st, sts), 1, min_seconds=10.0, max_cols=512)
            line["cpu_baseline"] = {"value": v, "unit": "columns/s", "cores": 1, "kind": "port",
                                    "sample": sample + " (C++ restatement of the reference Fortran, parity build -O2 "
                                              "-ffp-contract=off, one core: the reference's column loop is serial)",
                                    "published": PUBLISHED_REFERENCE}
        print(json.dumps(line))
        return

    # ---- N > 1: configs[2], strong scaling ----------------------------------------------------------------------------
    from climt_b200.sharding import shard_bounds
    nblk = NCOL2 // BLOCK2
    if nblk % world:
        raise SystemExit(f"--gpus must divide {nblk} (the configs[2] grid is {nblk} blocks of {BLOCK2} columns)")
    lo, hi = shard_bounds(NCOL2, world)[rank]
    mine = [mcica_block_states(b) for b in range(lo // BLOCK2, hi // BLOCK2)]
    lw2, sw2 = concat_columns([b[0] for b in mine]), concat_columns([b[1] for b in mine])
    del mine
    wl = Workload(lw2, sw2, NLAY2, local, world, engine_kw=mc_kw)
    ms = time_device(torch, dist, world, wl, W, K)
    wl.lw.check()
    wl.sw.check()
    launches = wl.launches() * K
    e2e_ms = time_host(torch, dist, world, wl, W, K)
    h2d, d2h = wl.transfer_bytes()
    wl.close()
    del wl, lw2, sw2
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_extras:
        # last round's series: the configs[1] grid per GPU, weak scaling
        st, sts = clear_sky_states(NCOL, NLAY, 20260925 + rank)
        wlw = Workload(st, sts, NLAY, local, world)
        wms = time_device(torch, dist, world, wlw, W, K)
        wes = time_host(torch, dist, world, wlw, W, K)
        wscan = wlw.lw.zero_scan_state
        wlw.close()
        del wlw
        torch.cuda.empty_cache()
        wms, wes = max_over_ranks(torch, dist, world, [wms, wes])
        extras["weak_clear_sky"] = {"workload": WORKLOAD + ", per GPU", "scaling": "weak", "value": world * NCOL * K / (wms * 1e-3),
                                    "unit": "columns/s", "ms_per_step": wms / K,
                                    "e2e": {"value": world * NCOL * K / (wes * 1e-3), "unit": "columns/s",
                                            "zero_input_scan_rank0": wscan}}
    ms, e2e_ms = max_over_ranks(torch, dist, world, [ms, e2e_ms])
    if not args.no_extras:
        # the strong-scaling denominator in the same run: rank 0 alone on the whole grid (the other ranks wait at the barrier)
        if rank == 0:
            blocks = [mcica_block_states(b) for b in range(nblk)]
            lwa, swa = concat_columns([b[0] for b in blocks]), concat_columns([b[1] for b in blocks])
            del blocks
            wl1 = Workload(lwa, swa, NLAY2, local, 1, engine_kw=mc_kw, host=False)
            K1 = max(3, K // 4)
            ms1 = time_device(torch, None, 1, wl1, 2, K1)
            extras["n1_same_workload"] = {"value": NCOL2 * K1 / (ms1 * 1e-3), "unit": "columns/s", "ms_per_step": ms1 / K1,
                                          "how": "rank 0 alone on the whole 131072-column grid, inputs resident in HBM, measured after "
                                                 "the timed region of this run (the other ranks idle)"}
            wl1.close()
        dist.barrier()
    if rank == 0:
        line = {
            "metric": METRIC2, "value": NCOL2 * K / (ms * 1e-3), "unit": "columns/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD2, "columns_global": NCOL2, "columns_per_gpu": NCOL2 // world, "levels": NLAY2,
                       "gpoints": "140 LW + 112 SW", "rng": "kissvec (per-column seeds, generated on the device), permuteseed 112",
                       "cache": "the engines' per-step workspace traffic exceeds the 126 MB L2: nothing survives between timed iterations",
                       "parallelism": f"fixed global grid cut into {world} contiguous column blocks (climt_b200/sharding.py), no "
                                      "data-path collective; one all_gather_into_tensor per step reassembles the 12 global flux / "
                                      "heating-rate fields on every rank, overlapped with the next step's kernels"},
            "e2e": {"value": NCOL2 * K / (e2e_ms * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world,
                    "note": "every rank's host-pointer C-ABI calls on its pinned block (H2D + kernels + D2H), barrier at the end; the "
                            "results stay sharded in host memory (a sharded consumer; SURVEY.md 8e). Bytes summed over ranks"},
            "gpu_launches": launches * world,
            "clocks": sampler_summary(sampler),
        }
        line.update(extras)
        print(json.dumps(line))
    dist.destroy_process_group()


def sampler_summary(sampler):
    sampler.stop.set()
    sampler.join(timeout=2)
    return sampler.summary()


if __name__ == "__main__":
    main()
