#!/usr/bin/env python
"""Benchmarks of the other BASELINE.json configurations (bench.py keeps the headline contract: configs[1], RRTMG LW+SW).

  python bench_extra.py --workload mcica   # configs[2] per-GPU share: RRTMG LW+SW with McICA clouds (KISS RNG), 16384 col x 72 lev
  python bench_extra.py --workload cork    # configs[3] per-GPU share: CORK correlated-k LW+SW, 65536 col x 60 lev
  python bench_extra.py --workload gmd     # configs[4] per-GPU share: the GMD radiative-convective physics step (Instellation, RRTMG LW+SW, Emanuel, SimplePhysics, SlabSurface), 8100 col x 60 lev
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
    ap.add_argument("--workload", choices=["mcica", "cork", "gmd"], required=True)
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
    elif args.workload == "gmd":
        # One radiative-convective physics step of the aquaplanet configuration: Instellation (zenith angle) -> RRTMG LW + SW (clear
        # sky, as bench.py) -> SlabSurface (surface energy balance from the engines' surface fluxes), and Emanuel convection on the
        # same columns.  Device leg: one state resident in HBM in the (level, column) layout; the shortwave engine reads the cosine
        # Instellation wrote, the convection engine reads the radiation engines' temperature / pressure tensors in place (layout 0),
        # the slab reads row 0 of the flux outputs in place.
        from climt_b200.engine import LWEngine, SWEngine, LW_IN, LW_OUT, SW_IN, lw_shapes
        import datetime
        from climt_b200 import emanuel, instellation as INST, simple_physics as SP, slab_surface as SLAB
        ncol, nlay, dt_conv = args.ncol or 8100, 60, 1200.0
        when = datetime.datetime(2026, 3, 20, 12, 0)
        jc = INST.julian_centuries(when)
        sfc = SY.make_surface_state(ncol, seed=20260925 + rank)
        st = SY.make_lw_state(ncol, nlay, seed=20260925 + rank)
        sts = SY.make_sw_state(ncol, nlay, seed=20260925 + rank)
        es = SY.make_emanuel_state(ncol, nlay, seed=20260925 + rank)
        abi, abis = H.to_abi(st), H.to_abi_sw(sts)
        # the convection state on the radiation grid: same pressures, the convecting soundings' temperature and humidity
        from climt_b200 import state as S
        st["tlay"] = np.ascontiguousarray(es["air_temperature"].T)
        st["tlev"] = np.ascontiguousarray(S.get_interface_values(st["tlay"], st["tsfc"], st["play"], st["plev"]))
        sts = dict(sts, tlay=st["tlay"], tlev=st["tlev"])
        abi, abis = H.to_abi(st), H.to_abi_sw(sts)
        lw, sw = LWEngine(device=local), SWEngine(device=local)
        epar = dict(minorig=1, elcrit=0.0011, tlcrit=-55.0, entp=1.5, sigd=0.05, sigs=0.12, omtrain=50.0, omtsnow=5.5, coeffr=1.0,
                    coeffs=0.8, cu=0.7, beta=10.0, dtmax=0.9, alpha=0.1, damp=0.1, cpd=1004.64, cpv=1846.0, cl=2500.0, rv=461.5, rd=287.0,
                    lv0=2.5e6, g=9.80665, rowl=1e3, delt0=300.0, t_rain=273.0)
        em = emanuel.EmanuelEngine(epar, device=local)
        _, outs = lw_shapes(ncol, nlay)
        d_in = {k: torch.from_numpy(abi[k]).cuda() for k in LW_IN}
        ds_in = {k: (d_in[k] if k in d_in and k in ("play", "plev", "tlay", "tlev", "tsfc") else torch.from_numpy(abis[k]).cuda()) for k in SW_IN}
        d_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in LW_OUT}
        ds_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in LW_OUT}
        tr = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda()  # noqa: E731
        de_in = {"t": d_in["tlay"], "p": d_in["play"], "ph": d_in["plev"], "q": tr(es["specific_humidity"]), "u": tr(es["eastward_wind"]),
                 "v": tr(es["northward_wind"]), "cbmf": torch.from_numpy(es["cloud_base_mass_flux"]).cuda()}
        _, eouts = em.shapes(ncol, nlay, 0)
        de_out = {k: torch.empty(eouts[k], dtype=torch.int32 if k == "iflag" else torch.float64, device="cuda") for k in eouts}
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=a.dtype)).pin_memory().numpy()  # noqa: E731
        h_in, hs_in = {k: pin(abi[k]) for k in LW_IN}, {k: pin(abis[k]) for k in SW_IN}
        h_out, hs_out = {k: pin(np.empty(outs[k])) for k in LW_OUT}, {k: pin(np.empty(outs[k])) for k in LW_OUT}
        # host leg of the convection call: the component's (column, level) arrays
        he_in = {"t": pin(np.ascontiguousarray(st["tlay"].T)), "q": pin(es["specific_humidity"]), "u": pin(es["eastward_wind"]),
                 "v": pin(es["northward_wind"]), "p": pin(np.ascontiguousarray(st["play"].T)), "ph": pin(np.ascontiguousarray(st["plev"].T)),
                 "cbmf": pin(es["cloud_base_mass_flux"])}
        _, eouts1 = em.shapes(ncol, nlay, 1)
        he_out = {k: pin(np.empty(eouts1[k], dtype=np.int32 if k == "iflag" else np.float64)) for k in eouts1}
        d_sfc = {k: torch.from_numpy(v).cuda() for k, v in sfc.items()}
        h_sfc = {k: pin(v) for k, v in sfc.items()}
        flux_names = {"downwelling_shortwave_flux_in_air": (1, "dflx"), "downwelling_longwave_flux_in_air": (0, "dflx"),
                      "upwelling_shortwave_flux_in_air": (1, "uflx"), "upwelling_longwave_flux_in_air": (0, "uflx")}
        d_slab = dict(d_sfc, **{k: (ds_out if w else d_out)[f] for k, (w, f) in flux_names.items()})
        slab_result = {}
        # SimplePhysics (condensation, surface fluxes, boundary layer) on the same columns: pressures in Pa, surface first
        sp_par = SP.SimplePhysics(device=local).params()
        hsp_in = {"t": h_in["tlay"], "q": pin(np.ascontiguousarray(es["specific_humidity"].T)), "u": pin(np.ascontiguousarray(es["eastward_wind"].T)),
                  "v": pin(np.ascontiguousarray(es["northward_wind"].T)), "pmid": pin(st["play"] * 100.0), "pint": pin(st["plev"] * 100.0),
                  "ps": pin(st["plev"][0] * 100.0), "ts": pin(st["tsfc"].copy()), "qsurf": pin(np.zeros(ncol)), "lat": h_sfc["latitude"]}
        dsp_in = {k: (d_in["tlay"] if k == "t" else (de_in[k] if k in ("q", "u", "v") else torch.from_numpy(v).cuda())) for k, v in hsp_in.items()}

        def step_device():
            _, ds_in["coszen"] = INST.instellation_device(d_sfc["latitude"], d_sfc["longitude"], jc, want_coszen=True)
            lw.run_device(ncol, nlay, d_in, d_out)
            sw.run_device(ncol, nlay, ds_in, ds_out, dyofyr=1)
            em.run_device(ncol, nlay, de_in, de_out, dt_conv, qs_mode=emanuel.QS_BOLTON, layout=0)
            o = SP.simple_physics_device(sp_par, dsp_in, dt_conv)
            d_slab["surface_upward_latent_heat_flux"], d_slab["surface_upward_sensible_heat_flux"] = o["lat_ht_flux"], o["sens_ht_flux"]
            slab_result["device"] = SLAB.slab_surface_device(d_slab, flux_layout="level_major")

        def step_host():
            hs_in["coszen"][:] = np.cos(INST.instellation_host(h_sfc["latitude"], h_sfc["longitude"], jc, local))
            lw.run_host(ncol, nlay, h_in, h_out, wait=False)
            sw.run_host(ncol, nlay, hs_in, hs_out, dyofyr=1, wait=False)
            em.run_host(he_in, dt_conv, qs_mode=emanuel.QS_BOLTON, out=he_out)
            o = SP.simple_physics_host(sp_par, hsp_in, dt_conv, device=local)
            lw.wait()
            sw.wait()
            # the engines' host outputs are (interface_levels, column): row 0 is the surface value of every column
            slab_result["host"] = SLAB.slab_surface_host(dict(h_sfc, surface_upward_latent_heat_flux=o["lat_ht_flux"], surface_upward_sensible_heat_flux=o["sens_ht_flux"],
                                                              **{k: (hs_out if w else h_out)[f][0] for k, (w, f) in flux_names.items()}), local)
        name = "RRTMG LW+SW + Emanuel convection columns/s (60 lev)"
        workload = (f"radiative-convective physics step (the components of examples/gmd_radiative_convective.py): Instellation -> RRTMG LW+SW clear sky, "
                    f"Emanuel convection, SimplePhysics -> SlabSurface (dt 1200 s), "
                    f"{ncol} columns x 60 levels per GPU (BASELINE.json configs[4] is 360 x 180 = 64800 columns on 8 GPUs)")
        launches = lambda: lw.last_launches + sw.last_launches + em.last_launches + 3  # noqa: E731  (+ k_instellation, k_simple_physics, k_slab_surface)
        e_h2d = sum(v.nbytes for v in he_in.values()) + 2 * 8 * ncol + (15 * 8 + 4) * ncol + sum(v.nbytes for v in hsp_in.values())
        e_d2h = sum(v.nbytes for v in he_out.values()) + 8 * ncol + 2 * 8 * ncol + (4 * nlay + 3) * 8 * ncol
        xfer = lambda: tuple(a + b + c for a, b, c in zip(lw.last_transfer_bytes, sw.last_transfer_bytes, (e_h2d, e_d2h)))  # noqa: E731

        def cpu():
            from oracle import emanuel as OE, adjacent as OA, simple_physics as OSP
            from climt_b200.constants import DEFAULTS as CDEF
            csp = {k: v[0] for k, v in CDEF.items()}
            n = 128
            sub = {k: v[:, :n] if v.ndim == 2 else (v[:, :n, :] if v.ndim == 3 and v.shape[-1] in (14, 16) else v[..., :n]) for k, v in st.items()}
            subs = {k: v[:, :n] if v.ndim == 2 else (v[:, :n, :] if v.ndim == 3 and v.shape[-1] in (14, 16) else v[..., :n]) for k, v in sts.items()}
            sube = {k: v[:n] for k, v in es.items()}
            sube["air_temperature"] = np.ascontiguousarray(st["tlay"].T)[:n]
            sube["air_pressure"] = np.ascontiguousarray(st["play"].T)[:n]
            sube["air_pressure_on_interface_levels"] = np.ascontiguousarray(st["plev"].T)[:n]
            olw, osw = H.lw_oracle(), H.sw_oracle()
            consts = {k: epar[k] for k in ("cpd", "cpv", "cl", "rv", "rd", "lv0", "g", "rowl")}
            t0, reps = time.perf_counter(), 0
            while time.perf_counter() - t0 < 10.0:
                subs["coszen"] = np.cos(OA.instellation(sfc["latitude"][:n], sfc["longitude"][:n], when))
                ol = H.run_lw_oracle(olw, sub)
                os_ = osw(subs, dyofyr=1)
                OE.fortran_component_call(sube, dt_conv, consts)
                OSP.component_call({"air_temperature": sub["tlay"], "specific_humidity": hsp_in["q"][:, :n], "eastward_wind": hsp_in["u"][:, :n],
                                    "northward_wind": hsp_in["v"][:, :n], "air_pressure": hsp_in["pmid"][:, :n],
                                    "air_pressure_on_interface_levels": hsp_in["pint"][:, :n], "surface_air_pressure": hsp_in["ps"][:n],
                                    "surface_temperature": hsp_in["ts"][:n], "surface_specific_humidity": hsp_in["qsurf"][:n],
                                    "latitude": sfc["latitude"][:n]}, dt_conv, csp)
                OA.slab_surface(dict({k: v[:n] for k, v in sfc.items()}, downwelling_shortwave_flux_in_air=os_["swdflx"][0],
                                     downwelling_longwave_flux_in_air=ol["dflx"][0], upwelling_shortwave_flux_in_air=os_["swuflx"][0],
                                     upwelling_longwave_flux_in_air=ol["uflx"][0]))
                reps += 1
            dt = time.perf_counter() - t0
            return reps * n / dt, f"{n} of the workload's {ncol} columns x {reps} repetitions in {dt:.1f} s (C++ restatements of RRTMG and CONVECT, numpy restatements of the small steps, 1 core)"
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
