#!/usr/bin/env python
"""Development tool: time the CORK LW / SW engines (device-resident inputs) on a synthetic grid."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import helpers as H  # noqa: E402
from climt_b200 import cork, synthetic as SY  # noqa: E402

ncol, nlev = int(os.environ.get("NCOL", 65536)), int(os.environ.get("NLAY", 60))
bands = os.environ.get("BANDS", "1") == "1"
rng = np.random.default_rng(3)
lw = SY.make_lw_state(ncol, nlev, seed=3)
s = {"T": lw["tlay"], "p": lw["play"] * 100.0, "p_int": lw["plev"] * 100.0, "T_surf": lw["tsfc"], "q": lw["h2o"] * 0.622,
     "co2": np.full((nlev, ncol), 4e-4), "emissivity": np.ones((14, ncol)), "tau_cloud_lw": np.zeros((nlev, ncol, 14)),
     "zenith": np.deg2rad(rng.uniform(0, 85, ncol)), "albedo": rng.uniform(0.05, 0.3, ncol),
     "tau_cloud_sw": np.zeros((nlev, ncol, 3)), "ssa_cloud": np.zeros((nlev, ncol, 3)), "g_cloud": np.zeros((nlev, ncol, 3))}
res = {"ncol": ncol, "nlev": nlev, "U": os.environ.get("CLIMT_B200_CORK_U", "default"), "so": os.environ.get("CLIMT_B200_SO", "default"),
       "bands": bands}
for which, tname in (("lw", "earth_low_res_lw"), ("sw", "earth_low_res_sw")):
    eng = cork.CorkEngine(tname)
    arrays = H.cork_arrays(s, which)
    ins, outs = eng.shapes(ncol, nlev)
    dev_in = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in arrays.items()}
    names = ["up_broad", "down_broad", "heating_rate"]
    if bands:
        names += ["up_band", "down_band", "tau_band", "hr_band"] + (["trans_band"] if which == "lw" else [])
    dev_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in names}
    fn = (lambda: eng.lw_device(ncol, nlev, dev_in, dev_out)) if which == "lw" else (lambda: eng.sw_device(ncol, nlev, dev_in, dev_out))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[which + "_step_ms"] = e0.elapsed_time(e1) / 5
    eng.enable_timing(True)
    fn()
    torch.cuda.synchronize()
    res[which + "_units_ms"] = eng.last_unit_kernel_ms
    res[which + "_col_per_s"] = ncol / (res[which + "_step_ms"] * 1e-3)
    res[which + "_checksum"] = float(dev_out["up_broad"].sum().item())
    eng.close()
print(json.dumps(res))
