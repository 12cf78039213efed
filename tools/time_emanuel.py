#!/usr/bin/env python
"""Development tool: time the Emanuel convection engine on a synthetic grid (device-resident inputs, both layouts, host path)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import helpers as H  # noqa: E402
from climt_b200 import emanuel, synthetic as SY  # noqa: E402

ncol, nlev, dt = int(os.environ.get("NCOL", 64800)), int(os.environ.get("NLAY", 60)), float(os.environ.get("DT", 1200))
st = SY.make_emanuel_state(ncol, nlev, seed=11)
par = dict(minorig=1, elcrit=0.0011, tlcrit=-55.0, entp=1.5, sigd=0.05, sigs=0.12, omtrain=50.0, omtsnow=5.5, coeffr=1.0, coeffs=0.8,
           cu=0.7, beta=10.0, dtmax=0.9, alpha=0.1, damp=0.1, cpd=1004.64, cpv=1846.0, cl=2500.0, rv=461.5, rd=287.0, lv0=2.5e6,
           g=9.80665, rowl=1e3, delt0=300.0, t_rain=273.0)
eng = emanuel.EmanuelEngine(par)
arrays = H.emanuel_arrays(st)
res = {"ncol": ncol, "nlev": nlev, "so": os.environ.get("CLIMT_B200_SO", "default")}
for layout in (0, 1):
    ins, outs = eng.shapes(ncol, nlev, layout)
    tens = {k: torch.from_numpy(np.ascontiguousarray(v if layout == 1 or v.ndim == 1 else v.T)).cuda() for k, v in arrays.items()}
    out = {k: torch.empty(outs[k], dtype=torch.int32 if k == "iflag" else torch.float64, device="cuda") for k in outs}
    fn = lambda: eng.run_device(ncol, nlev, tens, out, dt, qs_mode=1, layout=layout)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[f"layout{layout}_step_ms"] = e0.elapsed_time(e1) / 5
    eng.enable_timing(True)
    fn()
    torch.cuda.synchronize()
    eng.enable_timing(False)
    res[f"layout{layout}_kernel_ms"] = eng.last_kernel_ms
    res[f"layout{layout}_col_per_s"] = ncol / (res[f"layout{layout}_step_ms"] * 1e-3)
    fl = out["iflag"].cpu().numpy()
    res["flags"] = {int(f): int((fl == f).sum()) for f in np.unique(fl)}
    res[f"layout{layout}_checksum"] = float(out["ft"].abs().sum().item())
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory().numpy()  # noqa: E731
h_in = {k: pin(v) for k, v in arrays.items()}
ho = None
for i in range(4):
    if i == 1:
        t0 = time.perf_counter()
    ho = eng.run_host(h_in, dt, qs_mode=1, out=ho)
res["host_step_ms"] = (time.perf_counter() - t0) / 3 * 1e3
res["host_col_per_s"] = ncol / (res["host_step_ms"] * 1e-3)
print(json.dumps(res))
