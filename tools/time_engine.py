#!/usr/bin/env python
"""Development tool: time the LW / SW engines of whatever libclimt_b200 build $CLIMT_B200_SO points at."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import helpers as H  # noqa: E402
from climt_b200 import synthetic as SY  # noqa: E402
from climt_b200.engine import LWEngine, SWEngine, LW_IN, LW_OUT, SW_IN, lw_shapes  # noqa: E402

ncol, nlay = int(os.environ.get("NCOL", 8192)), int(os.environ.get("NLAY", 60))
clouds = os.environ.get("CLOUDS", "0") == "1"
mcica = os.environ.get("MCICA", "0") == "1"  # BASELINE configs[2]: McICA clouds, kissvec RNG (per-column seeds, generated on the device)
st = SY.make_lw_state(ncol, nlay, clouds=clouds or mcica)
sts = SY.make_sw_state(ncol, nlay, clouds=clouds or mcica, overcast_only=not mcica)
abi, abis = H.to_abi(st), H.to_abi_sw(sts)
_, outs = lw_shapes(ncol, nlay)
kw = dict(icld=2, mcica=True, irng=0, permuteseed=112) if mcica else {}
eng, engs = LWEngine(**kw), SWEngine(**kw)
d_in = {k: torch.from_numpy(abi[k]).cuda() for k in LW_IN}
ds_in = {k: torch.from_numpy(abis[k]).cuda() for k in SW_IN}
d_out = {k: torch.empty(outs[k], dtype=torch.float64, device="cuda") for k in LW_OUT}
res = {"so": os.environ.get("CLIMT_B200_SO", "default"), "ncol": ncol, "nlay": nlay, "clouds": clouds, "mcica": mcica}
for name, fn in (("lw", lambda: eng.run_device(ncol, nlay, d_in, d_out)),
                 ("sw", lambda: engs.run_device(ncol, nlay, ds_in, d_out, dyofyr=1))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[name + "_step_ms"] = e0.elapsed_time(e1) / 10
    e = eng if name == "lw" else engs
    e.enable_timing(True)
    ms, tms = [], []
    for _ in range(5):
        fn()
        torch.cuda.synchronize()
        ms.append(e.last_unit_kernel_ms)
        tms.append(e.last_taumol_kernel_ms)
    e.enable_timing(False)
    res[name + "_units_ms"] = float(np.mean(ms))
    res[name + "_taumol_ms"] = float(np.mean(tms))
res["lw_checksum"] = float(d_out["dflx"].sum().item())
print(json.dumps(res))
