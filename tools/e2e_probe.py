#!/usr/bin/env python
"""Development tool: where does the host-pointer (e2e) path spend its time?  PCIe copy rates vs engine host calls."""
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
from climt_b200 import synthetic as SY  # noqa: E402
from climt_b200.engine import LWEngine, SWEngine, LW_IN, LW_OUT, SW_IN, lw_shapes  # noqa: E402

res = {}
# 1. raw PCIe rates, pinned memory
n = 25 * 1024 * 1024  # 200 MB of doubles
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    res[name + "_GBs"] = 5 * n * 8 / (time.perf_counter() - t0) / 1e9
# strided (column-chunk) copies like the pipeline's gathers: (rows, 8192) -> (rows, 4096)
rows = n // 8192
h2 = h[: rows * 8192].view(rows, 8192)
d2 = torch.empty((rows, 4096), dtype=torch.float64, device="cuda")
d2.copy_(h2[:, :4096], non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    d2.copy_(h2[:, :4096], non_blocking=True)
    d2.copy_(h2[:, 4096:], non_blocking=True)
torch.cuda.synchronize()
res["h2d_strided_GBs"] = 5 * rows * 8192 * 8 / (time.perf_counter() - t0) / 1e9

ncol, nlay = 8192, 60
abi, abis = H.to_abi(SY.make_lw_state(ncol, nlay)), H.to_abi_sw(SY.make_sw_state(ncol, nlay))
_, outs = lw_shapes(ncol, nlay)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
np_in = {k: pin(abi[k]) for k in LW_IN}
nps_in = {k: pin(abis[k]) for k in SW_IN}
np_out = {k: pin(np.empty(outs[k])) for k in LW_OUT}
nps_out = {k: pin(np.empty(outs[k])) for k in LW_OUT}
pg_in = {k: np.array(abi[k]) for k in LW_IN}       # pageable copies
pg_out = {k: np.empty(outs[k]) for k in LW_OUT}


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for chunk in (1024, 2048, 4096, 8192):
    os.environ["CLIMT_B200_HOST_CHUNK"] = str(chunk)
    lw, sw = LWEngine(), SWEngine()
    r = {}
    r["lw_ms"] = timeit(lambda: lw.run_host(ncol, nlay, np_in, np_out))
    r["sw_ms"] = timeit(lambda: sw.run_host(ncol, nlay, nps_in, nps_out, dyofyr=1))

    def both():
        lw.run_host(ncol, nlay, np_in, np_out, wait=False)
        sw.run_host(ncol, nlay, nps_in, nps_out, dyofyr=1, wait=False)
        lw.wait(); sw.wait()
    r["both_async_ms"] = timeit(both)
    if chunk == 4096:
        r["lw_pageable_ms"] = timeit(lambda: lw.run_host(ncol, nlay, pg_in, pg_out))
        # python-side overhead of one call (argument packing only): time with a tiny problem
        t0 = time.perf_counter()
        h2d, d2h = lw.last_transfer_bytes
        r["lw_bytes"] = [h2d, d2h]
    res[f"chunk{chunk}"] = r
    lw.close(); sw.close()
print(json.dumps(res))
