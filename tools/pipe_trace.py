#!/usr/bin/env python
"""Development tool: stage timeline of the host-pointer pipeline (CLIMT_B200_PIPE_TRACE=1), LW and SW calls overlapped as in
bench.py's e2e leg.  Prints, per engine and chunk, when each stage completed on the GPU and when it was enqueued on the host."""
import os
import sys

os.environ["CLIMT_B200_PIPE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import helpers as H  # noqa: E402
from climt_b200 import synthetic as SY  # noqa: E402
from climt_b200.engine import LWEngine, SWEngine, LW_IN, LW_OUT, SW_IN, lw_shapes  # noqa: E402

ncol, nlay = 8192, 60
abi, abis = H.to_abi(SY.make_lw_state(ncol, nlay)), H.to_abi_sw(SY.make_sw_state(ncol, nlay))
_, outs = lw_shapes(ncol, nlay)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
np_in, nps_in = {k: pin(abi[k]) for k in LW_IN}, {k: pin(abis[k]) for k in SW_IN}
np_out, nps_out = {k: pin(np.empty(outs[k])) for k in LW_OUT}, {k: pin(np.empty(outs[k])) for k in LW_OUT}
lw, sw = LWEngine(), SWEngine()
for rep in range(4):
    print(f"---- repetition {rep}", file=sys.stderr, flush=True)
    lw.run_host(ncol, nlay, np_in, np_out, wait=False)
    sw.run_host(ncol, nlay, nps_in, nps_out, dyofyr=1, wait=False)
    lw.wait()
    sw.wait()
