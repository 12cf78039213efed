"""Development canary (GPU box): small, ragged column counts through both engines against the oracle -- run first, under a
short timeout, when a kernel with block-wide barriers has changed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from climt_b200 import synthetic as SY
from climt_b200.engine import LWEngine, SWEngine
for ncol in (32, 300, 1, 257):
    st = SY.make_lw_state(ncol, 60, seed=5, clouds=True, aerosol=True)
    e = LWEngine(); got = e.run_host(ncol, 60, H.to_abi(st)); e.close()
    ref = H.run_lw_oracle(H.lw_oracle(cloud_overlap=1), st)
    print("lw", ncol, max(H.rel_err(got[k], ref[k]) for k in ("uflx", "dflx", "uflxc", "dflxc")), flush=True)
    sts = SY.make_sw_state(ncol, 60, seed=5, clouds=True)
    e = SWEngine(); gots = e.run_host(ncol, 60, H.to_abi_sw(sts), dyofyr=80); e.close()
    refs = H.sw_oracle()(sts, dyofyr=80)
    print("sw", ncol, max(H.rel_err(gots[k], refs[kk]) for k, kk in (("uflx", "swuflx"), ("dflx", "swdflx"), ("uflxc", "swuflxc"))), flush=True)
