#!/usr/bin/env python
"""Development study (CPU only): how much accuracy would a *parallel-scan* form of the shortwave adding method cost?

The transfer kernel carries 7 values per (g-point, layer) through HBM between its two vertical sweeps because the adding
recurrences are serial in the level.  They are, however, compositions of maps that associate:
  rupd' = refd + trad^2 rupd / (1 - rupd refd)        -- a Moebius map of rupd: 2x2 matrices compose by multiplication
  rup'  = ref + trad ((tra - dbt) rupd + dbt rup) / (1 - rupd refd)   -- affine in rup once rupd is known
so a warp could own one (column, g-point) with the levels across its lanes and scan.  This script takes the layer properties the
kernel code itself produces (host emulation, scratch rows exported by tests/emul/sw_emul.cpp) for cloudy synthetic columns, redoes
the upward sweep (a) serially as the kernel does and (b) as prefix products of normalised matrices in a balanced-tree order, and
reports the difference.  Result recorded in DESIGN.md 6a.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from climt_b200 import synthetic as SY  # noqa: E402


def tree_prefix(mats):
    """inclusive prefix products P_l = M_l ... M_1 M_0 in a balanced (Hillis-Steele) order, each product renormalised"""
    P = mats.copy()
    n, d = P.shape[0], 1
    while d < n:
        Q = P.copy()
        Q[d:] = np.einsum("lij,ljk->lik", P[d:], P[:-d])
        Q /= np.abs(Q).max(axis=(1, 2), keepdims=True)      # Moebius maps are projective: any rescaling is the same map
        P, d = Q, 2 * d
    return P


def main():
    ncol, nlay = 24, 60
    st = SY.make_sw_state(ncol, nlay, seed=31, clouds=True)
    lib = H.emul_lib("sw")
    nscr = lib.emul_sw_nscr()
    scr = np.zeros((112, nscr, nlay, ncol))
    lib.emul_sw_export_scratch(scr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    rc, out = H.run_sw_emul(st)
    lib.emul_sw_export_scratch(None)
    assert rc == 0
    worst = {"serial_vs_kernel": 0.0, "scan_rupd": 0.0, "scan_rup": 0.0}
    for g in range(0, 112, 3):
        for c in range(ncol):
            cloudy = scr[g, 7:14, :, c].any()
            rows = scr[g, 7:14, :, c] if cloudy else scr[g, 0:7, :, c]
            ref, refd, tra, trad, dbt, rup_k, rupd_k = rows
            if not np.isfinite(rows).all() or not ref.any():
                continue
            # the kernel's starting values are not exported: recover them from its first step is ill-posed, so restart the
            # recurrence from level 1 with the kernel's own rup / rupd of level 0
            r, rd = rup_k[0], rupd_k[0]
            rup_s, rupd_s = [r], [rd]
            for l in range(1, nlay):
                z = 1.0 / (1.0 - rd * refd[l])
                r, rd = ref[l] + trad[l] * ((tra[l] - dbt[l]) * rd + dbt[l] * r) * z, refd[l] + trad[l] * trad[l] * rd * z
                rup_s.append(r)
                rupd_s.append(rd)
            rup_s, rupd_s = np.array(rup_s), np.array(rupd_s)
            worst["serial_vs_kernel"] = max(worst["serial_vs_kernel"], float(np.abs(rupd_s - rupd_k).max()), float(np.abs(rup_s - rup_k).max()))
            # (b) scan: Moebius matrices of levels 1..nlay-1 applied to rupd_k[0]
            M = np.zeros((nlay - 1, 2, 2))
            M[:, 0, 0] = trad[1:] ** 2 - refd[1:] ** 2
            M[:, 0, 1] = refd[1:]
            M[:, 1, 0] = -refd[1:]
            M[:, 1, 1] = 1.0
            P = tree_prefix(M)
            rupd_p = np.concatenate([[rupd_k[0]], (P[:, 0, 0] * rupd_k[0] + P[:, 0, 1]) / (P[:, 1, 0] * rupd_k[0] + P[:, 1, 1])])
            worst["scan_rupd"] = max(worst["scan_rupd"], float(np.abs(rupd_p - rupd_s).max()))
            # rup: affine maps r' = a_l r + b_l with the scanned rupd of the level below
            z = 1.0 / (1.0 - rupd_p[:-1] * refd[1:])
            a = trad[1:] * dbt[1:] * z
            b = ref[1:] + trad[1:] * (tra[1:] - dbt[1:]) * rupd_p[:-1] * z
            A = np.zeros((nlay - 1, 2, 2))
            A[:, 0, 0], A[:, 0, 1], A[:, 1, 1] = a, b, 1.0
            PA, d = A.copy(), 1
            while d < nlay - 1:
                Q = PA.copy()
                Q[d:] = np.einsum("lij,ljk->lik", PA[d:], PA[:-d])
                PA, d = Q, 2 * d
            rup_p = np.concatenate([[rup_k[0]], PA[:, 0, 0] * rup_k[0] + PA[:, 0, 1]])
            worst["scan_rup"] = max(worst["scan_rup"], float(np.abs(rup_p - rup_s).max()))
    print({k: f"{v:.3e}" for k, v in worst.items()}, "(absolute; reflectances are O(1))")


if __name__ == "__main__":
    main()
