"""The table-slot shortcut of the device code (climt_b200/csrc/cb_common.h: tbl_slot): int(1e4 * x / (bpade + x) + 0.5) evaluated with a
quotient that may be 1 ulp off the correctly rounded one, re-done with the IEEE division whenever the argument of the truncation
lies within 1e-9 of an integer.  Here the premise is checked in numpy: for quotients perturbed by one ulp in either direction the
shortcut either falls back or lands in the reference's slot -- never in a neighbouring one."""
import numpy as np


def _slot(q):
    v = np.float64(10000.0) * q + np.float64(0.5)     # two roundings, as __dadd_rn(__dmul_rn(.)) on the device
    return v, v.astype(np.int64)


def test_perturbed_quotient_never_changes_the_slot_without_falling_back():
    rng = np.random.default_rng(12)
    bpade = 1.0 / 0.278
    # optical depths as the kernels see them, plus arguments constructed to sit on slot boundaries
    x = np.concatenate([10.0 ** rng.uniform(np.log10(0.06), np.log10(600.0), 400_000),
                        rng.uniform(0.06, 5.0, 200_000)])
    k = rng.integers(100, 9999, 200_000).astype(np.float64)
    t = (k + 0.5 - 1e-13 * rng.integers(-3, 4, k.size)) / 10000.0     # quotients a hair away from a boundary
    x = np.concatenate([x, t * bpade / (1.0 - t)])
    q = x / (bpade + x)
    _, exact = _slot(q)
    fell_back = 0
    for q1 in (np.nextafter(q, 0.0), q, np.nextafter(q, 2.0)):
        v, fast = _slot(q1)
        near = np.abs(v - np.rint(v)) < 1e-9
        fell_back += int(near.sum())
        assert np.array_equal(fast[~near], exact[~near])
    assert 0 < fell_back < 3 * 0.3 * x.size      # the constructed boundary cases do fall back; ordinary arguments do not
    v, _ = _slot(q[:600_000])
    assert (np.abs(v - np.rint(v)) < 1e-9).mean() < 1e-6
