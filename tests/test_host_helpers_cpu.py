"""Host-side helpers of the host-pointer calls (climt_b200/csrc/engine_common.h), compiled for the CPU: the all-zero scan that lets
the engines replace the PCIe transfer of all-zero inputs by a device memset, the persistent worker pool under it, and the guard
that turns the scan off on hosts where it is slower than the copy it saves.  (The pipeline around them needs a GPU:
tests/test_host_pipeline_gpu.py.)"""
import ctypes

import numpy as np
import pytest

import helpers as H

_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def lib():
    L = H.host_pipe_emul_lib()
    L.emul_pool_stress.restype = ctypes.c_long
    return L


def _scan(L, arrays, ncol, c0, n):
    keep = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1, ncol) for a in arrays]
    base = (_dp * len(keep))(*[a.ctypes.data_as(_dp) for a in keep])
    rows = (ctypes.c_long * len(keep))(*[a.shape[0] for a in keep])
    zero = (ctypes.c_int * len(keep))()
    L.emul_all_zero(base, rows, len(keep), ctypes.c_long(ncol), ctypes.c_long(c0), ctypes.c_long(n), zero)
    return [bool(z) for z in zero]


def test_scan_sees_exactly_the_chunk_it_is_given(lib):
    ncol = 1000
    a = np.zeros((960, ncol))            # an aerosol-sized array (16 bands x 60 layers)
    b = np.zeros((60, ncol))
    c = np.zeros((1, ncol))
    assert _scan(lib, [a, b, c], ncol, 0, 256) == [True, True, True]
    a[959, 999] = 1e-300                 # last row, last column: only the last chunk is dirty
    b[0, 256] = -0.0                     # minus zero has a set bit: "not zero", the array is simply transferred
    c[0, 255] = np.nan
    assert _scan(lib, [a, b, c], ncol, 0, 256) == [True, True, False]
    assert _scan(lib, [a, b, c], ncol, 256, 256) == [True, False, True]
    assert _scan(lib, [a, b, c], ncol, 768, 232) == [False, True, True]
    assert _scan(lib, [a, b, c], ncol, 0, ncol) == [False, False, False]
    assert _scan(lib, [a], ncol, 999, 1) == [False] and _scan(lib, [a], ncol, 998, 1) == [True]


def test_scan_agrees_with_numpy_on_random_sparsity(lib):
    rng = np.random.default_rng(3)
    ncol = 777
    for trial in range(20):
        arrs = []
        for rows in (1, 7, 60, 640):
            a = np.zeros((rows, ncol))
            if rng.uniform() < 0.5:
                a[rng.integers(rows), rng.integers(ncol)] = rng.normal()
            arrs.append(a)
        c0 = int(rng.integers(0, ncol - 1))
        n = int(rng.integers(1, ncol - c0 + 1))
        want = [not a[:, c0:c0 + n].view(np.uint64).any() for a in arrs]
        assert _scan(lib, arrs, ncol, c0, n) == want


def test_pool_runs_every_task_once_under_two_concurrent_callers(lib):
    assert lib.emul_pool_size() >= 1
    assert lib.emul_pool_stress(200, 37) == 2 * 200 * 37
    assert lib.emul_pool_stress(50, 1) == 100


def test_guard_trips_after_three_slow_scans_in_a_row(lib):
    def trip(rates):
        r = np.array(rates, dtype=np.float64)
        return lib.emul_scan_guard(r.ctypes.data_as(_dp), len(r))
    assert trip([100, 80, 60, 120]) == -1
    assert trip([5, 5, 100, 5, 5, 100]) == -1          # never three in a row
    assert trip([100, 5, 6, 7, 100]) == 3
