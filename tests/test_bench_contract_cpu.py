"""bench.py, the parts that need no GPU: the reference arm's JSON line carries the keys the contract asks for (on the same
config / metric / unit as the CUDA arm), and the configs[2] grid is assembled from fixed-seed blocks whatever the rank count."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _ref(gpus):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus), "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_line_n1():
    import bench
    d = _ref(1)
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "columns/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == bench.WORKLOAD and d["config"]["same_config"] is True
    assert d["config"]["columns_per_step"] == bench.NCOL            # the whole grid, on every box
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and "-O3" in cb["flags"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 100.0


def test_reference_arm_follows_the_multi_gpu_workload():
    import bench
    d = _ref(2)
    assert d["metric"] == bench.METRIC2 and d["config"]["workload"] == bench.WORKLOAD2 and d["scaling"] == "strong"
    assert d["config"]["levels"] == 72 and d["n_gpus"] == 2


def test_config2_grid_is_independent_of_the_rank_count():
    import bench
    from climt_b200.sharding import shard_bounds
    assert bench.NCOL2 == 512 * 256 and bench.NCOL2 % bench.BLOCK2 == 0
    for world in (1, 2, 4, 8):
        b = shard_bounds(bench.NCOL2, world)
        assert all(lo % bench.BLOCK2 == 0 and hi % bench.BLOCK2 == 0 for lo, hi in b)
    a = bench.mcica_block_states(3)[0]
    b = bench.mcica_block_states(3)[0]
    assert all(np.array_equal(a[k], b[k]) for k in a)              # a block is a pure function of its index
    lw = bench.concat_columns([{"x": np.zeros((4, bench.BLOCK2)), "t": np.zeros((4, bench.BLOCK2, 16))},
                               {"x": np.ones((4, bench.BLOCK2)), "t": np.ones((4, bench.BLOCK2, 16))}])
    assert lw["x"].shape == (4, 2 * bench.BLOCK2) and lw["t"].shape == (4, 2 * bench.BLOCK2, 16)
    assert lw["x"][0, bench.BLOCK2] == 1.0 and lw["t"][0, bench.BLOCK2 - 1, 0] == 0.0
