"""CPU-side checks of bench.py's host logic (no GPU): the reference arm's JSON line, the bounded-sample sizing, the
roofline traffic lookup and the strong-scaling layout helpers."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_sample_grows_with_the_time_budget_and_keeps_the_ratios():
    prev = 0
    for budget in (5.0, 14.0, 30.0, 100.0):
        t, n, nev, nex = bench.ref_sample_shape("c2", 16, 0, budget)
        assert t == "d" and n >= prev and n <= 20000
        assert nev * 20 == n and nex * 50 == n  # nev / N = 1/20, nex / N = 1/50 as in C2
        prev = n
    assert bench.ref_sample_shape("c2", 16, 0, 30.0)[1] == 8000
    assert bench.ref_sample_shape("c2", 16, 4321, 30.0)[1] == 4321  # --ref-n override
    assert bench.ref_sample_shape("c2", 4, 0, 30.0)[1] < 8000  # fewer cores, smaller sample


def test_traffic_lookup_has_every_gpu_count_of_the_scaling_run():
    for key in ("c2", "c2_g2", "c2_g4", "c2_g8"):
        v = bench.hemm_traffic(key)
        assert isinstance(v, int) and v > 5e8, key
    assert bench.hemm_traffic("no such workload") is None


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "chase_ref_cpu_d")),
                    reason="reference CPU solver not built (needs /root/reference)")
def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference`: the unmodified reference CPU solver on a bounded sample, one JSON line with the
    contract keys; non-zero ranks of a torchrun launch print nothing."""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--ref-n", "1500"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "TFLOP/s" and j["higher_is_better"] is True
    assert j["scaling"] == "strong" and j["steps"] == 1 and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "N=1500" in j["config"]["sample"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out1 = subprocess.run(cmd, capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def test_strong_scaling_blocks_tile_the_fixed_problem():
    """The multi-GPU arm holds C2 fixed: for every grid of the scaling run the local blocks (block-cyclic 64) cover each
    global row / column exactly once, and the weak arm's N is the round-1 sequence."""
    from chase_b200 import bench_dist as bd
    from chase_b200 import dist as cd

    assert bd.BASE == ("d", 20000, 1000, 400)
    assert [bd.weak_n(g) for g in (2, 4, 8)] == [28288, 40000, 56576]
    for G in (1, 2, 4, 8):
        r, c = cd.grid_dims(G)
        assert r * c == G and r >= c
        rows = np.concatenate([cd.global_indices(20000, r, 64, i) for i in range(r)])
        cols = np.concatenate([cd.global_indices(20000, c, 64, j) for j in range(c)])
        assert np.array_equal(np.sort(rows), np.arange(20000)) and np.array_equal(np.sort(cols), np.arange(20000))
