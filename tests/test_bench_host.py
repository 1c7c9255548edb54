"""Host-side pieces of bench.py that can be checked without a GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_table_matches_baseline_configs():
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        cfgs = json.load(fh)["configs"]
    assert len(cfgs) == 5
    assert bench.WORKLOADS["c1"][:4] == (10000, 100, 5000, 10)
    assert bench.WORKLOADS["c2"][:4] == (240000, 24000, 3072, 1)
    assert bench.WORKLOADS["c3"][:4] == (300000, 30000, 3072, 1)
    assert bench.WORKLOADS["c4"][:4] == (50000, 50000, 2048, 4)          # k = 3 + self
    assert bench.WORKLOADS["c5"][:4] == (1000000, 30000, 49152, 10)
    n5, q5, d5, k5, _ = bench.WORKLOADS["c5"]
    assert bench.WORKLOADS["c5s"][:4] == (n5 // 8, q5, d5, k5)           # one rank's share at 8 GPUs


def test_image_like_rows_do_not_depend_on_the_sharding():
    """configs[4] rows are generated per rank; every sharding must see the same matrix."""
    d, n = 48, 10000
    whole = bench.synth_rows("c5", 0, n, d, torch.device("cpu"), 1000)
    assert whole.dtype == torch.float32 and whole.shape == (n, d)
    assert float(whole.abs().max()) <= 1.0
    for world in (2, 3, 8):
        per = (n + world - 1) // world
        parts = [bench.synth_rows("c5", min(n, per * r), min(n, per * (r + 1)), d, torch.device("cpu"), 1000) for r in range(world)]
        assert torch.equal(torch.cat(parts), whole)
    other = bench.synth_rows("c5", 0, 64, d, torch.device("cpu"), 500000)      # queries: another seed base
    assert not torch.equal(other, whole[:64])


def test_reference_arm_skips_non_zero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "small", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
