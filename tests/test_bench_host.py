"""Host-side pieces of bench.py that can be checked without a GPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_dci  # noqa: E402


def test_workload_table_matches_baseline_configs():
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        cfgs = json.load(fh)["configs"]
    assert len(cfgs) == 5
    assert bench.WORKLOADS["c1"][:4] == (10000, 100, 5000, 10)
    assert bench.WORKLOADS["c2"][:4] == (240000, 24000, 3072, 1)
    assert bench.WORKLOADS["c3"][:4] == (300000, 30000, 3072, 1)
    assert bench.WORKLOADS["c4"][:4] == (50000, 50000, 2048, 4)          # k = 3 + self
    assert bench.WORKLOADS["c5"][:4] == (1000000, 30000, 49152, 10)
    n5, q5, d5, k5 = bench.WORKLOADS["c5"][:4]
    assert bench.WORKLOADS["c5s"][:4] == (n5 // 8, q5, d5, k5)           # one rank's share at 8 GPUs
    # generators follow SURVEY.md 8d
    assert [bench.WORKLOADS[w][4] for w in ("c1", "c2", "c3", "c4", "c5")] == ["lowrank", "pixels", "gauss", "relu", "image"]
    for w in bench.WORKLOADS:
        cfg = bench.workload_config(w)
        assert set(cfg) == {"workload", "pool", "queries", "dim", "k", "features"}        # no model keys, same in both arms


@pytest.mark.parametrize("workload", ["c5", "c3", "c2", "c4", "c1"])
def test_rows_do_not_depend_on_the_sharding(workload):
    """Rows are generated per rank; every sharding must see the same matrix."""
    d, n = 48, 10000
    whole = bench.synth_rows(workload, 0, n, d, torch.device("cpu"), bench.POOL_SEED)
    want = torch.float64 if bench.GENERATORS[bench.WORKLOADS[workload][4]][0] == "float64" else torch.float32
    assert whole.dtype == want and whole.shape == (n, d)
    for world in (2, 3, 8):
        per = (n + world - 1) // world
        parts = [bench.synth_rows(workload, min(n, per * r), min(n, per * (r + 1)), d, torch.device("cpu"), bench.POOL_SEED) for r in range(world)]
        assert torch.equal(torch.cat(parts), whole)
    other = bench.synth_rows(workload, 0, 64, d, torch.device("cpu"), bench.QUERY_SEED)      # queries: another seed base
    assert not torch.equal(other, whole[:64])


def test_generators_follow_the_survey_distributions():
    d, n = 64, 8192
    dev = torch.device("cpu")
    g = bench.synth_rows("c3", 0, n, d, dev, 1).numpy()
    assert abs(g.mean()) < 0.02 and abs(g.std() - 1.0) < 0.02
    assert np.array_equal(g, g.astype(np.float32).astype(np.float64))            # float32 values widened (.astype(float64))
    p = bench.synth_rows("c2", 0, n, d, dev, 1).numpy()
    assert p.min() >= -1.0 and p.max() <= 1.0 and abs(p.std() - 0.5) < 0.05 and (np.abs(p) == 1.0).mean() > 0.01
    r = bench.synth_rows("c4", 0, n, d, dev, 1).numpy()
    assert r.dtype == np.float32 and r.min() == 0.0 and 0.45 < (r == 0).mean() < 0.55
    lo = bench.synth_rows("c1", 0, 512, 500, dev, 1).numpy()
    assert np.linalg.matrix_rank(lo) == 50                                       # dci_code/example.py:36-40: rank-50 data
    # the host generators (reference arm, NumPy) draw from the same distributions, per row block
    for wl in ("c3", "c2", "c4", "c1"):
        a = bench.synth_rows_host(wl, 0, 6000, 40, 7, threads=3)
        b = np.concatenate([bench.synth_rows_host(wl, 0, 2500, 40, 7, threads=1), bench.synth_rows_host(wl, 2500, 6000, 40, 7, threads=2)])
        assert a.dtype == np.float64 and np.array_equal(a, b)
        t = bench.synth_rows(wl, 0, 6000, 40, dev, 7).double().numpy()
        assert abs(a.mean() - t.mean()) < 0.05 * max(1.0, t.std()) and abs(a.std() - t.std()) < 0.05 * t.std()


def test_reference_arm_skips_non_zero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(not ref_dci.available(), reason="oracle/_ref/_dci.so not built")
def test_reference_arm_under_torchrun_environment():
    """torchrun exports OMP_NUM_THREADS=1 when nproc > 1 (round 1: the reference DCI then ran single-threaded while
    the line said otherwise).  The arm must use every core whatever it inherits, keep the requested step count,
    report the thread count in effect, score recall against the exact answer, and print the GPU arm's config."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "small",
                          "--steps", "3", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    cores = bench.host_cores()
    assert line["impl"] == "reference" and line["steps"] == 3 and line["warmup"] == 1
    assert line["cpu_baseline"]["threads"] == cores and line["cpu_baseline"]["cores"] == cores
    assert "threads in effect=%d" % cores in line["cpu_baseline"]["sample"]
    assert line["config"] == bench.workload_config("small")
    assert line["cpu_baseline"]["pool_rows"] == bench.WORKLOADS["small"][0]          # the full pool, not a subsample
    assert 0.0 < line["recall_at_k"] <= 1.0 and line["cpu_baseline"]["recall_at_k"] == line["recall_at_k"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["metric"] == bench.METRIC and line["higher_is_better"] is True and line["gpu_launches"] == 0


def test_b200_arm_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "small", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
