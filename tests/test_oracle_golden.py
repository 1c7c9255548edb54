"""The oracle against the golden vectors generated from the UNMODIFIED reference (tests/golden/make_golden.py).

The reference ships no tests for this path, so the pins are outputs of the compiled reference
itself: in exhaustive mode Prioritized DCI retrieves every point and is exact."""
import glob
import os

import numpy as np
import pytest

from oracle import knn_oracle as ko
from oracle import ref_dci

KNN_CASES = ["knn_gauss", "knn_lowrank", "knn_ties", "knn_k_gt_n", "knn_image_odd_dim"]


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return z["data"], z["query"], int(z["k"]), z["ref_idx"], z["ref_dist"]


@pytest.mark.parametrize("name", KNN_CASES)
def test_c_oracle_matches_reference_exhaustive(golden_dir, name):
    x, q, k, ri, rd = load(golden_dir, name)
    idx, dist = ko.exact_knn_c(x, q, k)
    assert idx.shape == ri.shape == (q.shape[0], min(k, x.shape[0]))
    # same loop as util.c:62-69 -> distances bit-identical
    assert np.array_equal(dist, rd)
    if name == "knn_ties":
        # exact duplicates: the reference's order among equal distances is unspecified; distances pin the answer
        ok, msg = ko.compare_knn(idx, dist, ri, rd, x, q)
        assert ok, msg
        same_d = (dist[:, :-1] == dist[:, 1:])
        assert np.all(idx[:, :-1][same_d] < idx[:, 1:][same_d]), "oracle resolves ties to the lower index"
    else:
        assert np.array_equal(idx, ri)


@pytest.mark.parametrize("name", KNN_CASES)
def test_numpy_oracle_matches_reference_exhaustive(golden_dir, name):
    x, q, k, ri, rd = load(golden_dir, name)
    idx, dist = ko.exact_knn_numpy(x, q, k, xblock=97)
    ok, msg = ko.compare_knn(idx, dist, ri, rd, x, q, tie_rtol=1e-12, dist_rtol=1e-12)
    assert ok, msg


def test_squared_and_float32_inputs():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((900, 33)).astype(np.float32)
    q = rng.standard_normal((21, 33)).astype(np.float32)
    i1, d1 = ko.exact_knn_c(x, q, 6)
    i2, d2 = ko.exact_knn_numpy(x, q, 6, squared=True)
    assert np.array_equal(i1, i2)
    np.testing.assert_allclose(d1 ** 2, d2, rtol=1e-12)


def test_compare_knn_rules():
    ri = np.array([[3, 7]], dtype=np.int32)
    rd = np.array([[1.0, 2.0]])
    assert ko.compare_knn(ri, rd, ri, rd)[0]
    assert ko.compare_knn(ri, rd * (1 + 5e-6), ri, rd)[0]                         # distance within 1e-5
    assert not ko.compare_knn(ri, rd * (1 + 5e-5), ri, rd)[0]
    # with the inputs at hand a wrong index is re-evaluated: claimed distances do not excuse it
    x = np.array([[1.0], [2.0]]); q = np.array([[0.0]])
    gi = np.array([[0, 1]], dtype=np.int32); gd = np.array([[1.0, 2.0]])
    swapped = np.array([[1, 0]], dtype=np.int32)
    assert not ko.compare_knn(swapped, gd, gi, gd, x, q)[0]
    x = np.array([[1.0], [1.0 + 5e-7]])
    tie_d = np.array([[1.0, 1.0 + 5e-7]])
    assert ko.compare_knn(swapped, tie_d, gi, tie_d, x, q)[0]                     # tie within 1e-6: excused
    assert not ko.compare_knn(np.zeros((1, 1), np.int32), np.zeros((1, 1)), ri, rd)[0]   # shape


def test_merge_restatement_equals_global_answer():
    rng = np.random.default_rng(11)
    x = rng.standard_normal((1003, 24))
    q = rng.standard_normal((50, 24))
    gi, gd = ko.exact_knn_c(x, q, 5)
    from inclusivegan_b200.sharding import shard_range
    parts_i, parts_d = [], []
    for r in range(4):
        a, b = shard_range(1003, 4, r)
        i, d = ko.exact_knn_c(x[a:b], q, 5)
        parts_i.append(i + a)
        parts_d.append(d)
    mi, md = ko.merge_topk_numpy(np.stack(parts_i), np.stack(parts_d))
    assert np.array_equal(mi, gi) and np.array_equal(md, gd)


@pytest.mark.skipif(not ref_dci.available(), reason="oracle/_ref/_dci.so not built")
def test_prebuilt_reference_reproduces_deterministic_approximate_run(golden_dir):
    """The harness bench.py uses for the CPU baseline (oracle/ref_dci.py over the prebuilt extension)
    reproduces a deterministic approximate reference run recorded through the reference's own dci.py."""
    z = np.load(os.path.join(golden_dir, "approx_levels1.npz"))
    db = ref_dci.RefDCI(z["data"].shape[1], 2, 7)
    db.proj_vec[...] = z["proj_vec"]
    data = np.ascontiguousarray(z["data"])
    db.add(data, num_levels=1, prop_to_retrieve=0.05)
    fi, fd, cnt = db.query(z["query"], int(z["k"]), prop_to_retrieve=0.05)
    assert np.array_equal(cnt, z["counts"])
    assert np.array_equal(fi, z["flat_idx"])
    assert np.array_equal(fd, z["flat_dist"])
    # and its reported distances are exact for the points it did return
    qrow = np.repeat(np.arange(len(cnt)), cnt)
    np.testing.assert_array_equal(ko.pair_dist(data, z["query"], qrow, fi), fd)
    db.clear()


@pytest.mark.skipif(not ref_dci.available(), reason="oracle/_ref/_dci.so not built")
def test_reference_exhaustive_live(golden_dir):
    """Live (not recorded) run of the compiled reference in exhaustive mode equals the oracle."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((700, 48))
    q = rng.standard_normal((19, 48))
    db = ref_dci.RefDCI(48, 2, 7)
    db.add(x, num_levels=1, prop_to_retrieve=1.0, prop_to_visit=1.0)
    fi, fd, cnt = db.query(q, 3, prop_to_retrieve=1.0, prop_to_visit=1.0)
    oi, od = ko.exact_knn_c(x, q, 3)
    assert np.array_equal(fi.reshape(19, 3), oi) and np.array_equal(fd.reshape(19, 3), od)
    db.clear()
