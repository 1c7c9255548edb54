"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI (via the DCI drop-in class and the
device-pointer wrapper), against the float64 oracle on the same seeded inputs.

Acceptance (BASELINE.json north_star): neighbour indices equal the oracle's except for ties within 1e-6
relative distance; distances within 1e-5 relative.  The tolerances live in oracle.knn_oracle.compare_knn.
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import knn_oracle as ko

pytestmark = pytest.mark.gpu

FLAG_SQUARED, FLAG_NO_CERTIFY, FLAG_FORCE_SCAN = 1, 2, 4
KNN_CASES = ["knn_gauss", "knn_lowrank", "knn_ties", "knn_k_gt_n", "knn_image_odd_dim"]


@pytest.fixture(scope="module")
def lib(native_lib):
    assert native_lib.b200knn_device_count() >= 1, "no sm_100 device visible: the CUDA path cannot run (no CPU fallback)"
    return native_lib


def make(kind, n, q, d, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    if kind == "gauss":
        x, y = rng.standard_normal((n, d)), rng.standard_normal((q, d))
    elif kind == "image":            # [-1, 1]-ranged pixels, training_loop.py:379
        x, y = np.clip(0.5 * rng.standard_normal((n, d)), -1, 1), np.clip(0.5 * rng.standard_normal((q, d)), -1, 1)
    elif kind == "cluster":          # queries next to pool rows: small NN gaps relative to norms
        x = rng.standard_normal((n, d))
        y = x[rng.integers(0, n, q)] + 0.05 * rng.standard_normal((q, d))
    elif kind == "relu":             # non-negative, Inception-pool-like (config 4)
        x, y = np.maximum(rng.standard_normal((n, d)), 0), np.maximum(rng.standard_normal((q, d)), 0)
    elif kind == "lowrank":          # dci_code/example.py:36-40
        lat = 2 * rng.random((n + q, 50)) - 1
        t = 2 * rng.random((50, d)) - 1
        z = lat @ t
        x, y = z[:n], z[n:]
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(x.astype(dtype)).copy(), np.ascontiguousarray(y.astype(dtype)).copy()


def check(db, x, y, k, flags=0, squared=False):
    idx, dist = db.query_arrays(y, k, squared=squared, flags=flags)
    ri, rd = ko.exact_knn_numpy(x, y, k, squared=squared)
    if squared:
        ok, msg = ko.compare_knn(idx, np.sqrt(dist), ri, np.sqrt(rd), x, y)
    else:
        ok, msg = ko.compare_knn(idx, dist, ri, rd, x, y)
    assert ok, msg
    return idx, dist


# ------------------------------------------------------------------------------------------------ goldens
@pytest.mark.parametrize("name", KNN_CASES)
def test_golden_vectors_through_dci_class(lib, golden_dir, name):
    """Same inputs the reference answered (exhaustive mode) -> same indices/distances, via DCI.add/query."""
    from inclusivegan_b200 import DCI
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    x, q, k = np.ascontiguousarray(z["data"]).copy(), z["query"], int(z["k"])
    db = DCI(x.shape[1], 2, 7)
    db.add(x, num_levels=2, field_of_view=10, prop_to_retrieve=0.002)
    idx, dist = db.query(q, num_neighbours=k, field_of_view=100, prop_to_retrieve=0.05)
    assert isinstance(idx, list) and len(idx) == q.shape[0] and idx[0].dtype == np.int32 and dist[0].dtype == np.float64
    ok, msg = ko.compare_knn(np.array(idx), np.array(dist), z["ref_idx"], z["ref_dist"], x, q)
    assert ok, msg
    if name != "knn_ties":
        assert np.array_equal(np.array(idx), z["ref_idx"])
    st = db.stats()
    assert st["kernel_launches"] > 0


# ------------------------------------------------------------------------------------------------ shapes
@pytest.mark.parametrize("kind,n,q,d,k,dtype", [
    ("gauss", 1000, 10, 64, 1, np.float64),          # one tile, one K block
    ("gauss", 3000, 130, 200, 1, np.float64),        # K tail (200 = 3*64 + 8), 2 query tiles
    ("gauss", 5000, 300, 129, 3, np.float64),        # dim % 8 != 0: padded BF16 pitch, scalar convert path
    ("gauss", 5000, 64, 1024, 10, np.float32),       # float32 extension, C = 32 shortlist
    ("image", 20000, 500, 3072, 1, np.float64),      # IMLE feature shape (32x32x3 pixels in [-1,1])
    ("cluster", 70000, 1000, 512, 1, np.float64),    # many shortlists per query (chunked sweep)
    ("relu", 6000, 6000, 256, 4, np.float32),        # config-4-like
    ("lowrank", 10000, 100, 5000, 10, np.float64),   # config 1 (dci_code/example.py data), full size
    ("gauss", 257, 129, 72, 16, np.float64),         # k = 16: last k served by the 32-entry shortlist
    ("cluster", 20000, 300, 512, 25, np.float64),    # k = num_samples_factor default of training_loop() (:154): 64-entry shortlist
    ("gauss", 500, 40, 64, 32, np.float64),          # k = 32 boundary of the tensor path
    ("gauss", 500, 40, 64, 33, np.float64),          # first k on the exact-scan path
    ("gauss", 300, 20, 64, 40, np.float64),          # k > 32: exact scan + segmented sort
    ("gauss", 1, 3, 8, 1, np.float64),               # single pool row
    ("image", 3000, 300, 49152, 10, np.float32),     # config-5 feature shape (128x128x3 raw pixels), k=10, reduced N
])
def test_parity_vs_oracle(lib, kind, n, q, d, k, dtype):
    from inclusivegan_b200 import DCI
    x, y = make(kind, n, q, d, seed=n + q + d, dtype=dtype)
    db = DCI(d, 3, 15)
    db.add(x)
    assert db.num_points == n
    check(db, x, y, k)


@pytest.mark.parametrize("flags", [FLAG_FORCE_SCAN, FLAG_NO_CERTIFY, 0])
def test_each_code_path_alone(lib, flags):
    """exact scan only / tensor pass without the certificate / full pipeline — all must be exact here."""
    from inclusivegan_b200 import DCI
    x, y = make("cluster", 30000, 700, 384, seed=9)
    db = DCI(384)
    db.add(x)
    check(db, x, y, 5, flags=flags)


def test_squared_distances(lib):
    from inclusivegan_b200 import DCI
    x, y = make("relu", 4000, 100, 2048, seed=4, dtype=np.float32)
    db = DCI(2048)
    db.add(x)
    check(db, x, y, 4, squared=True)


def test_query_self_equals_querying_the_pool_rows(lib):
    """b200knn_query_self (pool rows as queries, nothing re-uploaded or re-converted) == query(pool) bit for bit."""
    from inclusivegan_b200 import DCI
    x, _ = make("relu", 40000, 1, 384, seed=18, dtype=np.float32)         # > one device pass (32768 rows)
    db = DCI(384)
    db.add(x)
    i1, d1 = db.query_self_arrays(4, squared=True)
    i2, d2 = db.query_arrays(x, 4, squared=True)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    assert np.array_equal(i1[:, 0], np.arange(40000)) and np.all(d1[:, 0] == 0.0)


def test_self_knn_radius(lib):
    """precision_recall.py:74-90 pattern: k+1 smallest incl. self; self must be rank 0 at distance 0."""
    from inclusivegan_b200 import DCI
    x, _ = make("relu", 5000, 1, 512, seed=8, dtype=np.float32)
    db = DCI(512)
    db.add(x)
    idx, dist = db.query_arrays(x, 4, squared=True)
    assert np.array_equal(idx[:, 0], np.arange(5000)) and np.all(dist[:, 0] == 0.0)
    ri, rd = ko.exact_knn_numpy(x, x, 4, squared=True)
    np.testing.assert_allclose(dist[:, 3], rd[:, 3], rtol=1e-5)


# ------------------------------------------------------------------------------------------------ API behaviour
def test_trainer_call_pattern(lib):
    """training_loop.py:367-406: reset(); add(...); batched query(k=1) -> np.array(idx)[:,0]; second add is refused."""
    from inclusivegan_b200 import DCI
    x, y = make("image", 4000, 96, 768, seed=12)
    db = DCI(768, num_comp_indices=3, num_simp_indices=15)
    nearest, dists = [], []
    for round_ in range(2):                                   # two refreshes: reset must fully drop the old pool
        pool = x if round_ == 0 else np.ascontiguousarray(x[::-1]).copy()
        db.reset()
        db.add(pool, num_levels=3, field_of_view=10, prop_to_retrieve=0.002)
        assert db.num_points == 4000 and db.num_levels == 3
        with pytest.raises(RuntimeError, match="does not support insertion of more than one array"):
            db.add(pool)
        nearest, dists = [], []
        for s in range(0, 96, 24):                            # 2*minibatch rows per call (24 with README settings)
            i, dd = db.query(y[s:s + 24], num_neighbours=1, field_of_view=200, prop_to_retrieve=1.0)
            nearest += list(np.array(i)[:, 0])
            dists += list(np.array(dd)[:, 0])
        ri, rd = ko.exact_knn_c(pool, y, 1)
        ok, msg = ko.compare_knn(np.array(nearest)[:, None], np.array(dists)[:, None], ri, rd, pool, y)
        assert ok, msg
    db.clear()
    assert db.num_points == 0
    with pytest.raises(RuntimeError):
        db.query(y[:2], num_neighbours=1)                     # empty index: loud, like num_neighbours checks in the reference


def test_exclusive_mode_k_equals_factor(lib):
    """training_loop.py:383: num_neighbours = num_samples_factor (10), every query gets exactly k results."""
    from inclusivegan_b200 import DCI
    x, y = make("gauss", 3000, 48, 300, seed=21)
    db = DCI(300, 3, 15)
    db.add(x, num_levels=3, field_of_view=10, prop_to_retrieve=0.002)
    idx, dist = db.query(y, num_neighbours=10, field_of_view=200, prop_to_retrieve=1.0)
    a = np.array(idx)
    assert a.shape == (48, 10)
    ri, rd = ko.exact_knn_c(x, y, 10)
    ok, msg = ko.compare_knn(a, np.array(dist), ri, rd, x, y)
    assert ok, msg
    assert np.all(np.diff(np.array(dist), axis=1) >= 0)


def test_index_selection_remaps_to_original_rows(lib):
    """dci.py:224-270,315-316: `indices` picks rows; results refer to rows of the array passed in."""
    from inclusivegan_b200 import DCI
    x, y = make("gauss", 2000, 40, 96, seed=30)
    for sel in (slice(500, 1500), slice(1, 2000, 3), np.array([5, 1999, 700, 3, 1200] + list(range(100, 160)), dtype=np.intc),
                (np.arange(2000) % 7 == 0)):
        rows = np.arange(2000)[sel]
        db = DCI(96)
        db.add(x, indices=sel)
        assert db.num_points == len(rows)
        idx, dist = db.query_arrays(y, 3)
        ri, rd = ko.exact_knn_c(np.ascontiguousarray(x[rows]), y, 3)
        ok, msg = ko.compare_knn(idx, dist, rows[ri].astype(np.int32), rd, x, y)
        assert ok, msg


def test_all_neighbours_when_k_negative(lib):
    """dci.py:278-279: num_neighbours < 0 -> all points, ascending."""
    from inclusivegan_b200 import DCI
    x, y = make("gauss", 150, 6, 32, seed=31)
    db = DCI(32)
    db.add(x)
    idx, dist = db.query(y, num_neighbours=-1)
    a = np.array(idx)
    assert a.shape == (6, 150) and all(sorted(r) == list(range(150)) for r in a.tolist())
    ri, rd = ko.exact_knn_c(x, y, 150)
    ok, msg = ko.compare_knn(a, np.array(dist), ri, rd, x, y)
    assert ok, msg


@pytest.mark.parametrize("n,q,d,k,dtype", [(5000, 70, 96, 100, np.float64),       # select 100 of 5000 (several 1024-key chunks per row)
                                           (3000, 40, 64, 3000, np.float64),      # k = n > one chunk: sort only
                                           (70000, 33, 32, 2500, np.float32),     # long rows of keys, float32 pool
                                           (2049, 10, 16, 2048, np.float64)])     # k = n - 1: the last key is dropped
def test_large_k_select_and_sort(lib, n, q, d, k, dtype):
    """k > 32 runs the exact scan + scan_topk_kernel (radix select, index-ordered compaction, stable radix sort)."""
    from inclusivegan_b200 import DCI
    x, y = make("gauss", n, q, d, seed=n + k, dtype=dtype)
    db = DCI(d)
    db.add(x)
    idx, dist = check(db, x, y, k)
    assert all(len(set(r)) == k for r in idx.tolist())
    assert np.all(np.diff(dist, axis=1) >= 0)


def test_large_k_ties_come_out_in_index_order(lib):
    """Equal distances (duplicated pool rows, lattice points) must come out by ascending index — also across the
    selection threshold, where only the lowest-indexed of the tied keys belong to the answer."""
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(12)
    base = rng.integers(-3, 4, size=(400, 12)).astype(np.float64)          # small integer lattice: many exact ties
    x = np.ascontiguousarray(np.concatenate([base, base[::-1], base, base[:123]]))      # every row 3-4 times
    y = np.ascontiguousarray(rng.integers(-3, 4, size=(25, 12)).astype(np.float64))
    db = DCI(12)
    db.add(x)
    d2 = ((y[:, None, :] - x[None, :, :]) ** 2).sum(-1)                      # exact in float64 (small integers)
    for k in (40, 700, x.shape[0]):
        idx, dist = db.query_arrays(y, k, squared=True)
        order = np.lexsort((np.broadcast_to(np.arange(x.shape[0]), d2.shape), d2), axis=1)[:, :k]
        assert np.array_equal(idx, order.astype(np.int32)), k
        assert np.array_equal(dist, np.take_along_axis(d2, order, axis=1)), k


def test_audit_cross_checks_a_sample_by_the_exact_scan(lib, monkeypatch):
    """DCI(audit=N) / $B200KNN_AUDIT: N rows of every query call answered again by the float64 scan and compared."""
    from inclusivegan_b200 import DCI
    x, y = make("cluster", 20000, 900, 256, seed=77)
    db = DCI(256, audit=50)
    db.add(x)
    check(db, x, y, 3)
    assert db.audited_queries == 50
    db.query(y[:24], num_neighbours=1)                 # the list-returning face audits too (24 rows: all of them)
    assert db.audited_queries == 74
    monkeypatch.setenv("B200KNN_AUDIT", "7")
    db2 = DCI(256)
    db2.add(x)
    db2.query_arrays(y, 1, squared=True)
    assert db2.audited_queries == 7
    # a corrupted answer is caught: feed the audit a result whose first distance is off
    i, d = db.query_arrays(y, 3)
    d[0, 0] *= 1.001
    with pytest.raises(RuntimeError, match="audit"):
        db._audit_answers(np.ascontiguousarray(y), 3, 0, i, d)


def test_non_contiguous_and_mixed_dtype_queries(lib):
    from inclusivegan_b200 import DCI
    x, y = make("gauss", 3000, 64, 128, seed=33)
    db = DCI(128)
    db.add(x)
    yy = np.asfortranarray(y)                                  # auto-fixed like dci.py:121-127
    i1, d1 = db.query_arrays(yy, 2)
    i2, d2 = db.query_arrays(y.astype(np.float32), 2)          # f32 queries against an f64 pool
    ri, rd = ko.exact_knn_c(x, y, 2)
    assert ko.compare_knn(i1, d1, ri, rd, x, y)[0]
    ri32, rd32 = ko.exact_knn_c(x, y.astype(np.float32).astype(np.float64), 2)
    assert ko.compare_knn(i2, d2, ri32, rd32, x, y.astype(np.float32).astype(np.float64))[0]
    i3, d3 = db.query_arrays(y.astype(np.int64), 1)            # silently cast to float64 like the reference
    assert i3.shape == (64, 1)


def test_uncertified_queries_are_answered_exactly(lib):
    """High-dimensional i.i.d. data: neighbour gaps are of the order of the BF16 rounding perturbation, so a good part
    of the shortlists cannot be certified; the second (collect) tensor pass must still produce the exact answer."""
    from inclusivegan_b200 import DCI
    x, y = make("gauss", 4000, 60, 4096, seed=77)
    db = DCI(4096)
    db.add(x)
    check(db, x, y, 10)
    assert db.stats()["uncertified"] > 0


def test_common_offset_is_removed_by_centering(lib):
    """Rows far from the origin with tiny mutual distances (norm ~640, distances ~0.02): uncertifiable if rounded as
    they are, trivial once the pool mean is subtracted before the BF16 rounding (translation changes no distance)."""
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(77)
    base = 40.0 + rng.standard_normal((1, 256))
    x = np.ascontiguousarray(base + 1e-3 * rng.standard_normal((4000, 256)))
    y = np.ascontiguousarray(base + 1e-3 * rng.standard_normal((50, 256)))
    db = DCI(256)
    db.add(x)
    check(db, x, y, 3)
    assert db.stats()["uncertified"] == 0


def test_overflowing_second_pass_falls_to_the_exact_scan(lib):
    """Two tight clusters far apart: after centering every row is +-40 plus noise far below BF16 resolution, all
    scores inside a cluster tie, the collect lists overflow (2000 > 1024) and the exact float64 scan answers."""
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(78)
    sign = np.where(np.arange(4000) % 2 == 0, 1.0, -1.0)[:, None]
    x = np.ascontiguousarray(40.0 * sign + 1e-3 * rng.standard_normal((4000, 256)))
    y = np.ascontiguousarray(40.0 * np.where(np.arange(50) % 2 == 0, 1.0, -1.0)[:, None] + 1e-3 * rng.standard_normal((50, 256)))
    db = DCI(256)
    db.add(x)
    check(db, x, y, 3)
    st = db.stats()
    assert st["uncertified"] > 0 and st["exact_scanned"] > 0


# ------------------------------------------------------------------------------------------------ device ABI + merge
def test_device_pointer_abi_and_merge_kernel(lib):
    """b200knn_add_device / query_device / merge_topk_device as bench.py uses them for row-sharded pools."""
    torch = pytest.importorskip("torch")
    from inclusivegan_b200.dci import DeviceKNN, F64
    from inclusivegan_b200.sharding import shard_range
    x, y = make("gauss", 9001, 333, 160, seed=40)
    dev = torch.device("cuda:0")
    ty = torch.from_numpy(y).to(dev)
    k, world = 4, 3
    all_i = torch.empty(world, 333, k, dtype=torch.int32, device=dev)
    all_d = torch.empty(world, 333, k, dtype=torch.float64, device=dev)
    keep = []
    for r in range(world):
        a, b = shard_range(9001, world, r)
        tx = torch.from_numpy(x[a:b]).to(dev)
        ix = DeviceKNN(160, 0)
        ix.set_stream(torch.cuda.current_stream().cuda_stream)
        ix.add(tx.data_ptr(), F64, b - a, index_base=a)
        assert ix.query(ty.data_ptr(), F64, 333, k, all_i[r].data_ptr(), all_d[r].data_ptr()) == k
        keep.append((tx, ix))
    out_i = torch.empty(333, k, dtype=torch.int32, device=dev)
    out_d = torch.empty(333, k, dtype=torch.float64, device=dev)
    keep[0][1].merge(all_i.data_ptr(), all_d.data_ptr(), world, 333, k, out_i.data_ptr(), out_d.data_ptr(),
                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ri, rd = ko.exact_knn_c(x, y, k)
    ok, msg = ko.compare_knn(out_i.cpu().numpy(), out_d.cpu().numpy(), ri, rd, x, y)
    assert ok, msg
    # the numpy restatement used by the gloo test agrees with the kernel
    mi, md = ko.merge_topk_numpy(all_i.cpu().numpy(), all_d.cpu().numpy())
    assert np.array_equal(mi, out_i.cpu().numpy()) and np.array_equal(md, out_d.cpu().numpy())


def test_multi_device_handle_if_available(lib):
    """Single-process row sharding over several GPUs (DCI(devices=[...])); skipped on a 1-GPU box."""
    if lib.b200knn_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from inclusivegan_b200 import DCI
    devs = list(range(min(4, lib.b200knn_device_count())))
    x, y = make("gauss", 20000, 300, 256, seed=50)
    db = DCI(256, devices=devs)
    db.add(x)
    check(db, x, y, 5)
    # several chunks per call, float32 rows, a common offset (the centring vector is the GLOBAL mean), clustered queries
    x, y = make("cluster", 60000, 9000, 384, seed=51, dtype=np.float32)
    x += np.float32(3.0); y += np.float32(3.0)
    db = DCI(384, devices=devs)
    db.add(x)
    i1, d1 = db.query_arrays(y, 3)
    ri, rd = ko.exact_knn_numpy(x, y, 3)
    ok, msg = ko.compare_knn(i1, d1, ri, rd, x.astype(np.float64), y.astype(np.float64))
    assert ok, msg
    one = DCI(384)                                  # bit-identical to the single-device handle (canonical summation order)
    one.add(x)
    i0, d0 = one.query_arrays(y, 3)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    # the trainer's 24-row calls through the collective path
    li, ld = db.query(y[:24], num_neighbours=1)
    assert np.array_equal(np.array(li)[:, 0], i1[:24, 0])
    # all neighbours (dci.py:278-279 num_neighbours=-1) of a pool smaller than k per shard: padded shard lists, scan path
    xs, ys = make("gauss", 50, 40, 32, seed=52)
    db = DCI(32, devices=devs)
    db.add(xs)
    li, ld = db.query(ys)                           # num_neighbours=-1 -> 50 per query
    ri, rd = ko.exact_knn_c(xs, ys, 50)
    ok, msg = ko.compare_knn(np.array(li), np.array(ld), ri, rd, xs, ys)
    assert ok, msg
    li, ld = db.query(ys, num_neighbours=7)         # k > rows per shard (13 or 12): collective path with padded lists
    ri, rd = ko.exact_knn_c(xs, ys, 7)
    ok, msg = ko.compare_knn(np.array(li), np.array(ld), ri, rd, xs, ys)
    assert ok, msg


def test_multi_device_next_entry_points_if_available(lib):
    """Self-kNN, the projected entry points and ball membership on a multi-device handle (SURVEY 8f rows on the row-sharded
    pool): bit-identical to the single-device handle.  Skipped on a 1-GPU box."""
    if lib.b200knn_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from inclusivegan_b200 import DCI
    devs = list(range(min(4, lib.b200knn_device_count())))
    # ---- self-kNN (k + self), float32 features, ragged shards, duplicated rows across a shard boundary
    x, _ = make("relu", 9001, 1, 512, seed=60, dtype=np.float32)
    x[len(x) // len(devs)] = x[len(x) // len(devs) - 1]          # first row of shard 1 == last row of shard 0
    one, many = DCI(512), DCI(512, devices=devs)
    one.add(x)
    many.add(x)
    i0, d0 = one.query_self_arrays(4, squared=True)
    i1, d1 = many.query_self_arrays(4, squared=True)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    ri, rd = ko.exact_knn_numpy(x, x[:700], 4, squared=True)
    ok, msg = ko.compare_knn(i1[:700], np.sqrt(d1[:700]), ri, np.sqrt(rd), x, x[:700])
    assert ok, msg
    # ---- ball membership against the sharded pool
    probe = np.ascontiguousarray(np.maximum(np.random.default_rng(61).standard_normal((900, 512)), 0).astype(np.float32))
    r2 = np.ascontiguousarray(d0[:, 3])
    assert np.array_equal(one.ball_membership(probe, r2), many.ball_membership(probe, r2))
    # ---- random projection on every shard: add_projected / query_projected
    rng = np.random.default_rng(62)
    proj = rng.normal(0.0, 1.0 / 256, size=(1536, 256))
    pool = rng.standard_normal((7000, 1536)).astype(np.float32)
    reals = rng.standard_normal((600, 1536)).astype(np.float32)
    a, b = DCI(256), DCI(256, devices=devs)
    for h in (a, b):
        h.set_projector(proj)
        h.add_projected(pool)
    ia, da = a.query_projected_arrays(reals, 5)
    ib, dbb = b.query_projected_arrays(reals, 5)
    assert np.array_equal(ia, ib) and np.array_equal(da, dbb)
    xp, yp = pool.astype(np.float64) @ proj, reals.astype(np.float64) @ proj
    ri, rd = ko.exact_knn_numpy(xp, yp, 5)
    ok, msg = ko.compare_knn(ib, dbb, ri, rd, xp, yp)
    assert ok, msg


def test_add_device_on_a_multi_device_handle_if_available(lib):
    """b200knn_add_device with rows resident on one GPU, handle sharded over several: the library copies the slices over
    NVLink; answers equal the host add()."""
    if lib.b200knn_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch
    from inclusivegan_b200 import DCI
    devs = list(range(min(4, lib.b200knn_device_count())))
    x, y = make("cluster", 30000, 400, 320, seed=63)
    ref = DCI(320, devices=devs)
    ref.add(x)
    i0, d0 = ref.query_arrays(y, 3)
    h = DCI(320, devices=devs)
    xd = torch.from_numpy(x).to("cuda:0")
    rc = lib.b200knn_add_device(h._handle, ctypes.c_void_p(xd.data_ptr()), 0, x.shape[0], x.shape[1], 0)
    assert rc == 0, lib.b200knn_last_error()
    assert h.num_points == x.shape[0]
    i1, d1 = h.query_arrays(y, 3)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


# ------------------------------------------------------------------------------------------------ full size
def test_full_size_config3_properties(lib):
    """BASELINE config 3 shape (300k pool x 30k queries, d=3072, k=1) through size-independent properties:
      * planted neighbours: query i = pool[p_i] + noise much smaller than any inter-point distance -> index p_i;
      * a 2048-query subsample of unplanted rows checked against the oracle over the full pool;
      * idempotence: the same call twice gives identical output."""
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(300)
    n, q, d = 300000, 30000, 3072
    x = rng.standard_normal((n, d), dtype=np.float32)
    plant = rng.integers(0, n, q)
    y = x[plant] + 0.05 * rng.standard_normal((q, d), dtype=np.float32)
    ns = 2048
    y[:ns] = rng.standard_normal((ns, d), dtype=np.float32)             # unplanted rows for the oracle subsample
    db = DCI(d, 3, 15)
    db.add(x, num_levels=3, field_of_view=10, prop_to_retrieve=0.002)
    idx, dist = db.query_arrays(y, 1)
    assert np.array_equal(idx[ns:, 0], plant[ns:].astype(np.int32))
    assert np.all(np.abs(dist[ns:, 0] - 0.05 * np.sqrt(d)) < 0.2)
    ri, rd = ko.exact_knn_numpy(x, y[:ns], 1, qblock=1024)
    ok, msg = ko.compare_knn(idx[:ns], dist[:ns], ri, rd)
    assert ok, msg
    idx2, dist2 = db.query_arrays(y, 1)
    assert np.array_equal(idx, idx2) and np.array_equal(dist, dist2)


def test_full_size_config4_self_knn(lib):
    """BASELINE config 4 shape: self-kNN radii (k = 3 + the row itself) of a 50k x 2048 non-negative float32 feature set
    (relu(N(0,1)), Inception-pool-like), the call ManifoldEstimator makes (metrics/precision_recall.py:73-90).
      * every row finds itself first, at distance 0;
      * a 2048-row subsample checked against the oracle over the full set (squared distances, like the metric);
      * the radii equal the oracle's k-th order statistic."""
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(400)
    n, d, k = 50000, 2048, 4
    x = np.maximum(rng.standard_normal((n, d), dtype=np.float32), 0)
    db = DCI(d)
    db.add(x)
    idx, d2 = db.query_self_arrays(k, squared=True)
    assert idx.shape == (n, k) and np.array_equal(idx[:, 0], np.arange(n, dtype=np.int32)) and np.all(d2[:, 0] == 0.0)
    assert np.all(np.diff(d2, axis=1) >= 0)
    sel = rng.choice(n, 2048, replace=False)
    ri, rd = ko.exact_knn_numpy(x, x[sel], k, squared=True, qblock=1024)
    ok, msg = ko.compare_knn(idx[sel], d2[sel], ri, rd)
    assert ok, msg
    st = db.stats()
    assert st["exact_scanned"] == 0, st          # structured or not, this shape stays on the tensor path


# ------------------------------------------------------------------------------------------------ certificate's error model
def _bf16_round(a):
    """Round-to-nearest-even float32 -> bfloat16 (returned as float64 values), the conversion the convert kernel does."""
    f = np.ascontiguousarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def _tf32_round(a):
    """Round-to-nearest (ties away from zero) float32 -> tf32 (10 explicit mantissa bits): cvt.rna.tf32.f32."""
    f = np.ascontiguousarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


KC = 4096          # $B200KNN_KC default: K elements per tensor-core accumulation unit


def _acc_geom(kp, tier):
    """Shard::acc_geom: (k_unit, n_units) the error model charges."""
    kelems = 32 if tier == 2 else 64
    nkb = -(-kp // kelems) * (3 if tier == 1 else 1)
    kbu = KC // kelems
    per = kbu if nkb >= 2 * kbu else nkb
    return min(per, nkb) * kelems, -(-nkb // per)


@pytest.mark.parametrize("kind,n,q,d,tier", [
    ("gauss", 20000, 256, 3072, "bf16"), ("image", 20000, 256, 3072, "bf16"), ("relu", 30000, 512, 2048, "bf16"),
    ("gauss", 6000, 130, 5000, "bf16"),
    ("gauss", 3000, 130, 49152, "bf16"),          # long rows: K-chunked accumulation (12 units of 4096)
    ("mixed", 5000, 130, 12288, "bf16"),          # magnitudes spread over six decades inside every row
    ("gauss", 20000, 256, 3072, "bf16x3"), ("mixed", 5000, 130, 12288, "bf16x3"), ("gauss", 3000, 130, 49152, "bf16x3"),
    ("gauss", 20000, 256, 3072, "tf32"), ("image", 6000, 130, 9000, "tf32")])
def test_tensor_scores_within_the_certified_error_model(lib, kind, n, q, d, tier):
    """The certificate is only as good as its error model (EMPIRICALLY validated, not proven: the accumulation model of the
    tensor core is an assumption).  Pull the raw tensor-core scores of the shortlists and check them against float64
    arithmetic on the rounded (mean-centred) operands of the tier: |s~_gpu - s~_exact| must stay below the eps_acc the
    kernels assume (csrc/rerank.cuh make_err_model), and every kept score must bracket the true distance through the
    exact perturbation norms ||q-q^||, ||x-x^||."""
    from inclusivegan_b200 import DCI
    if kind == "mixed":
        rng = np.random.default_rng(d + n)
        scale = 10.0 ** rng.uniform(-3, 3, d)
        x = rng.standard_normal((n, d)) * scale
        y = rng.standard_normal((q, d)) * scale
    else:
        x, y = make(kind, n, q, d, seed=d + n)
    t = {"bf16": 0, "bf16x3": 1, "tf32": 2}[tier]
    db = DCI(d, precision=tier)
    db.add(x)
    idx, dist = db.query_arrays(y, 1, flags=FLAG_NO_CERTIFY)
    scores, rows = db.debug_shortlists()
    assert scores.shape[0] == q and rows.max() < n
    mu = x.mean(axis=0)                                        # the library subtracts the pool mean before rounding
    xc, yc = x - mu, y - mu
    if t == 2:
        xh, yh = _tf32_round(xc), _tf32_round(yc)
        xl = yl = None
        xr, yr = xh, yh
    else:
        xh, yh = _bf16_round(xc), _bf16_round(yc)
        if t == 1:
            xl, yl = _bf16_round(xc - xh), _bf16_round(yc - yh)
            xr, yr = xh + xl, yh + yl
        else:
            xl = yl = None
            xr, yr = xh, yh
    xn = np.einsum("ij,ij->i", xr, xr)
    qn = np.einsum("ij,ij->i", yr, yr)
    kp = (d + 7) // 8 * 8
    k_unit, n_units = _acc_geom(kp, t)
    worst_ratio = 0.0
    err_q = np.linalg.norm(yc - yr, axis=1)
    err_x_max = np.linalg.norm(xc - xr, axis=1).max()
    xl_max = np.linalg.norm(xl, axis=1).max() if t == 1 else 0.0
    for i in range(0, q, 7):                                   # a spread of query rows
        valid = rows[i] >= 0
        r = rows[i][valid]
        s_gpu = scores[i][valid].astype(np.float64)
        if t == 1:                                             # what the three MMAs compute: hi.hi + hi.lo + lo.hi
            ql = np.linalg.norm(yl[i])
            s_ref = xn[r] - 2.0 * (xh[r] @ yh[i] + xl[r] @ yh[i] + xh[r] @ yl[i])
            qh, xhh = np.sqrt(qn[i]) + ql, np.sqrt(xn.max()) + xl_max
            mag = np.sqrt((2 * qh * qh + ql * ql) * (2 * xhh * xhh + xl_max * xl_max))
            extra = 2.0 * ql * xl_max * (1 + 1e-6)
        else:
            s_ref = xn[r] - 2.0 * (xr[r] @ yr[i])
            mag = np.sqrt(qn[i] * xn.max())
            extra = 0.0
        eps_mma = (k_unit + 8.0) * 2.4e-7 * mag * 1.001 + n_units * 1.2e-7 * mag
        eps_mma += (kp / 16.0 + 8.0) * 1.2e-7 * (xn.max() + qn[i]) if t == 0 else 4.8e-7 * (xn.max() + qn[i] + 2.0 * mag)
        worst_ratio = max(worst_ratio, float(np.max(np.abs(s_gpu - s_ref)) / eps_mma))
        # bracket of the true distance (the inequality the pruning rule and the certificate rely on)
        eps = eps_mma + extra
        d_true = np.linalg.norm(x[r] - y[i], axis=1)
        eta = err_q[i] + err_x_max
        lo = np.sqrt(np.maximum(s_gpu + qn[i] - eps, 0.0)) - eta
        hi = np.sqrt(np.maximum(s_gpu + qn[i] + eps, 0.0)) + eta
        assert np.all(lo <= d_true * (1 + 1e-12)) and np.all(d_true <= hi * (1 + 1e-12))
    assert worst_ratio < 1.0, "tensor-core accumulation error exceeds the modelled eps_acc (ratio %.3f)" % worst_ratio
    print("%s %s d=%d: max |s_gpu - s_ref| / eps_acc = %.4f (k_unit %d, units %d)" % (tier, kind, d, worst_ratio, k_unit, n_units))


@pytest.mark.parametrize("tier", ["bf16x3", "tf32"])
@pytest.mark.parametrize("kind,n,q,d,k,dtype", [("gauss", 20011, 300, 256, 10, np.float64), ("cluster", 15000, 700, 3072, 1, np.float32),
                                               ("image", 9001, 260, 9000, 3, np.float32), ("lowrank", 10000, 100, 5000, 10, np.float64)])
def test_precision_tiers_are_exact(lib, tier, kind, n, q, d, k, dtype):
    """the split-BF16 (three MMAs) and TF32 flavours of the tensor pass answer exactly like the BF16 one —
    bit-identical results (the exact re-rank and its canonical summation order do not depend on the tier) — and leave
    fewer queries to the second pass."""
    from inclusivegan_b200 import DCI
    x, y = make(kind, n, q, d, seed=n + d, dtype=dtype)
    base = DCI(d)
    base.add(x)
    i0, d0 = check(base, x.astype(np.float64), y.astype(np.float64), k) if dtype == np.float64 else base.query_arrays(y, k)
    db = DCI(d, precision=tier)
    db.add(x)
    i1, d1 = db.query_arrays(y, k)
    ri, rd = ko.exact_knn_numpy(x, y, k)
    ok, msg = ko.compare_knn(i1, d1, ri, rd, x.astype(np.float64), y.astype(np.float64))
    assert ok, msg
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    assert db.stats()["uncertified"] <= base.stats()["uncertified"]
    # self-kNN uses the pool's own tier operands
    si, sd = db.query_self_arrays(min(k + 1, 4))
    assert np.array_equal(si[:, 0], np.arange(n, dtype=np.int32)) and np.all(sd[:, 0] == 0)
    # switching the tier of a loaded index rebuilds its operands
    base.set_precision(tier)
    i2, d2 = base.query_arrays(y, k)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_unstructured_long_rows_stay_on_the_tensor_path(lib):
    """VERDICT r1 / SURVEY 8f-4: i.i.d. Gaussian rows at d = 49152 — every pairwise distance within a fraction of a percent
    of every other, BF16 rounding moves a point further than the neighbour gaps.  Round 1 answered these on the exact
    float64 CUDA-core scan (~1k queries/s).  With the split-BF16 tier (rounding 2^-17) and K-chunked accumulation (error
    bound per 4096 products instead of per 49152) the tensor path certifies them: nothing falls to the scan."""
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(4915)
    n, q, d = 20000, 512, 49152
    x = rng.standard_normal((n, d), dtype=np.float32)
    y = rng.standard_normal((q, d), dtype=np.float32)
    db = DCI(d, precision="bf16x3")
    db.add(x)
    idx, dist = db.query_arrays(y, 1)
    ri, rd = ko.exact_knn_numpy(x, y, 1, qblock=512)
    ok, msg = ko.compare_knn(idx, dist, ri, rd)
    assert ok, msg
    st = db.stats()
    assert st["exact_scanned"] == 0, st
    print("bf16x3, i.i.d. d=49152: %d of %d queries needed the second (collection) pass, none the exact scan" % (st["uncertified"], q))


# ------------------------------------------------------------------------------------------------ C ABI corner cases
def test_cabi_strided_rows_and_pageable_uploads(lib):
    """ld > dim for both matrices (rows embedded in a wider host array) straight through the C ABI, large enough
    that the pageable upload goes through the pinned ring, float64 pool + float32 queries."""
    rng = np.random.default_rng(60)
    n, q, d, ld = 40000, 5000, 200, 264
    big_x = rng.standard_normal((n, ld))
    big_q = rng.standard_normal((q, ld)).astype(np.float32)
    h = ctypes.c_void_p()
    assert lib.b200knn_create(d, 0, None, ctypes.byref(h)) == 0
    assert lib.b200knn_add(h, big_x.ctypes.data, 0, n, ld) == 0, lib.b200knn_last_error()
    assert lib.b200knn_add(h, big_x.ctypes.data, 0, n, ld) == -2          # second add refused (dci.py:228-229)
    oi = np.empty((q, 3), np.int32); od = np.empty((q, 3))
    kk = ctypes.c_int(0)
    assert lib.b200knn_query(h, big_q.ctypes.data, 1, q, ld, 3, 0, oi.ctypes.data, od.ctypes.data, ctypes.byref(kk)) == 0, lib.b200knn_last_error()
    assert kk.value == 3
    x = np.ascontiguousarray(big_x[:, :d]); y = np.ascontiguousarray(big_q[:, :d]).astype(np.float64)
    ri, rd = ko.exact_knn_numpy(x, y, 3)
    ok, msg = ko.compare_knn(oi, od, ri, rd, x, y)
    assert ok, msg
    assert lib.b200knn_clear(h) == 0 and lib.b200knn_num_points(h) == 0
    assert lib.b200knn_query(h, big_q.ctypes.data, 1, q, ld, 3, 0, oi.ctypes.data, od.ctypes.data, None) == -2
    assert lib.b200knn_destroy(h) == 0


def test_large_query_batches_are_chunked_consistently(lib):
    """More queries than one device pass (32768) and than one upload chunk: chunk boundaries must not show."""
    from inclusivegan_b200 import DCI
    x, y = make("gauss", 3000, 70000, 64, seed=61, dtype=np.float32)
    db = DCI(64)
    db.add(x)
    idx, dist = db.query_arrays(y, 2)
    ri, rd = ko.exact_knn_numpy(x, y, 2)
    ok, msg = ko.compare_knn(idx, dist, ri, rd, x, y)
    assert ok, msg


def test_extension_stand_in_functions(lib):
    """inclusivegan_b200._dci (stand-in for the reference's compiled `_dci`, INTEGRATION.md option B) called the way
    dci_code/src/dci.py calls the extension: add(inst, data, start, end, ...), query(...) -> flat idx, flat dist, counts."""
    from inclusivegan_b200 import _dci
    x, y = make("gauss", 2500, 40, 96, seed=70)
    inst = _dci.new(96, 2, 7)
    assert _dci.get_num_points(inst) == 0 and _dci.get_proj_vec(inst).shape == (14, 96)
    _dci.add(inst, x, 500, 2500, 2, False, -1, -1, 1.0, 0.002, -1)                  # dci.py:263 argument order
    assert _dci.get_num_points(inst) == 2000 and _dci.get_num_levels(inst) == 2
    flat_i, flat_d, counts = _dci.query(inst, y, 3, False, -1, -1, 1.0, 0.05, 100)  # dci.py:313
    assert counts.tolist() == [3] * 40 and flat_i.dtype == np.int32 and flat_d.dtype == np.float64
    ri, rd = ko.exact_knn_c(np.ascontiguousarray(x[500:]), y, 3)
    ok, msg = ko.compare_knn(flat_i.reshape(40, 3), flat_d.reshape(40, 3), ri + 500, rd, x, y)
    assert ok, msg
    _dci.reset(inst)
    assert _dci.get_num_points(inst) == 0


def test_results_do_not_depend_on_batching_or_path(lib):
    """Every emitting kernel sums a distance in one canonical order (csrc/rerank.cuh canon_d2): the same query answered in
    one big call, in 24-row calls (different kernel flavour, 1024-thread re-rank), or via the second pass gives
    bit-identical distances and indices."""
    from inclusivegan_b200 import DCI
    x, y = make("gauss", 4000, 600, 4096, seed=90)          # high-d i.i.d.: a good share of rows needs the second pass
    db = DCI(4096)
    db.add(x)
    i_all, d_all = db.query_arrays(y, 10)
    assert db.stats()["uncertified"] > 0
    i_parts = np.concatenate([db.query_arrays(y[s:s + 24], 10)[0] for s in range(0, 600, 24)])
    d_parts = np.concatenate([db.query_arrays(y[s:s + 24], 10)[1] for s in range(0, 600, 24)])
    assert np.array_equal(i_all, i_parts) and np.array_equal(d_all, d_parts)
    i_nc, d_nc = db.query_arrays(y, 10, flags=FLAG_NO_CERTIFY)
    certified_rows = np.all(i_nc == i_all, axis=1)
    assert certified_rows.mean() > 0.2
    assert np.array_equal(d_nc[certified_rows], d_all[certified_rows])


@pytest.mark.parametrize("kind,n,q,d,k,dtype", [("cluster", 30000, 2500, 384, 1, np.float64), ("relu", 9000, 2100, 2048, 4, np.float32),
                                                ("gauss", 6000, 2200, 520, 20, np.float64)])
def test_rerank_flavours_are_bit_identical(lib, kind, n, q, d, k, dtype, monkeypatch):
    """The re-rank has a block-per-query flavour (few queries: lowest latency) and a warp-per-query flavour (many queries:
    24 resident per SM).  Same pruning, same canonical summation order: outputs must be bit-identical, and exact."""
    from inclusivegan_b200 import DCI
    x, y = make(kind, n, q, d, seed=n + k, dtype=dtype)
    out = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("B200KNN_RERANK_WARP", mode)          # read when the handle's device state is created
        db = DCI(d)
        db.add(x)
        out[mode] = check(db, x, y, k)
        small = db.query_arrays(y[:40], k)                        # 40 rows: auto mode would take the block flavour
        assert np.array_equal(small[0], out[mode][0][:40]) and np.array_equal(small[1], out[mode][1][:40])
        db.clear()
    assert np.array_equal(out["0"][0], out["2"][0]) and np.array_equal(out["0"][1], out["2"][1])


def test_one_second_pass_per_call_over_several_chunks(lib, monkeypatch):
    """A host-row call of several chunks keeps every chunk's converted rows and runs ONE collection pass at its end
    (Shard::CallAccum): uncertified rows of different chunks are answered exactly, identically to chunk-sized calls, with
    and without the whole-call buffers, from pageable and from page-locked rows."""
    import torch
    from inclusivegan_b200 import DCI
    x, y = make("gauss", 4000, 2600, 4096, seed=91)           # high-d i.i.d. (as in test_results_do_not_depend_on_batching_or_path):
    db = DCI(4096)                                              # a share of the rows is uncertified
    db.add(x)
    monkeypatch.setenv("B200KNN_UPLOAD_RAMP", "2")             # cut the call into several (growing) chunks whatever the source
    i_all, d_all = check(db, x, y, 10)
    unc = db.stats()["uncertified"]
    assert unc >= 2
    parts = [db.query_arrays(y[s:s + 500], 10) for s in range(0, 2600, 500)]
    assert np.array_equal(i_all, np.concatenate([p[0] for p in parts])) and np.array_equal(d_all, np.concatenate([p[1] for p in parts]))
    yp = torch.from_numpy(y).pin_memory().numpy()              # page-locked rows
    i_pin, d_pin = db.query_arrays(yp, 10)
    assert np.array_equal(i_all, i_pin) and np.array_equal(d_all, d_pin)
    monkeypatch.setenv("B200KNN_CALL_BUFFER_MB", "0")          # the round-2-first-session path: stage buffers recycled, a pass per chunk
    monkeypatch.setenv("B200KNN_UPLOAD_RAMP", "0")
    i_old, d_old = db.query_arrays(y, 10)
    assert np.array_equal(i_all, i_old) and np.array_equal(d_all, d_old)


@pytest.mark.parametrize("cg", ["1", "2"])
def test_round_wide_lockstep_schedule(lib, cg, monkeypatch):
    """Long-K schedule (Shard::plan: a grid of query tiles x pool streams advancing through K together) forced at a
    size with several rounds and ragged chunks; answers must not depend on the schedule."""
    from inclusivegan_b200 import DCI
    x, y = make("cluster", 21000, 2900, 320, seed=33, dtype=np.float32)
    ref = DCI(320)
    ref.add(x)
    i0, d0 = ref.query_arrays(y, 4)
    monkeypatch.setenv("B200KNN_WIDE", "2")
    monkeypatch.setenv("B200KNN_CTA_GROUP", cg)
    db = DCI(320)
    db.add(x)
    i1, d1 = check(db, x, y, 4)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    # the second pass (collect mode) and the ball-membership filter run on the same schedule
    xg, yg = make("gauss", 4000, 600, 4096, seed=90)
    dg = DCI(4096)
    dg.add(xg)
    check(dg, xg, yg, 10)
    assert dg.stats()["uncertified"] > 0
    r2 = np.full(21000, 0.9 * float(np.median(d1[:, 0])) ** 2)
    member = db.ball_membership(y, r2)
    assert np.array_equal(member.astype(bool), d1[:, 0] ** 2 <= r2[0])


def test_long_rows_pick_the_wide_schedule_and_stay_exact(lib):
    """d = 49152 (config 5's raw pixels) on image-like rows: first pass on the tensor path, exact answers."""
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(5)
    d, n, q = 49152, 6000, 700
    basis = (0.125 * rng.standard_normal((16, d))).astype(np.float32)
    x = np.clip(rng.standard_normal((n, 16)).astype(np.float32) @ basis + 0.05 * rng.standard_normal((n, d)).astype(np.float32), -1, 1)
    y = np.clip(rng.standard_normal((q, 16)).astype(np.float32) @ basis + 0.05 * rng.standard_normal((q, d)).astype(np.float32), -1, 1)
    db = DCI(d)
    db.add(x)
    check(db, x, y, 10)
    st = db.stats()
    assert st["exact_scanned"] == 0, st
