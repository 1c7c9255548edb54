"""Random projection on the device: the trainer's `rows.astype(float64) @ projector`
(training/training_loop.py:205-212,362-365,379-381) computed by libb200knn instead of NumPy on the host.

Parity: projected rows against NumPy float64 matmul (1e-12 relative to the row scale — only the summation order
differs), and the search on unprojected rows against the float64 oracle run on host-projected rows.
"""
import numpy as np
import pytest

from oracle import knn_oracle as ko

pytestmark = pytest.mark.gpu


def trainer_projector(in_dim, proj_dim, seed):
    # training_loop.py:211: N(0, 1/proj_dim) entries, float64
    return np.random.default_rng(seed).normal(0.0, 1.0 / float(proj_dim), size=(in_dim, proj_dim)).astype(np.float64)


def images(n, shape, seed):
    # generator output: float32 NCHW in [-1, 1]
    rng = np.random.default_rng(seed)
    lat = rng.standard_normal((n, 12)).astype(np.float32)
    basis = (0.3 * rng.standard_normal((12, int(np.prod(shape))))).astype(np.float32)
    return np.clip(lat @ basis + 0.05 * rng.standard_normal((n, int(np.prod(shape)))).astype(np.float32), -1, 1).reshape((n,) + shape)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,in_dim,dim", [(300, 3072, 200), (1000, 1000, 129), (37, 77, 5), (260, 4099, 256)])
def test_projected_rows_match_numpy_float64(native_lib, dtype, n, in_dim, dim):
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(n + in_dim)
    rows = rng.standard_normal((n, in_dim)).astype(dtype)
    proj = trainer_projector(in_dim, dim, 3)
    db = DCI(dim)
    db.set_projector(proj)
    got = db.project_rows(rows)
    want = rows.astype(np.float64) @ proj
    scale = np.linalg.norm(rows.astype(np.float64), axis=1, keepdims=True) * np.linalg.norm(proj, axis=0, keepdims=True)
    assert np.max(np.abs(got - want) / scale) < 1e-14
    # one FMA chain per element: chunking the rows differently cannot change a bit
    half = db.project_rows(rows[: n // 2])
    assert np.array_equal(half, got[: n // 2])


def test_search_on_unprojected_images_equals_oracle_on_host_projected_rows(native_lib):
    from inclusivegan_b200 import DCI
    shape = (3, 32, 32)
    pool, reals = images(6000, shape, 1), images(500, shape, 2)
    proj = trainer_projector(3072, 384, 7)
    xp = pool.reshape(len(pool), -1).astype(np.float64) @ proj          # what training_loop.py:365 hands to DCI.add
    yp = reals.reshape(len(reals), -1).astype(np.float64) @ proj        # training_loop.py:381
    db = DCI(384, 3, 15)
    db.set_projector(proj)
    db.add_projected(pool, num_levels=3, field_of_view=10)
    assert db.num_points == 6000
    idx, dist = db.query_projected_arrays(reals, 10)
    ri, rd = ko.exact_knn_numpy(xp, yp, 10)
    ok, msg = ko.compare_knn(idx, dist, ri, rd, xp, yp)
    assert ok, msg
    # same answer as the host-matmul route through add()/query()
    ref = DCI(384)
    ref.add(np.ascontiguousarray(xp))
    i2, d2 = ref.query_arrays(yp, 10)
    assert np.array_equal(idx, i2)
    np.testing.assert_allclose(dist, d2, rtol=1e-9)
    # list-returning flavour, trainer's 24-row calls
    li, ld = db.query_projected(reals[:24], num_neighbours=1)
    assert len(li) == 24 and all(a.shape == (1,) for a in li)
    assert np.array_equal(np.array(li)[:, 0], idx[:24, 0])
    db.reset()
    assert db.num_points == 0
    db.add_projected(pool[:100])                       # projector survives reset()
    assert db.num_points == 100


def test_projection_argument_errors(native_lib):
    from inclusivegan_b200 import DCI
    db = DCI(16)
    with pytest.raises(RuntimeError):
        db.add_projected(np.zeros((4, 32), np.float32))            # no projector yet
    with pytest.raises(ValueError):
        db.set_projector(np.zeros((32, 15)))                        # wrong output width
    db.set_projector(trainer_projector(32, 16, 0))
    with pytest.raises(ValueError):
        db.add_projected(np.zeros((4, 31), np.float32))            # wrong input width
    db.add_projected(np.random.default_rng(0).standard_normal((50, 32)).astype(np.float32))
    with pytest.raises(RuntimeError):
        db.add_projected(np.zeros((4, 32), np.float32))            # one array per index
