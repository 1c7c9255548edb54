"""k-NN precision/recall adaptor: oracle sanity on CPU, parity of the B200 path on the GPU."""
import numpy as np
import pytest

from oracle import pr_oracle


def two_sets(n, m, d, seed, spread=1.25, latent=8):
    """Features on a low-dimensional manifold (non-negative, Inception-pool-like); the eval set is wider than the
    reference set, so precision and recall both land strictly between 0 and 1."""
    rng = np.random.default_rng(seed)
    w = rng.standard_normal((latent, d)) / np.sqrt(latent)
    ref = np.maximum(rng.standard_normal((n, latent)) @ w + 0.05 * rng.standard_normal((n, d)), 0).astype(np.float32)
    ev = np.maximum(spread * rng.standard_normal((m, latent)) @ w + 0.05 * rng.standard_normal((m, d)), 0).astype(np.float32)
    return ref, ev


def test_oracle_definitions_small():
    """Hand-checkable case: points on a line."""
    ref = np.array([[0.0], [1.0], [3.0], [7.0]])
    D = pr_oracle.manifold_radii(ref, [1, 2])
    # squared distance to 1st / 2nd other point
    assert np.allclose(D, [[1, 9], [1, 4], [4, 9], [16, 36]])
    ev = np.array([[0.5], [5.2], [20.0]])
    pred, realism, nearest, _ = pr_oracle.evaluate(ref, D, ev)
    assert pred.tolist() == [[1, 1], [1, 1], [0, 0]]          # 5.2 is within 4 of 7 (r^2=16); 20 is outside everything
    assert nearest.tolist() == [0, 3, 3]
    assert np.isclose(realism[0], 1 / 0.25)


def _load(golden_dir, name):
    import os
    with np.load(os.path.join(golden_dir, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def test_oracle_float16_mode_reproduces_the_reference_class(golden_dir):
    """pr_generic.npz: outputs of the reference's own ManifoldEstimator (float16 storage, metrics/precision_recall.py:72-73,100)
    fed by a NumPy distance block.  The oracle with the same storage casts must reproduce them bit for bit."""
    g = _load(golden_dir, "pr_generic")
    nh = [int(v) for v in g["nhood_sizes"]]
    for feats, key in ((g["ref"], "D_ref"), (g["ev"], "D_ev")):
        D = pr_oracle.manifold_radii(feats, nh, row_batch=int(g["row_batch"]), store_dtype=np.float16)
        assert D.dtype == np.float16 and np.array_equal(D, g[key])
    pred, realism, nearest, _ = pr_oracle.evaluate(g["ref"], g["D_ref"], g["ev"], row_batch=int(g["row_batch"]), store_dtype=np.float16)
    assert np.array_equal(pred, g["precision"]) and np.array_equal(nearest, g["nearest"])
    np.testing.assert_array_equal(realism.astype(np.float32), g["realism"])
    rec = pr_oracle.evaluate(g["ev"], g["D_ev"], g["ref"], row_batch=int(g["row_batch"]), store_dtype=np.float16)[0]
    assert np.array_equal(rec, g["recall"])
    # and the float64 definition differs from the float16 pipeline only by float16 rounding of the radii
    D64 = pr_oracle.manifold_radii(g["ref"], nh)
    np.testing.assert_allclose(D64, g["D_ref"].astype(np.float64), rtol=2.0 ** -10)


def test_oracle_float64_mode_equals_the_reference_where_float16_is_exact(golden_dir):
    """pr_lattice.npz: integer features, every squared distance an integer <= 2048 — float16 holds them exactly, so the
    reference's outputs are the real-number answer and the float64 oracle must equal them."""
    g = _load(golden_dir, "pr_lattice")
    nh = [int(v) for v in g["nhood_sizes"]]
    D_ref = pr_oracle.manifold_radii(g["ref"], nh)
    D_ev = pr_oracle.manifold_radii(g["ev"], nh)
    assert np.array_equal(D_ref, g["D_ref"].astype(np.float64)) and np.array_equal(D_ev, g["D_ev"].astype(np.float64))
    pred, realism, nearest, _ = pr_oracle.evaluate(g["ref"], D_ref, g["ev"])
    assert np.array_equal(pred, g["precision"]) and np.array_equal(nearest, g["nearest"])
    fin = np.isfinite(g["realism"])
    assert np.array_equal(np.isfinite(realism), fin)
    np.testing.assert_allclose(realism[fin], g["realism"][fin], rtol=2.0 ** -10)       # the reference divides in float16
    assert np.array_equal(pr_oracle.evaluate(g["ev"], D_ev, g["ref"])[0], g["recall"])


@pytest.mark.gpu
def test_gpu_path_equals_reference_outputs_on_the_lattice(native_lib, golden_dir):
    """The B200 path against outputs of the reference's own class (exact where float16 is exact)."""
    from inclusivegan_b200.precision_recall import ManifoldEstimator, knn_precision_recall_features
    g = _load(golden_dir, "pr_lattice")
    nh = [int(v) for v in g["nhood_sizes"]]
    est = ManifoldEstimator(None, g["ref"], nhood_sizes=nh)
    assert np.array_equal(est.D, g["D_ref"].astype(np.float64))
    pred, realism, nearest = est.evaluate(g["ev"], return_realism=True, return_neighbors=True)
    assert np.array_equal(pred, g["precision"]) and np.array_equal(nearest, g["nearest"])
    fin = np.isfinite(g["realism"])
    np.testing.assert_allclose(realism[fin], g["realism"][fin], rtol=2.0 ** -10)
    state = knn_precision_recall_features(g["ref"], g["ev"], nhood_sizes=nh)
    np.testing.assert_array_equal(state.recall, g["recall"])
    np.testing.assert_allclose(state.knn_precision, g["precision"].mean(axis=0))
    np.testing.assert_allclose(state.knn_recall, g["recall"].mean(axis=0))


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,d,nhoods", [(2000, 1500, 256, [3]), (1200, 1000, 2048, [3]), (1500, 1200, 128, [1, 3, 5])])
def test_precision_recall_matches_float64_oracle(native_lib, n, m, d, nhoods):
    from inclusivegan_b200.precision_recall import ManifoldEstimator, knn_precision_recall_features
    assert native_lib.b200knn_device_count() >= 1
    ref, ev = two_sets(n, m, d, seed=n + d)
    D_ref = pr_oracle.manifold_radii(ref, nhoods)
    est = ManifoldEstimator(None, ref, nhood_sizes=nhoods)
    np.testing.assert_allclose(est.D, D_ref, rtol=1e-9)                      # radii: exact self-kNN
    pred, realism, nearest = est.evaluate(ev, return_realism=True, return_neighbors=True)
    o_pred, o_real, o_near, margin = pr_oracle.evaluate(ref, D_ref, ev)
    disagree = pred != o_pred
    assert np.all(margin[disagree] < 1e-9), "membership differs away from a ball surface"
    assert 0.05 < o_pred.mean() < 0.95                                       # the case exercises both outcomes
    assert np.array_equal(nearest, o_near)
    np.testing.assert_allclose(realism, o_real, rtol=1e-5)
    state = knn_precision_recall_features(ref, ev, nhood_sizes=nhoods)
    D_ev = pr_oracle.manifold_radii(ev, nhoods)
    o_rec = pr_oracle.evaluate(ev, D_ev, ref)[0]
    np.testing.assert_allclose(state.knn_precision, o_pred.mean(axis=0), atol=2.0 / m)
    np.testing.assert_allclose(state.knn_recall, o_rec.mean(axis=0), atol=2.0 / n)


@pytest.mark.gpu
def test_membership_deep_inside_and_far_outside(native_lib):
    """Lists that overflow (a point inside thousands of balls) and empty lists both decide correctly."""
    from inclusivegan_b200 import DCI
    rng = np.random.default_rng(3)
    x = rng.standard_normal((6000, 64))
    db = DCI(64)
    db.add(x)
    r2 = np.full(6000, 400.0)                      # huge balls: everything nearby is inside all of them
    q = np.concatenate([rng.standard_normal((50, 64)), 100.0 + rng.standard_normal((50, 64))])
    flags = db.ball_membership(q, r2)
    assert flags[:50].all() and not flags[50:].any()
    r2 = np.zeros(6000)                            # degenerate balls: only exact duplicates are members
    flags = db.ball_membership(np.concatenate([x[:10], x[:10] + 1e-3]), r2)
    assert flags[:10].all() and not flags[10:].any()
