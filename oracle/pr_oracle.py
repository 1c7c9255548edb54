"""TEST INFRASTRUCTURE — not product code.

NumPy restatement of the k-NN precision/recall arithmetic of the reference's metrics/precision_recall.py
(ManifoldEstimator.__init__ :60-94 and .evaluate :96-134).  Blocked over rows like the reference.

`store_dtype` selects the precision distances and radii are STORED in:
  * float64 (default): the real-number definition, which the B200 path implements exactly;
  * float16: the reference's own storage (`self.D` and `distance_batch` are float16 arrays, :72-73, :100) — in this mode
    the functions reproduce the reference's ManifoldEstimator bit for bit.

Pinned by reference outputs: tests/golden/pr_lattice.npz and pr_generic.npz are produced by the reference's own class
(tests/golden/make_golden_pr.py: TensorFlow stubbed out, a NumPy distance block passed to the constructor — the TF fp16
DistanceBlock :38-57 cannot run here); tests/test_pr_metric.py checks float16 mode == reference on generic data and
float64 mode == reference where float16 is exact (integer lattice).
"""
import numpy as np


def pairwise_sq(u, v):
    """batch_pairwise_distances (:20-34): squared Euclidean, clamped at 0 — here by direct differences in float64."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    out = np.empty((u.shape[0], v.shape[0]))
    for i in range(u.shape[0]):
        diff = v - u[i]
        out[i] = np.einsum("ij,ij->i", diff, diff)
    return out


def manifold_radii(features, nhood_sizes, row_batch=256, store_dtype=np.float64):
    """ManifoldEstimator.__init__ :73-90: squared distance to the k-th neighbour, index 0 = the sample itself."""
    n = features.shape[0]
    D = np.zeros((n, len(nhood_sizes)), dtype=store_dtype)
    seq = np.arange(max(nhood_sizes) + 1)
    for b in range(0, n, row_batch):
        d = pairwise_sq(features[b:b + row_batch], features).astype(store_dtype)
        D[b:b + row_batch] = np.partition(d, seq, axis=1)[:, nhood_sizes]
    return D


def evaluate(ref_features, D, eval_features, row_batch=256, store_dtype=np.float64):
    """ManifoldEstimator.evaluate :96-134 -> (in-manifold flags [Q, nhoods], realism [Q], nearest index [Q]),
    plus the margin |d2 - D| of the closest call per query (to excuse float-level ties in tests)."""
    q = eval_features.shape[0]
    pred = np.zeros((q, D.shape[1]), dtype=np.int32)
    realism = np.zeros(q)
    nearest = np.zeros(q, dtype=np.int32)
    margin = np.zeros((q, D.shape[1]))
    for b in range(0, q, row_batch):
        d = pairwise_sq(eval_features[b:b + row_batch], ref_features).astype(store_dtype)
        inside = d[:, :, None] <= D[None, :, :]
        pred[b:b + row_batch] = np.any(inside, axis=1)
        with np.errstate(over="ignore", invalid="ignore"):
            rel = np.abs(d[:, :, None].astype(np.float64) - D[None, :, :].astype(np.float64)) / np.maximum(D[None, :, :].astype(np.float64), 1e-300)
        margin[b:b + row_batch] = rel.min(axis=1)
        nearest[b:b + row_batch] = np.argmin(d, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            realism[b:b + row_batch] = D[nearest[b:b + row_batch], 0] / d.min(axis=1)
    return pred, realism, nearest, margin
