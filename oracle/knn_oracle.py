"""TEST INFRASTRUCTURE — not product code.

Exact float64 brute-force kNN oracle (see oracle/knn_oracle.c for what it restates and why).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; inclusivegan_b200/ never does.

Two implementations of the same contract:
  * exact_knn_c      — ctypes over oracle/liboracle.so: sequential-sum Euclidean distance in the
                       loop order of the reference's compute_dist (dci_code/src/util.c:62-69);
                       distances bit-identical to the reference's for the same pair.
  * exact_knn_numpy  — blocked BLAS selection (||q||^2+||x||^2-2q.x shortlists a superset), then the
                       shortlist is re-evaluated with the direct float64 sum of squared differences;
                       used where the C loop would take too long (full-size property tests).
Both return (idx int32 [Q, kk], dist float64 [Q, kk]) ascending by (distance, index), kk=min(k, N).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(quiet=True):
    """Compile liboracle.so (and oracle/_ref/_dci.so when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.knn_oracle_f64.restype = ctypes.c_int
        lib.knn_oracle_f64.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.knn_oracle_pair_dist_f64.restype = None
        lib.knn_oracle_pair_dist_f64.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def exact_knn_c(data, query, k, squared=False):
    data = np.ascontiguousarray(data, dtype=np.float64)
    query = np.ascontiguousarray(query, dtype=np.float64)
    n, d = data.shape
    nq = query.shape[0]
    assert query.shape[1] == d
    kk = min(int(k), n)
    idx = np.empty((nq, kk), dtype=np.int32)
    dist = np.empty((nq, kk), dtype=np.float64)
    r = _lib().knn_oracle_f64(data.ctypes.data, n, query.ctypes.data, nq, d, int(k), int(bool(squared)),
                              idx.ctypes.data, dist.ctypes.data)
    if r != kk:
        raise RuntimeError("knn_oracle_f64 failed (%d)" % r)
    return idx, dist


def pair_dist(data, query, qrow, xrow):
    """Float64 Euclidean distances of explicit (query row, data row) pairs."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    query = np.ascontiguousarray(query, dtype=np.float64)
    qrow = np.ascontiguousarray(qrow, dtype=np.int64)
    xrow = np.ascontiguousarray(xrow, dtype=np.int64)
    out = np.empty(qrow.shape[0], dtype=np.float64)
    _lib().knn_oracle_pair_dist_f64(data.ctypes.data, query.ctypes.data, data.shape[1], qrow.shape[0],
                                    qrow.ctypes.data, xrow.ctypes.data, out.ctypes.data)
    return out


def exact_knn_numpy(data, query, k, squared=False, qblock=256, xblock=65536, slack=8):
    """Blocked float64 brute force.  `data`/`query` may be float32 or float64 (upcast block by block, so a
    full-size float32 pool never has to exist in float64).  Selection by the GEMM form keeps a shortlist of
    kk+slack per query over pool blocks; the shortlist is then re-evaluated with the direct float64
    sum of squared differences, and a completeness guard falls back to a full direct evaluation."""
    n, d = data.shape
    nq = query.shape[0]
    kk = min(int(k), n)
    c = min(n, kk + slack)
    idx = np.empty((nq, kk), dtype=np.int32)
    dist = np.empty((nq, kk), dtype=np.float64)
    xn_max = 0.0
    for s in range(0, nq, qblock):
        q = np.asarray(query[s:s + qblock], dtype=np.float64)
        b = q.shape[0]
        qn = np.einsum("ij,ij->i", q, q)
        best_v = np.full((b, c), np.inf)
        best_i = np.zeros((b, c), dtype=np.int64)
        for x0 in range(0, n, xblock):
            x = np.asarray(data[x0:x0 + xblock], dtype=np.float64)
            xn = np.einsum("ij,ij->i", x, x)
            xn_max = max(xn_max, float(xn.max()))
            approx = xn[None, :] - 2.0 * (q @ x.T)              # + ||q||^2 is constant per row
            m = min(c, x.shape[0])
            part = np.argpartition(approx, m - 1, axis=1)[:, :m] if m < x.shape[0] else np.broadcast_to(np.arange(x.shape[0]), (b, x.shape[0]))
            cand_v = np.concatenate([best_v, np.take_along_axis(approx, part, axis=1)], axis=1)
            cand_i = np.concatenate([best_i, part + x0], axis=1)
            keep = np.argpartition(cand_v, c - 1, axis=1)[:, :c]
            best_v = np.take_along_axis(cand_v, keep, axis=1)
            best_i = np.take_along_axis(cand_i, keep, axis=1)
        # exact re-evaluation of the shortlist
        d2 = np.empty((b, c))
        for j in range(c):
            diff = np.asarray(data[best_i[:, j]], dtype=np.float64) - q
            d2[:, j] = np.einsum("ij,ij->i", diff, diff)
        order = np.lexsort((best_i, d2), axis=1)
        cand_s = np.take_along_axis(best_i, order, axis=1)
        d2_s = np.take_along_axis(d2, order, axis=1)
        if c < n:
            boundary = best_v.max(axis=1) + qn                  # <= the GEMM-form value of every excluded point
            eps = 1e-9 * (qn + xn_max)
            for r in np.nonzero(d2_s[:, kk - 1] > boundary - eps)[0]:      # not provably complete: direct scan
                fd2 = np.empty(n)
                for x0 in range(0, n, xblock):
                    diff = np.asarray(data[x0:x0 + xblock], dtype=np.float64) - q[r]
                    fd2[x0:x0 + xblock] = np.einsum("ij,ij->i", diff, diff)
                o = np.lexsort((np.arange(n), fd2))[:c]
                cand_s[r], d2_s[r] = o, fd2[o]
        idx[s:s + qblock] = cand_s[:, :kk]
        dist[s:s + qblock] = d2_s[:, :kk] if squared else np.sqrt(d2_s[:, :kk])
    return idx, dist


def merge_topk_numpy(all_idx, all_dist):
    """CPU restatement of the k-way merge of per-shard results (contract of b200knn_merge_topk_device):
    inputs [G, Q, kk] ascending per shard -> [Q, kk] ascending by (distance, index); index -1 = empty slot."""
    all_idx = np.asarray(all_idx)
    all_dist = np.asarray(all_dist, dtype=np.float64)
    g, nq, kk = all_idx.shape
    ci = np.transpose(all_idx, (1, 0, 2)).reshape(nq, g * kk)
    cd = np.transpose(all_dist, (1, 0, 2)).reshape(nq, g * kk).copy()
    cd[ci < 0] = np.inf
    order = np.lexsort((ci, cd), axis=1)[:, :kk]
    return np.take_along_axis(ci, order, axis=1).astype(np.int32), np.take_along_axis(cd, order, axis=1)


def compare_knn(idx, dist, ref_idx, ref_dist, data=None, query=None, tie_rtol=1e-6, dist_rtol=1e-5):
    """north_star acceptance test.  Returns (ok, message).

    indices must equal the oracle's, except where the oracle's own distances tie within
    `tie_rtol` relative (then the returned index must be one of the tied candidates at an
    equivalent distance); distances must agree within `dist_rtol` relative.
    """
    idx = np.asarray(idx)
    ref_idx = np.asarray(ref_idx)
    dist = np.asarray(dist, dtype=np.float64)
    ref_dist = np.asarray(ref_dist, dtype=np.float64)
    if idx.shape != ref_idx.shape or dist.shape != ref_dist.shape:
        return False, "shape mismatch %s vs %s" % (idx.shape, ref_idx.shape)
    scale = np.maximum(np.abs(ref_dist), 1e-300)
    derr = np.abs(dist - ref_dist) / scale
    exact_zero = (ref_dist == 0) & (np.abs(dist) <= 1e-12)
    derr[exact_zero] = 0.0
    if derr.size and derr.max() > dist_rtol:
        q, r = np.unravel_index(np.argmax(derr), derr.shape)
        return False, "distance mismatch at query %d rank %d: %r vs oracle %r (rel %.3g)" % (
            q, r, dist[q, r], ref_dist[q, r], derr[q, r])
    bad = np.argwhere(idx != ref_idx)
    for q, r in bad:
        # allowed only if the distance of what we returned ties the oracle's at that rank
        a, b = dist[q, r], ref_dist[q, r]
        if data is not None and query is not None:
            a = float(pair_dist(data, query, [q], [idx[q, r]])[0])
        if abs(a - b) > tie_rtol * max(abs(b), 1e-300):
            return False, "index mismatch at query %d rank %d: %d (d=%r) vs oracle %d (d=%r)" % (
                q, r, idx[q, r], a, ref_idx[q, r], b)
    return True, "ok (%d tie-excused index differences)" % len(bad)
