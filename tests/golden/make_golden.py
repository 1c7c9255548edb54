"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Needs /root/reference (for dci_code/src/dci.py) and oracle/_ref/_dci.so (make -C oracle ref).  The GPU box
has neither the reference nor this need: tests read only the committed .npz files.

What is recorded
  knn_*.npz      inputs (float64) + the reference's answers in EXHAUSTIVE mode (num_levels=1,
                 prop_to_visit=prop_to_retrieve=1.0: Prioritized DCI visits and retrieves every point, so its
                 output is the exact kNN) through the reference's own Python wrapper dci.py -> _dci -> dci.c.
                 These pin oracle/knn_oracle.{c,py} (bit-identical distances expected: same loop as util.c:62-69).
  approx_*.npz   a deterministic APPROXIMATE run of the reference (num_levels=1 so no drand48 level assignment,
                 proj_vec overwritten with seeded values as dci.py:97-105 allows): pins the harness
                 oracle/ref_dci.py + the prebuilt _dci.so that bench.py times on the GPU box.
  select_rows.npz  the reference's index-selection semantics (dci.py:146-221 _check_and_fix_indices) on a table
                 of selector cases: pins inclusivegan_b200.dci.DCI._select_rows.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_dci  # noqa: E402


def gen_lowrank(rng, n, d, intrinsic=50):
    # dci_code/example.py:36-40
    latent = 2 * rng.random((n, intrinsic)) - 1
    t = 2 * rng.random((intrinsic, d)) - 1
    return latent @ t


def cases():
    rng = np.random.default_rng(20260101)
    out = {}
    x = rng.standard_normal((512, 40)); q = rng.standard_normal((33, 40))
    out["knn_gauss"] = (x, q, 5)
    allpts = gen_lowrank(rng, 416, 100)
    out["knn_lowrank"] = (np.copy(allpts[:400]), np.copy(allpts[400:]), 10)
    x = rng.standard_normal((150, 16)); x = np.concatenate([x, x[:150:3]], axis=0)      # exact duplicates
    q = np.concatenate([x[:10], rng.standard_normal((10, 16))], axis=0)                 # zero-distance queries
    out["knn_ties"] = (np.copy(x), np.copy(q), 4)
    x = rng.standard_normal((7, 16)); q = rng.standard_normal((5, 16))
    out["knn_k_gt_n"] = (x, q, 10)
    x = np.clip(0.5 * rng.standard_normal((300, 129)), -1, 1); q = np.clip(0.5 * rng.standard_normal((17, 129)), -1, 1)
    out["knn_image_odd_dim"] = (x, q, 1)
    return out


def main():
    mod = ref_dci.import_reference_wrapper()
    for name, (x, q, k) in cases().items():
        db = mod.DCI(x.shape[1], 2, 7)
        db.add(x, num_levels=1, prop_to_visit=1.0, prop_to_retrieve=1.0)
        idx, dist = db.query(q, num_neighbours=k, prop_to_visit=1.0, prop_to_retrieve=1.0)
        kk = min(k, x.shape[0])
        assert all(len(a) == kk for a in idx), "exhaustive reference must return min(k, N) per query"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), data=x, query=q, k=k,
                            ref_idx=np.array(idx, dtype=np.int32), ref_dist=np.array(dist, dtype=np.float64))
        db.clear()
        print(name, x.shape, q.shape, k)

    # deterministic approximate run
    rng = np.random.default_rng(7)
    x = rng.standard_normal((600, 32)); q = rng.standard_normal((12, 32))
    db = mod.DCI(32, 2, 7)
    pv = rng.standard_normal(db.proj_vec.shape); pv /= np.linalg.norm(pv, axis=1, keepdims=True)
    db.proj_vec = pv
    db.add(x, num_levels=1, prop_to_retrieve=0.05)
    idx, dist = db.query(q, num_neighbours=5, prop_to_retrieve=0.05)
    counts = np.array([len(a) for a in idx], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "approx_levels1.npz"), data=x, query=q, k=5, proj_vec=pv,
                        flat_idx=np.concatenate(idx).astype(np.int32), flat_dist=np.concatenate(dist), counts=counts)
    print("approx_levels1", counts)

    # index-selection semantics of the reference wrapper
    db = mod.DCI(4, 2, 7)
    data = np.zeros((10, 4))
    sel_cases = [
        ("none", None), ("slice_0_10", slice(0, 10)), ("slice_2_7", slice(2, 7)), ("slice_neg", slice(-4, -1)),
        ("slice_step2", slice(1, 9, 2)), ("slice_over", slice(5, 100)), ("int_3", 3), ("int_neg1", -1),
        ("arr_intc", np.array([5, 1, 8], dtype=np.intc)), ("arr_int64_neg", np.array([0, -1, 4], dtype=np.int64)),
        ("arr_bool", np.arange(10) % 3 == 0), ("list_int", [9, 0, 2]), ("list_bool", [True, False] * 5),
    ]
    rec = {}
    for nm, sel in sel_cases:
        contig, val = db._check_and_fix_indices(data, sel)
        rec[nm + "__contig"] = np.array(bool(contig))
        rec[nm + "__val"] = np.asarray(val, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "select_rows.npz"), **rec)
    print("select_rows", len(sel_cases))


if __name__ == "__main__":
    main()
