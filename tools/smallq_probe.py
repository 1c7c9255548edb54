"""Trainer-granularity latency (development aid): 24-row query() calls against a 300k x 3072 pool, + PR metric at 50k."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200 import DCI
rng = np.random.default_rng(0)
n, d = 300000, 3072
pool = rng.standard_normal((n, d), dtype=np.float32)
queries = rng.standard_normal((4800, d), dtype=np.float32).astype(np.float64)
db = DCI(d, 3, 15); db.add(pool)
for rows in (24, 12, 64, 256):
    for _ in range(5): db.query(queries[:rows], num_neighbours=1)
    db.set_profiling(True); db._lib.b200knn_reset_stats(db._handle)
    t = time.time(); reps = 200
    for s0 in range(0, rows * reps, rows):
        s0 %= (4800 - rows)
        db.query(queries[s0:s0 + rows], num_neighbours=1, field_of_view=200, prop_to_retrieve=1.0)
    ms = (time.time() - t) / reps * 1e3
    st = db.stats(); db.set_profiling(False)
    print("%3d-row calls: %.3f ms/call (%.0f q/s) | kernels per call: convert %.3f dist %.3f rerank %.3f second %.3f ms, uncert/call %.2f" % (
        rows, ms, rows / ms * 1e3, st["ms_convert"] / reps, st["ms_distance"] / reps, st["ms_rerank"] / reps, st["ms_scan"] / reps, st["uncertified"] / reps))
del db
# PR metric at config-4 scale
from inclusivegan_b200.precision_recall import knn_precision_recall_features
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_pr_metric import two_sets
ref, ev = two_sets(50000, 50000, 2048, seed=1)
for rep in range(2):
    t = time.time(); st = knn_precision_recall_features(ref, ev, nhood_sizes=[3]); dt = time.time() - t
    print("PR metric 50k+50k x 2048, k=3: %.3f s  precision %.4f recall %.4f" % (dt, st.knn_precision[0], st.knn_recall[0]))
