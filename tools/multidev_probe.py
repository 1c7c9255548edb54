"""Single-process multi-device handle (DCI(devices=[...])) at C3 scale with NumPy inputs (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200 import DCI, load_library
ng = load_library().b200knn_device_count()
n, q, d = 300000, 30000, 3072
rng = np.random.default_rng(0)
pool = np.empty((n, d), dtype=np.float64)
for i in range(0, n, 20000):
    pool[i:i + 20000] = rng.standard_normal((min(20000, n - i), d), dtype=np.float32)
queries = rng.standard_normal((q, d), dtype=np.float32).astype(np.float64)
ref = None
for g in sorted(set([1, 2, min(4, ng), ng])):
    if g > ng: continue
    db = DCI(d, 3, 15, devices=list(range(g)))
    for rep in range(2):
        db.reset()
        t = time.time(); db.add(pool); ta = time.time() - t
        t = time.time(); idx, dist = db.query_arrays(queries, 1); tq = time.time() - t
    if ref is None: ref = (idx, dist)
    same = bool(np.array_equal(idx, ref[0]) and np.array_equal(dist, ref[1]))
    print("devices=%d: add %.3f s, query %.1f ms (%.0f q/s), identical to 1-device result: %s" % (g, ta, tq * 1e3, q / tq, same)); sys.stdout.flush()
    del db
