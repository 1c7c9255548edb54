"""Real-world drop-in path (development aid): DCI.add / DCI.query with pageable NumPy float64 arrays at C3 scale."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200 import DCI
n, q, d = 300000, 30000, 3072
rng = np.random.default_rng(0)
t = time.time()
pool = np.empty((n, d), dtype=np.float64)
for i in range(0, n, 20000):
    pool[i:i + 20000] = rng.standard_normal((min(20000, n - i), d), dtype=np.float32)
queries = rng.standard_normal((q, d), dtype=np.float32).astype(np.float64)
print("generated in %.1fs" % (time.time() - t)); sys.stdout.flush()
for threads in (8, 1):
    os.environ["B200KNN_COPY_THREADS"] = str(threads)
    db = DCI(d, 3, 15)
    for rep in range(3):
        db.reset()
        t = time.time(); db.add(pool, num_levels=3, field_of_view=10, prop_to_retrieve=0.002); ta = time.time() - t
        t = time.time(); idx, dist = db.query(queries, num_neighbours=1, field_of_view=200, prop_to_retrieve=1.0); tq = time.time() - t
        t = time.time(); i2, d2 = db.query_arrays(queries, 1); tq2 = time.time() - t
        print("copy_threads=%d rep %d: add %.3f s (%.1f GB/s)  query(list API) %.1f ms  query_arrays %.1f ms (%.0f q/s)" % (
            threads, rep, ta, pool.nbytes / ta / 1e9, tq * 1e3, tq2 * 1e3, q / tq2)); sys.stdout.flush()
    # trainer-granularity calls: 24 rows per call
    t = time.time()
    for s0 in range(0, 24 * 200, 24):
        db.query(queries[s0:s0 + 24], num_neighbours=1, field_of_view=200, prop_to_retrieve=1.0)
    print("   24-row calls: %.3f ms/call" % ((time.time() - t) / 200 * 1e3))
    del db
