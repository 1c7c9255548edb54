"""Throughput of the on-device random projection beside NumPy's float64 matmul on the host.
usage: python tools/project_probe.py [rows] [in_dim] [proj_dim]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from inclusivegan_b200 import DCI  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
in_dim = int(sys.argv[2]) if len(sys.argv) > 2 else 49152
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
rng = np.random.default_rng(0)
rows = rng.standard_normal((n, in_dim), dtype=np.float32)
proj = rng.normal(0.0, 1.0 / dim, size=(in_dim, dim))
db = DCI(dim)
t0 = time.perf_counter()
db.set_projector(proj)
t_set = time.perf_counter() - t0
db.add_projected(rows[:512])       # warm-up (allocations, module load)
db.reset()
t0 = time.perf_counter()
db.add_projected(rows)
t_add = time.perf_counter() - t0
flop = 2.0 * n * in_dim * dim
print("set_projector %.2f s (%.1f GB); add_projected %d x %d -> %d: %.3f s = %.1f TFLOP/s float64 incl. upload + convert"
      % (t_set, proj.nbytes / 1e9, n, in_dim, dim, t_add, flop / t_add / 1e12))
t0 = time.perf_counter()
idx, dist = db.query_projected_arrays(rows[:24], 1)
t_q = time.perf_counter() - t0
print("query_projected 24 rows: %.2f ms; self-match %s" % (1e3 * t_q, bool((idx[:, 0] == np.arange(24)).all())))
m = 256                               # the trainer's chunk (training_loop.py:362-365)
t0 = time.perf_counter()
ref = rows[:m].astype(np.float64) @ proj
t_np = time.perf_counter() - t0
print("NumPy float64 matmul of one %d-row chunk: %.3f s = %.2f TFLOP/s -> %d rows would take %.1f s"
      % (m, t_np, 2.0 * m * in_dim * dim / t_np / 1e12, n, t_np * n / m))
got = db.project_rows(rows[:m])
print("max |device - numpy| / scale = %.2e" % (np.max(np.abs(got - ref)) / np.max(np.abs(ref))))
