"""compute-sanitizer target for the kernels and host paths of round 2's second session: the warp-per-query re-rank
(persistent, work counter), the k > 32 select + sort kernel, one second pass per call over whole-call buffers (several
chunks from host rows, self-kNN beyond one pass), the upload ramp."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200KNN_RERANK_WARP"] = "2"            # the warp flavour also for small batches
os.environ["B200KNN_UPLOAD_RAMP"] = "2"
from inclusivegan_b200 import DCI
rng = np.random.default_rng(0)
x = rng.standard_normal((3000, 200)); y = rng.standard_normal((1300, 200))
db = DCI(200); db.add(x)
for k in (1, 10, 20):
    db.query_arrays(y, k)                          # warp re-rank, C = 16 / 32 / 64; 1300 rows = several ramp chunks, one second pass
db.query_arrays(y[:24], 1)
xf = rng.standard_normal((2500, 129)).astype(np.float32); dbf = DCI(129); dbf.add(xf); dbf.query_arrays(xf[:600], 4, squared=True)
print("warp re-rank done", db.stats()["uncertified"])
# k > 32: select + sort (k < n over several 1024-key chunks, k = n, ties)
db.query_arrays(y[:9], 100); db.query_arrays(y[:5], 3000); db.query_arrays(y[:5], 2999)
lat = rng.integers(-2, 3, size=(1500, 6)).astype(np.float64); dbt = DCI(6); dbt.add(lat); dbt.query_arrays(lat[:7], 700)
print("large k done")
# uncertified rows in several chunks of one call -> one collection pass at the end
hx = rng.standard_normal((2000, 2048)); hy = rng.standard_normal((1100, 2048))
dbh = DCI(2048); dbh.add(hx); dbh.query_arrays(hy, 5)
print("accumulated second pass done", dbh.stats()["uncertified"])
