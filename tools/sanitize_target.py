"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200 import DCI
from inclusivegan_b200.dci import FLAG_FORCE_SCAN
rng = np.random.default_rng(0)
x = rng.standard_normal((3000, 200)); y = rng.standard_normal((300, 200))
db = DCI(200); db.add(x)
db.query_arrays(y, 1); db.query_arrays(y, 10); db.query_arrays(y[:24], 1); db.query_arrays(y, 3, flags=FLAG_FORCE_SCAN); db.query_arrays(y[:40], 40)
base = 40.0 + rng.standard_normal((1, 64)); xx = base + 1e-3 * rng.standard_normal((2000, 64)); yy = base + 1e-3 * rng.standard_normal((30, 64))
db2 = DCI(64); db2.add(np.ascontiguousarray(xx)); db2.query_arrays(np.ascontiguousarray(yy), 3)      # uncertified -> collect -> overflow -> scan
r2 = np.full(3000, 150.0); db.ball_membership(y, r2)
xf = rng.standard_normal((1500, 129)).astype(np.float32); dbf = DCI(129); dbf.add(xf); dbf.query_arrays(xf[:100], 4, squared=True)
print("sanitize target done", db.stats()["kernel_launches"], db2.stats())
# later additions: self-kNN, long rows (batched canonical sums), forced grid schedule (round-wide lockstep), random projection
db.query_self_arrays(4)
xl = rng.standard_normal((600, 3300)).astype(np.float32); dbl = DCI(3300); dbl.add(xl); dbl.query_arrays(xl[:70], 10)
os.environ["B200KNN_WIDE"] = "2"
for cg in ("1", "2"):
    os.environ["B200KNN_CTA_GROUP"] = cg
    dbw = DCI(200); dbw.add(x); dbw.query_arrays(y, 4); dbw.query_arrays(np.vstack([y] * 3), 1)
del os.environ["B200KNN_WIDE"], os.environ["B200KNN_CTA_GROUP"]
P = rng.normal(0, 0.01, size=(777, 129)); dbp = DCI(129); dbp.set_projector(P)
rows = rng.standard_normal((300, 777)).astype(np.float32)
dbp.add_projected(rows); dbp.query_projected_arrays(rows[:50], 3); dbp.project_rows(rows[:33].astype(np.float64))
print("sanitize target (additions) done")
# few queries x many shortlists: the 1024-thread re-rank flavour (group-parallel sweep) and the 32-warp list re-rank
xs = rng.standard_normal((20000, 64)); dbs = DCI(64); dbs.add(xs); dbs.query_arrays(rng.standard_normal((24, 64)), 10)
print("sanitize target (small-call flavours) done", dbs.stats()["uncertified"])
