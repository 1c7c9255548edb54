import numpy as np, sys
sys.path.insert(0, ".")
from inclusivegan_b200 import DCI
rng = np.random.default_rng(0)
xs = rng.standard_normal((20000, 64)); dbs = DCI(64); dbs.add(xs)
i, d = dbs.query_arrays(rng.standard_normal((24, 64)), 10)
xf = rng.standard_normal((20000, 3300)).astype(np.float32); dbf = DCI(3300); dbf.add(xf); dbf.query_arrays(xf[:8] + 0.01, 10)
print("mini done", dbs.stats()["uncertified"], dbf.stats()["uncertified"])
