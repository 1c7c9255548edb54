"""GPU diagnostic: runs the three code paths (exact scan, tensor pass without certificate, full) on a
ladder of shapes and prints mismatch counts against the float64 oracle.  Development aid, not a test."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DCI, FLAG_NO_CERTIFY, FLAG_FORCE_SCAN
from oracle import knn_oracle as ko

def run(N, Q, d, k, dtype=np.float64, kind="gauss", seed=0, paths=("scan", "tensor_nocert", "full")):
    rng = np.random.default_rng(seed)
    if kind == "gauss":
        X = rng.standard_normal((N, d)); Y = rng.standard_normal((Q, d))
    elif kind == "cluster":
        X = rng.standard_normal((N, d)); Y = X[rng.integers(0, N, Q)] + 0.1 * rng.standard_normal((Q, d))
    elif kind == "image":
        X = np.clip(0.5 * rng.standard_normal((N, d)), -1, 1); Y = np.clip(0.5 * rng.standard_normal((Q, d)), -1, 1)
    X = np.ascontiguousarray(X.astype(dtype)); Y = np.ascontiguousarray(Y.astype(dtype))
    t = time.time(); ri, rd = ko.exact_knn_numpy(X, Y, k); t_or = time.time() - t
    db = DCI(d, 2, 7)
    t = time.time(); db.add(X); t_add = time.time() - t
    out = []
    for path in paths:
        flags = {"scan": FLAG_FORCE_SCAN, "tensor_nocert": FLAG_NO_CERTIFY, "full": 0}[path]
        try:
            t = time.time(); i, dd = db.query_arrays(Y, k, flags=flags); tq = time.time() - t
            ok, msg = ko.compare_knn(i, dd, ri, rd, X, Y)
            nbad = int((i != ri).sum())
            maxrel = float(np.max(np.abs(dd - rd) / np.maximum(rd, 1e-300))) if dd.size else 0.0
            out.append("%s: ok=%s idx_mismatch=%d/%d max_rel_dist=%.2e t=%.3fs [%s]" % (path, ok, nbad, i.size, maxrel, tq, msg))
        except Exception as e:
            out.append("%s: EXC %s: %s" % (path, type(e).__name__, e))
            break
    st = db.stats()
    print("N=%d Q=%d d=%d k=%d %s %s | add %.3fs oracle %.2fs | uncert=%d scanned=%d launches=%d" % (
        N, Q, d, k, np.dtype(dtype).name, kind, t_add, t_or, st["uncertified"], st["exact_scanned"], st["kernel_launches"]))
    for o in out: print("    " + o)
    sys.stdout.flush()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "ladder"
    if which == "ladder":
        run(1000, 10, 64, 1)
        run(1000, 10, 64, 5)
        run(3000, 130, 200, 1)
        run(5000, 300, 129, 3)
        run(5000, 64, 1024, 10, dtype=np.float32)
        run(70000, 1000, 512, 1)
        run(70000, 1000, 512, 1, kind="cluster")
        run(20000, 500, 3072, 1, kind="image")
        run(10000, 100, 5000, 10)
        run(300, 20, 64, 40)          # k > 32: segmented sort path
        run(7, 5, 16, 10)             # k > N
        run(100000, 3000, 3072, 1, dtype=np.float32, paths=("full",))
        run(50000, 5000, 2048, 4, dtype=np.float32, kind="image", paths=("full",))
