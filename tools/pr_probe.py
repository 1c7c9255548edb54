import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from inclusivegan_b200.precision_recall import knn_precision_recall_features, ManifoldEstimator
from test_pr_metric import two_sets
ref, ev = two_sets(50000, 50000, 2048, seed=1)
for rep in range(5):
    t = time.time(); st = knn_precision_recall_features(ref, ev, nhood_sizes=[3]); dt = time.time() - t
    print("PR metric 50k+50k x 2048, k=3: %.3f s  precision %.4f recall %.4f" % (dt, st.knn_precision[0], st.knn_recall[0])); sys.stdout.flush()
t = time.time(); m = ManifoldEstimator(None, ref, nhood_sizes=[3]); t1 = time.time() - t
t = time.time(); p = m.evaluate(ev); t2 = time.time() - t
t = time.time(); p = m.evaluate(ev, return_realism=True, return_neighbors=True); t3 = time.time() - t
print("breakdown: estimator (add + self-kNN) %.3f s, membership %.3f s, membership+realism+neighbours %.3f s; stats %s" % (t1, t2, t3, m._index.stats()))
