"""Generate tests/golden/pr_*.npz from the UNMODIFIED reference metrics/precision_recall.py (build container only).

    python tests/golden/make_golden_pr.py

The reference's ManifoldEstimator (metrics/precision_recall.py:60-134) takes its distance block as a constructor
argument, so it runs without TensorFlow: `tensorflow`, `dnnlib`, `metrics.metric_base` and `training.misc` are stubbed
in sys.modules (the module only needs them to import), and a NumPy block stands in for DistanceBlock (:38-57, a TF
fp16 matmul).  The class itself stores distances and radii in float16 whatever the block returns (:72-73, :100), so

  pr_lattice.npz   small-integer features: every squared distance is an integer <= 2048, exact in float16 — the
                   reference's output IS the real-number answer; pins oracle/pr_oracle.py (float64 mode) and, on the
                   GPU, inclusivegan_b200.precision_recall directly against reference outputs.
  pr_generic.npz   non-negative Gaussian-like features: float16 rounding matters; pins the oracle's
                   `store_dtype=float16` mode (the same arithmetic with the reference's storage casts) bit for bit,
                   which ties the float64 definition the B200 path implements to the reference's code path.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference_pr():
    for name in ("tensorflow", "dnnlib", "dnnlib.tflib", "metrics", "metrics.metric_base", "training", "training.misc"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["dnnlib"].tflib = sys.modules["dnnlib.tflib"]
    sys.modules["dnnlib"].EasyDict = dict
    sys.modules["metrics"].metric_base = sys.modules["metrics.metric_base"]
    sys.modules["metrics.metric_base"].MetricBase = object            # base class of PR (:172), never instantiated here
    sys.modules["training"].misc = sys.modules["training.misc"]
    spec = importlib.util.spec_from_file_location("_reference_precision_recall", os.path.join(REF, "metrics", "precision_recall.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class NumpyDistanceBlock(object):
    """Stand-in for DistanceBlock.pairwise_distances (:54-56): squared Euclidean distances, float64, direct differences."""

    def pairwise_distances(self, U, V):
        U = np.asarray(U, dtype=np.float64)
        V = np.asarray(V, dtype=np.float64)
        out = np.empty((U.shape[0], V.shape[0]))
        for i in range(U.shape[0]):
            diff = V - U[i]
            out[i] = np.einsum("ij,ij->i", diff, diff)
        return out


def cases():
    rng = np.random.default_rng(20260202)
    out = {}
    # integer lattice: coordinates in {0,1,2,3}, d = 24 -> squared distances <= 24 * 9 = 216, exact in float16
    ref = rng.integers(0, 4, (700, 24)).astype(np.float32)
    ev = np.concatenate([rng.integers(0, 4, (400, 24)), rng.integers(0, 7, (200, 24))]).astype(np.float32)
    out["pr_lattice"] = (ref, ev, [1, 3, 5], 128, 256)
    w = rng.standard_normal((6, 96)) / np.sqrt(6)
    ref = np.maximum(rng.standard_normal((600, 6)) @ w + 0.05 * rng.standard_normal((600, 96)), 0).astype(np.float32)
    ev = np.maximum(1.3 * rng.standard_normal((500, 6)) @ w + 0.05 * rng.standard_normal((500, 96)), 0).astype(np.float32)
    out["pr_generic"] = (ref, ev, [3], 100, 250)
    return out


def main():
    mod = import_reference_pr()
    block = NumpyDistanceBlock()
    for name, (ref, ev, nhoods, rb, cb) in cases().items():
        ref_m = mod.ManifoldEstimator(block, ref, rb, cb, nhoods)
        ev_m = mod.ManifoldEstimator(block, ev, rb, cb, nhoods)
        with np.errstate(divide="ignore", invalid="ignore"):
            precision, realism, nearest = ref_m.evaluate(ev, return_realism=True, return_neighbors=True)
            recall = ev_m.evaluate(ref)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), ref=ref, ev=ev, nhood_sizes=np.array(nhoods), row_batch=rb, col_batch=cb,
                            D_ref=ref_m.D, D_ev=ev_m.D, precision=precision, realism=realism, nearest=nearest, recall=recall)
        print(name, "D dtype", ref_m.D.dtype, "precision", precision.mean(axis=0), "recall", recall.mean(axis=0))


if __name__ == "__main__":
    main()
