"""k-NN precision / recall (Kynkäänniemi et al.) on the B200 exact-kNN engine.

Mirror of the kNN part of the reference's metrics/precision_recall.py (ManifoldEstimator :60-134,
knn_precision_recall_features :138-167) with the same names, arguments and result fields, so
`PR._evaluate` (:180-223) can call it unchanged apart from the import.  What changes underneath:

  * the reference fills `[row_batch, N]` float16 host matrices from a TF fp16 matmul
    (DistanceBlock :38-57) and takes `np.partition` / `np.any` / `np.argmin` over them; here the k-th
    neighbour radii are a self-kNN (`k+1` smallest incl. the point itself, squared distances, :74-90) and the
    manifold test is `b200knn_ball_membership` — both tensor-core filtered and decided exactly in float64,
    the N x N distance matrix never exists.
  * results are exact; the reference's are float16-rounded (`self.D` is float16, :72), so counts can differ
    for points within float16 resolution of a ball surface.  `feature_net`, `row_batch_size`,
    `col_batch_size`, `num_gpus` are accepted for signature compatibility and unused.

No TensorFlow; features are NumPy arrays (float32 as the reference produces, or float64).
"""
import numpy as np

from .dci import DCI


class _State(dict):
    """Attribute-access dict standing in for dnnlib.EasyDict (precision_recall.py:141)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class ManifoldEstimator(object):
    """Estimate of the manifold of `features`: one ball per sample, radius = distance to its k-th neighbour."""

    def __init__(self, distance_block, features, row_batch_size=None, col_batch_size=None, nhood_sizes=(3,),
                 clamp_to_percentile=None, devices=None):
        features = np.ascontiguousarray(features)
        if features.dtype not in (np.float32, np.float64):
            features = features.astype(np.float32)
        if features.base is not None:
            features = features.copy()
        self.nhood_sizes = list(nhood_sizes)
        self.num_nhoods = len(self.nhood_sizes)
        self._ref_features = features
        self._index = DCI(features.shape[1], devices=devices)
        self._index.add(features)
        # k-th nearest neighbour of each sample among the samples themselves: index 0 is the sample itself
        # (precision_recall.py:73-90, `np.partition(..., seq)[:, nhood_sizes]` with seq = 0..max(k))
        kmax = max(self.nhood_sizes) + 1
        _, d2 = self._index.query_self_arrays(min(kmax, features.shape[0]), squared=True)   # rows are already on the device
        cols = [min(k, d2.shape[1] - 1) for k in self.nhood_sizes]
        self.D = np.ascontiguousarray(d2[:, cols])                       # float64 (reference: float16)
        if clamp_to_percentile is not None:                              # precision_recall.py:92-94
            max_distances = np.percentile(self.D, clamp_to_percentile, axis=0)
            self.D[self.D > max_distances] = 0

    def evaluate(self, eval_features, return_realism=False, return_neighbors=False):
        """Are the new feature vectors inside the estimated manifold?  (precision_recall.py:96-134)"""
        eval_features = np.ascontiguousarray(eval_features)
        if eval_features.dtype not in (np.float32, np.float64):
            eval_features = eval_features.astype(np.float32)
        num_eval = eval_features.shape[0]
        batch_predictions = np.zeros([num_eval, self.num_nhoods], dtype=np.int32)
        for j in range(self.num_nhoods):
            batch_predictions[:, j] = self._index.ball_membership(eval_features, self.D[:, j])
        if not (return_realism or return_neighbors):
            return batch_predictions
        idx, d2 = self._index.query_arrays(eval_features, 1, squared=True)
        nearest_indices = idx[:, 0].astype(np.int32)
        with np.errstate(divide="ignore", invalid="ignore"):
            realism_score = (self.D[nearest_indices, 0] / d2[:, 0]).astype(np.float32)      # :125
        if return_realism and return_neighbors:
            return batch_predictions, realism_score, nearest_indices
        if return_realism:
            return batch_predictions, realism_score
        return batch_predictions, nearest_indices


def knn_precision_recall_features(ref_features, eval_features, feature_net=None, nhood_sizes=(3,),
                                  row_batch_size=None, col_batch_size=None, num_gpus=None, devices=None):
    """k-NN precision and recall of eval_features w.r.t. ref_features (precision_recall.py:138-167)."""
    state = _State()
    state.ref_features = ref_features
    state.eval_features = eval_features
    state.ref_manifold = ManifoldEstimator(None, ref_features, row_batch_size, col_batch_size, nhood_sizes, devices=devices)
    state.eval_manifold = ManifoldEstimator(None, eval_features, row_batch_size, col_batch_size, nhood_sizes, devices=devices)
    # precision: how many eval points are in the ref manifold
    state.precision, state.realism_scores, state.nearest_neighbors = state.ref_manifold.evaluate(
        eval_features, return_realism=True, return_neighbors=True)
    state.knn_precision = state.precision.mean(axis=0)
    # recall: how many ref points are in the eval manifold
    state.recall = state.eval_manifold.evaluate(ref_features)
    state.knn_recall = state.recall.mean(axis=0)
    return state
