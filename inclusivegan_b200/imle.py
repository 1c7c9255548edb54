"""Host-side helpers for the IMLE matching step of the trainer (training/training_loop.py:353-406).

The reference loops over the real images 2*minibatch rows at a time (:374-403), one `dci_db.query` per 24 rows, and
in exclusive mode scans a growing Python list per candidate (`idx not in nearest_indices_list`, :386-396 — O(Q^2)).
`match_all` is the batched equivalent: ONE query for all rows (tensor-bound instead of 1 250 HBM-bound calls), then
the same first-unused-neighbour rule with a set.  The result is what the loop would have produced for the same
exact neighbours, in the same row order.
"""
import numpy as np


def exclusive_assign(indices, dists):
    """Greedy exclusive assignment of training_loop.py:386-396.

    indices / dists: [Q, k] neighbours of each real image, ascending distance.  Row i takes its first neighbour that no
    earlier row has taken; if all k are taken it falls back to its nearest (index 0) — exactly the reference's rule,
    including that fallback picks may repeat.  Returns (chosen index [Q], chosen distance [Q])."""
    indices = np.asarray(indices)
    dists = np.asarray(dists)
    q, k = indices.shape
    taken = set()
    out_i = np.empty(q, dtype=indices.dtype)
    out_d = np.empty(q, dtype=dists.dtype)
    for i in range(q):
        row = indices[i]
        pick = 0
        for j in range(k):
            if int(row[j]) not in taken:
                pick = j
                break
        out_i[i] = row[pick]
        out_d[i] = dists[i, pick]
        taken.add(int(row[pick]))
    return out_i, out_d


def match_all(dci_db, reals, exclusive_retrieved_code=False, num_samples_factor=10):
    """All real feature rows against the indexed pool in one call (replaces the while-loop at :374-403).

    reals: [data_size, dim] features in the order the loop would have visited them.
    Returns (nearest_indices [data_size], nearest_dists [data_size]) as arrays; `latent_candidates[nearest_indices]`
    and `np.percentile(nearest_dists, pct)` then follow as in :404-406."""
    k = int(num_samples_factor) if exclusive_retrieved_code else 1
    if hasattr(dci_db, "query_arrays"):
        idx, dist = dci_db.query_arrays(reals, k)
    else:                                   # any object with the reference's DCI.query signature
        li, ld = dci_db.query(reals, num_neighbours=k, field_of_view=200, prop_to_retrieve=1.0)
        idx, dist = np.array(li), np.array(ld)
    if exclusive_retrieved_code:
        return exclusive_assign(idx, dist)
    return idx[:, 0].copy(), dist[:, 0].copy()
