"""Host-side row-sharding arithmetic for multi-GPU matching (one process per GPU).

The pool is split into contiguous row ranges, rank g owning rows [g*ceil(N/G), (g+1)*ceil(N/G)) — the same
rule libb200knn uses inside b200knn_add for multi-device handles — so a shard-local row index plus
`index_base` (the range start) is the global row index (the reference's data_idx_offset, py_dci.c:185).
Queries are replicated; each rank produces an exact local top-k; the lists are all-gathered as
[G][Q][kk] (index int32, distance float64) and merged by b200knn_merge_topk_device.
"""


def shard_range(n_rows, world_size, rank):
    """Contiguous [start, stop) of pool rows owned by `rank`."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    per = (n_rows + world_size - 1) // world_size
    return min(n_rows, per * rank), min(n_rows, per * (rank + 1))


def pad_local_topk(idx, dist, kk):
    """Pad a shard's [Q, k_local] lists (k_local < kk when the shard holds fewer than kk rows) to [Q, kk] with
    (-1, +inf): the empty-slot encoding the merge kernel skips."""
    import numpy as np
    q, kl = idx.shape
    if kl == kk:
        return idx, dist
    pi = np.full((q, kk), -1, dtype=np.int32)
    pd = np.full((q, kk), np.inf, dtype=np.float64)
    pi[:, :kl] = idx
    pd[:, :kl] = dist
    return pi, pd
