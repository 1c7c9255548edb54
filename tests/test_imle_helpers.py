"""Host logic of the batched IMLE matching helper against a transcription-free restatement of the trainer's loop."""
import numpy as np

from inclusivegan_b200.imle import exclusive_assign, match_all


def loop_semantics(indices, dists):
    """What training_loop.py:386-396 computes, restated with the reference's own data structure (a growing list and
    `not in`), kept O(Q^2) on purpose."""
    chosen_i, chosen_d = [], []
    for i in range(indices.shape[0]):
        added = False
        for j in range(indices.shape[1]):
            if indices[i, j] not in chosen_i:
                chosen_i.append(indices[i, j]); chosen_d.append(dists[i, j]); added = True
                break
        if not added:
            chosen_i.append(indices[i, 0]); chosen_d.append(dists[i, 0])
    return np.array(chosen_i), np.array(chosen_d)


def test_exclusive_assign_matches_the_loop():
    rng = np.random.default_rng(0)
    for q, k, pool in ((200, 10, 60), (500, 3, 2000), (50, 1, 5), (300, 25, 100)):
        idx = np.stack([rng.choice(pool, size=k, replace=False) if pool >= k else rng.integers(0, pool, k) for _ in range(q)]).astype(np.int32)
        dist = np.sort(rng.random((q, k)), axis=1)
        a_i, a_d = exclusive_assign(idx, dist)
        b_i, b_d = loop_semantics(idx, dist)
        assert np.array_equal(a_i, b_i) and np.array_equal(a_d, b_d)


class _FakeDCI(object):
    """Stands in for a DCI object: answers from a precomputed exact table (no GPU in this test)."""

    def __init__(self, pool):
        self.pool = pool

    def query(self, q, num_neighbours, field_of_view=100, prop_to_retrieve=0.05):
        d = np.linalg.norm(q[:, None, :] - self.pool[None, :, :], axis=2)
        order = np.argsort(d, axis=1, kind="stable")[:, :num_neighbours]
        return [o.astype(np.int32) for o in order], [np.take_along_axis(d, order, axis=1)[i] for i in range(q.shape[0])]


def test_match_all_equals_the_batched_loop():
    rng = np.random.default_rng(1)
    pool = rng.standard_normal((300, 8)); reals = rng.standard_normal((96, 8))
    db = _FakeDCI(pool)
    # non-exclusive: the loop appends idx[:,0] batch by batch (:398-402)
    li, ld = [], []
    for s in range(0, 96, 24):
        i, d = db.query(reals[s:s + 24], num_neighbours=1)
        li += list(np.array(i)[:, 0]); ld += list(np.array(d)[:, 0])
    mi, md = match_all(db, reals)
    assert np.array_equal(mi, np.array(li)) and np.allclose(md, np.array(ld))
    # exclusive with k = 10
    i, d = db.query(reals, num_neighbours=10)
    ei, ed = loop_semantics(np.array(i), np.array(d))
    mi, md = match_all(db, reals, exclusive_retrieved_code=True, num_samples_factor=10)
    assert np.array_equal(mi, ei) and np.allclose(md, ed)
    assert len(set(mi.tolist())) == 96                      # pool is large enough: every real got its own sample
