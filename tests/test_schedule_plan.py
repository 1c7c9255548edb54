"""Invariants of the distance kernel's work schedule (Shard::plan_schedule, reached through b200knn_debug_plan; pure
host code, runs without a GPU): every (query tile, pool tile) pair is computed exactly once, shortlist slots are
consistent, lockstep groups are well-formed — for the L2-resident group schedule and the long-row grid schedule."""
import ctypes

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

BM, BN = 128, 256


def plan(lib, n, nq, kp, num_sms=148, cg=0, max_slots=64, budget=64, wide=1):
    lib.b200knn_debug_plan.restype = ctypes.c_int
    lib.b200knn_debug_plan.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    geo = np.zeros(8, np.int32)
    assert lib.b200knn_debug_plan(n, nq, kp, num_sms, cg, max_slots, budget, wide, None, 0, geo.ctypes.data) == 0, lib.b200knn_last_error()
    cgr, workers, rounds = int(geo[0]), int(geo[1]), int(geo[2])
    items = np.zeros((rounds * workers, 4), np.int32)
    assert lib.b200knn_debug_plan(n, nq, kp, num_sms, cg, max_slots, budget, wide, items.ctypes.data, items.shape[0], geo.ctypes.data) == 0
    keys = ("cg", "workers", "rounds", "qt", "nt", "max_slots", "qg", "wide")
    return dict(zip(keys, (int(v) for v in geo))), items.reshape(rounds, workers, 4)


def check_invariants(g, items, n, nq, max_slots):
    qrows = BM * g["cg"]
    assert g["qt"] == -(-nq // qrows) and g["nt"] == -(-n // BN)
    assert g["max_slots"] <= max_slots
    cover = np.zeros((g["qt"], g["nt"]), np.int32)
    for r in range(g["rounds"]):
        active = items[r][items[r][:, 0] >= 0]
        assert len(active) > 0
        word = active[:, 3].astype(np.uint32)
        slot, sharers, wide = word & 0xFFFF, (word >> 16) & 0xFF, word >> 24
        assert slot.max() < g["max_slots"]
        for (qt, t0, t1, _), sl in zip(active, slot):
            assert 0 <= t0 < t1 <= g["nt"]
            cover[qt, t0:t1] += 1
        # one shortlist slot per (query tile, pool stream): a query tile never meets the same slot twice
        assert len({(int(a[0]), int(s)) for a, s in zip(active, slot)}) == len(active)
        # workers sharing a pool stream sweep identical tile ranges and agree on how many they are
        for s in np.unique(slot):
            grp = active[slot == s]
            assert len({(int(a[1]), int(a[2])) for a in grp}) == 1
            assert (sharers[slot == s] == len(grp)).all()
        if wide.any():            # round-wide lockstep: every active worker carries the same head count, chunk lengths differ by <= 1
            assert (wide == len(active)).all()
            lens = active[:, 2] - active[:, 1]
            assert lens.max() - lens.min() <= 1
            assert len(active) % len(np.unique(slot)) == 0
    assert (cover == 1).all()


def test_headline_shapes(native_lib):
    g, items = plan(native_lib, 300000, 30000, 3072)                 # config 3: L2-resident groups of query tiles
    assert (g["cg"], g["wide"]) == (2, 0) and g["qg"] * BM * 2 * 3072 * 2 <= 64 << 20
    check_invariants(g, items, 300000, 30000, 64)
    g, items = plan(native_lib, 125000, 30000, 49152)                # config 5 share: the grid schedule
    assert g["wide"] == 1 and g["qg"] >= 6 and g["max_slots"] <= 16
    check_invariants(g, items, 125000, 30000, 64)
    g, items = plan(native_lib, 300000, 24, 3072)                    # the trainer's 24-row call: one tile over all SMs
    assert (g["cg"], g["qt"], g["rounds"], g["workers"]) == (1, 1, 1, 148)
    check_invariants(g, items, 300000, 24, 256)
    g, items = plan(native_lib, 125000, 30000, 49152, wide=0)
    assert g["wide"] == 0
    check_invariants(g, items, 125000, 30000, 64)


@settings(max_examples=150, deadline=None)
@given(n=st.integers(1, 400000), nq=st.integers(1, 40000), kp=st.sampled_from([8, 64, 512, 2048, 3072, 5000, 12288, 24576, 49152, 98304]),
       num_sms=st.sampled_from([2, 16, 132, 148]), cg=st.sampled_from([0, 1, 2]), max_slots=st.sampled_from([1, 2, 16, 64, 128, 256]),
       budget=st.sampled_from([1, 26, 64, 104]), wide=st.sampled_from([0, 1, 2]))
def test_every_tile_pair_is_scheduled_exactly_once(native_lib, n, nq, kp, num_sms, cg, max_slots, budget, wide):
    g, items = plan(native_lib, n, nq, kp, num_sms, cg, max_slots, budget, wide)
    check_invariants(g, items, n, nq, max_slots)


# ------------------------------------------------------------------------------------------------ multi-GPU host logic
@settings(max_examples=120, deadline=None)
@given(n=st.integers(1, 400000), nq=st.integers(1, 70000), kp=st.sampled_from([64, 512, 3072, 5000, 49152]),
       cap=st.sampled_from([256, 1000, 4096, 30000, 32768]), world=st.integers(1, 16))
def test_collective_query_chunks_and_slices(native_lib, n, nq, kp, cap, world):
    """b200knn_exchange_query: the chunks tile the query rows in order, none exceeds the exchange's capacity, and the
    per-rank upload slices tile every chunk (the exact re-rank finds the owner of row q as q // slice)."""
    native_lib.b200knn_debug_chunks.restype = ctypes.c_int
    native_lib.b200knn_debug_chunks.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    count = ctypes.c_int64(0)
    assert native_lib.b200knn_debug_chunks(n, nq, kp, 148, cap, world, None, 0, ctypes.byref(count)) == 0
    per = 2 + 2 * world
    out = np.zeros(count.value * per, dtype=np.int64)
    assert native_lib.b200knn_debug_chunks(n, nq, kp, 148, cap, world, out.ctypes.data, out.size, ctypes.byref(count)) == 0
    rows = out.reshape(count.value, per)
    cap_eff = max(256, cap // 256 * 256)
    pos = 0
    for r in rows:
        first, cnt = int(r[0]), int(r[1])
        assert first == pos and 0 < cnt <= cap_eff
        sl = r[2:].reshape(world, 2)
        assert sl[0, 0] == 0 and sl[-1, 1] == cnt and np.all(sl[1:, 0] == sl[:-1, 1]) and np.all(sl[:, 1] >= sl[:, 0])
        slice_rows = -(-cnt // world)
        for q in (0, cnt // 2, cnt - 1):                       # owner lookup of the re-rank
            owner = q // slice_rows
            assert sl[owner, 0] <= q < sl[owner, 1]
        pos += cnt
    assert pos == nq
    if count.value > 1:                                        # the ragged remainder goes first, the rest are equal
        assert len({int(c) for c in rows[1:, 1]}) == 1 and rows[0, 1] <= rows[1, 1]


@pytest.mark.parametrize("n,nq,dim,esz,k,pinned", [(300000, 30000, 3072, 8, 1, 1), (300000, 30000, 3072, 8, 1, 0), (240000, 24000, 3072, 8, 1, 1),
                                                   (300000, 9472, 3072, 8, 1, 1), (50000, 50000, 2048, 4, 4, 1), (10000, 100, 5000, 8, 10, 1),
                                                   (20000, 30011, 512, 8, 25, 1), (300000, 1025, 3072, 4, 1, 0), (1000000, 30000, 49152, 4, 10, 1)])
def test_host_upload_ramp_covers_the_call(native_lib, n, nq, dim, esz, k, pinned):
    """b200knn_query's chunk plan for a single-device host-row call (csrc/b200knn.cu plan_host_chunks): consecutive chunks
    that cover every row once, whole query tiles except the last, bounded by the stage budget; when the call is
    compute-bound (big pool) the first chunk — the only upload nothing hides — is a small share of the call."""
    native_lib.b200knn_debug_host_chunks.restype = ctypes.c_int
    native_lib.b200knn_debug_host_chunks.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    kp = (dim + 7) // 8 * 8
    count = ctypes.c_int64(0)
    assert native_lib.b200knn_debug_host_chunks(n, nq, kp, dim, esz, k, 148, pinned, None, 0, ctypes.byref(count)) == 0
    out = np.zeros(2 * count.value, dtype=np.int64)
    assert native_lib.b200knn_debug_host_chunks(n, nq, kp, dim, esz, k, 148, pinned, out.ctypes.data, out.size, ctypes.byref(count)) == 0
    rows = out.reshape(-1, 2)
    cap = max(256, min(32768, (512 << 20) // (dim * esz) // 256 * 256))
    pos = 0
    for first, cnt in rows.tolist():
        assert first == pos and 0 < cnt <= cap
        pos += cnt
    assert pos == nq
    # whole 256-row query tiles (the 2-CTA kernel's tile) except for one ragged chunk: the last of a ramp, the first of the
    # long-row fallback (remainder first, then whole groups)
    assert sum(1 for c in rows[:, 1].tolist() if c % 256) <= 1
    if n >= 240000 and nq >= 20000 and dim == 3072:             # compute-bound: a ramp, not equal chunks
        assert len(rows) >= 4 and rows[0, 1] <= nq // 10 and rows[0, 1] <= rows[1, 1] <= rows[2, 1]
    # deterministic
    out2 = np.zeros_like(out)
    assert native_lib.b200knn_debug_host_chunks(n, nq, kp, dim, esz, k, 148, pinned, out2.ctypes.data, out2.size, ctypes.byref(count)) == 0
    assert np.array_equal(out, out2)
