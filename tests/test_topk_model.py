"""CPU model of scan_topk_kernel (inclusivegan_b200/csrc/scan.cuh): the k smallest float64 keys of a row, ascending by
(key, index), found by an 8-pass radix select on the bit patterns, an index-ordered compaction in 1024-key chunks, and a
stable LSD radix sort of the k selected pairs.  The model follows the kernel's arithmetic step by step (same chunking,
same prefix-sum bookkeeping, same skipped passes) and is checked against np.lexsort — it pins the algorithm the GPU
tests (`test_large_k_select_and_sort`, `test_large_k_ties_come_out_in_index_order`) exercise on the device."""
import numpy as np
import pytest

CHUNK = 1024


def radix_select(keys, kk):
    """phase A: the kk-th smallest key T and how many keys == T belong to the answer (lowest indices first)"""
    prefix, need = np.uint64(0), kk
    for p in range(7, -1, -1):
        hi_shift = np.uint64(8 * (p + 1))
        if p < 7:
            live = (keys >> hi_shift) == (prefix >> hi_shift)
        else:
            live = np.ones(keys.shape, dtype=bool)
        digit = ((keys >> np.uint64(8 * p)) & np.uint64(255)).astype(np.int64)
        hist = np.bincount(digit[live], minlength=256)
        b, rem = 0, need
        while b < 255 and hist[b] < rem:
            rem -= hist[b]
            b += 1
        prefix |= np.uint64(b) << np.uint64(8 * p)
        need = rem
    return prefix, need


def compact(keys, kk, t, need):
    """phase B: chunks of 1024 keys in index order; packed (less, eq) exclusive prefix sums per chunk"""
    n = len(keys)
    out_k = np.zeros(kk, dtype=np.uint64)
    out_i = np.zeros(kk, dtype=np.int64)
    out_base, eq_seen = 0, 0
    for j0 in range(0, n, CHUNK):
        kc = keys[j0:j0 + CHUNK]
        less = np.ones(len(kc), dtype=bool) if kk == n else kc < t
        eq = np.zeros(len(kc), dtype=bool) if kk == n else kc == t
        less_before = np.cumsum(less) - less
        eq_before = eq_seen + np.cumsum(eq) - eq
        take = less | (eq & (eq_before < need))
        pos = out_base + less_before + (np.minimum(eq_before, need) - min(eq_seen, need))
        out_k[pos[take]] = kc[take]
        out_i[pos[take]] = j0 + np.nonzero(take)[0]
        eq_total = eq_seen + int(eq.sum())
        out_base += int(less.sum()) + (min(eq_total, need) - min(eq_seen, need))
        eq_seen = eq_total
    assert out_base == kk
    return out_k, out_i


def lsd_sort(k, i):
    """phase C: 8-bit digits from the bottom; a pass whose digit every key shares is skipped; stable within a pass because
    chunks, warps and lanes are visited in index order"""
    passes = 0
    for p in range(8):
        digit = ((k >> np.uint64(8 * p)) & np.uint64(255)).astype(np.int64)
        hist = np.bincount(digit, minlength=256)
        if hist.max() == len(k):
            continue
        passes += 1
        base = np.cumsum(hist) - hist
        dst_k, dst_i = np.empty_like(k), np.empty_like(i)
        for c0 in range(0, len(k), CHUNK):
            d = digit[c0:c0 + CHUNK]
            for w0 in range(0, len(d), 32):                       # a warp: rank among the lanes with the same digit
                dw = d[w0:w0 + 32]
                for lane, dg in enumerate(dw):
                    rank = int(np.sum(dw[:lane] == dg))
                    dst = base[dg] + rank
                    dst_k[dst] = k[c0 + w0 + lane]
                    dst_i[dst] = i[c0 + w0 + lane]
                for dg in np.unique(dw):
                    base[dg] += int(np.sum(dw == dg))
        k, i = dst_k, dst_i
    return k, i, passes


def topk_model(d2, kk):
    keys = d2.view(np.uint64)
    n = len(keys)
    t, need = (np.uint64(0xFFFFFFFFFFFFFFFF), 0) if kk == n else radix_select(keys, kk)
    ck, ci = compact(keys, kk, t, need)
    sk, si, passes = lsd_sort(ck, ci)
    return sk.view(np.float64), si, passes


@pytest.mark.parametrize("n,kk,kind", [(5000, 100, "gauss"), (3000, 3000, "gauss"), (2049, 2048, "gauss"), (1500, 700, "lattice"),
                                       (1500, 1500, "lattice"), (4100, 33, "lattice"), (1024, 1, "const"), (2500, 1300, "const")])
def test_model_equals_lexsort(n, kk, kind):
    rng = np.random.default_rng(n + kk)
    if kind == "gauss":
        d2 = (rng.standard_normal(n) ** 2 * 1e3 + 4000.0)              # distances of one query: same top bytes, as in the kernel's case
    elif kind == "lattice":
        d2 = rng.integers(0, 40, size=n).astype(np.float64)             # many exact ties, also across the selection threshold
    else:
        d2 = np.full(n, 7.25)
    got_d, got_i, passes = topk_model(np.ascontiguousarray(d2), kk)
    order = np.lexsort((np.arange(n), d2))[:kk]
    assert np.array_equal(got_i, order)
    assert np.array_equal(got_d, d2[order])
    if kind == "const":
        assert passes == 0                                              # every digit shared: nothing to sort
