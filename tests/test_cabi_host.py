"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/b200knn.h
declares; the Python DCI mirror validates arguments like the reference's dci.py; without a GPU every compute
entry point fails LOUDLY (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "b200knn.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200knn_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(native_lib, s), "libb200knn.so does not export %s" % s
    assert native_lib.b200knn_abi_version() == 1


def test_handle_lifecycle_without_device(native_lib):
    h = ctypes.c_void_p()
    assert native_lib.b200knn_create(64, 0, None, ctypes.byref(h)) == 0
    assert native_lib.b200knn_dim(h) == 64
    assert native_lib.b200knn_num_points(h) == 0
    assert native_lib.b200knn_clear(h) == 0
    # query on an empty index is a state error whatever the machine
    q = np.zeros((2, 64))
    oi = np.zeros((2, 1), np.int32); od = np.zeros((2, 1))
    rc = native_lib.b200knn_query(h, q.ctypes.data, 0, 2, 64, 1, 0, oi.ctypes.data, od.ctypes.data, None)
    assert rc == -2 and b"empty" in native_lib.b200knn_last_error()
    assert native_lib.b200knn_destroy(h) == 0
    bad = ctypes.c_void_p()
    assert native_lib.b200knn_create(0, 0, None, ctypes.byref(bad)) == -1
    assert b"dim" in native_lib.b200knn_last_error()


def test_argument_errors_from_the_c_abi(native_lib):
    h = ctypes.c_void_p()
    assert native_lib.b200knn_create(8, 0, None, ctypes.byref(h)) == 0
    x = np.zeros((4, 8))
    assert native_lib.b200knn_add(h, x.ctypes.data, 7, 4, 8) == -1          # bad dtype code
    assert native_lib.b200knn_add(h, x.ctypes.data, 0, 4, 4) == -1          # ld < dim
    assert native_lib.b200knn_add(h, None, 0, 4, 8) == -1                   # NULL data
    assert native_lib.b200knn_add(h, x.ctypes.data, 0, -1, 8) == -1
    native_lib.b200knn_destroy(h)


def test_no_cpu_fallback(native_lib):
    """On a machine without an sm_100 GPU, add() must fail with ENODEVICE — never compute on the CPU."""
    if native_lib.b200knn_device_count() > 0:
        pytest.skip("a B200 is present; the loud-failure path is exercised on CPU-only boxes")
    from inclusivegan_b200 import DCI, B200KNNError
    db = DCI(8)
    with pytest.raises(B200KNNError) as ei:
        db.add(np.zeros((4, 8)))
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)
    assert db.num_points == 0


def test_collective_and_tier_entry_points_fail_loudly_without_a_gpu(native_lib):
    """The round-2 entry points (precision tiers, the multi-GPU collective protocol) validate their arguments and, like
    everything else, refuse to do anything without a device."""
    from inclusivegan_b200 import dci as mod
    assert mod.PRECISION_TIERS == {"bf16": 0, "bf16x3": 1, "tf32": 2}
    h = ctypes.c_void_p()
    assert native_lib.b200knn_create(8, 0, None, ctypes.byref(h)) == 0
    assert native_lib.b200knn_set_precision(h, 7) == -1 and b"tier" in native_lib.b200knn_last_error()
    assert native_lib.b200knn_set_precision(None, 1) == -1
    ex = ctypes.c_void_p()
    assert native_lib.b200knn_exchange_create_for_queries(0, 0, 1, 0, 256, 1, ctypes.byref(ex)) == -1       # dim must be positive
    assert native_lib.b200knn_exchange_create_for_queries(0, 3, 2, 8, 256, 1, ctypes.byref(ex)) == -1       # rank >= world
    if native_lib.b200knn_device_count() == 0:
        assert native_lib.b200knn_set_precision(h, 1) == -3
        assert native_lib.b200knn_exchange_create_for_queries(0, 0, 1, 8, 256, 1, ctypes.byref(ex)) in (-3, -4)
        assert ex.value is None
    x = np.zeros((4, 8))
    assert native_lib.b200knn_exchange_add(None, h, x.ctypes.data, 0, 4, 8, 0) == -1                       # NULL exchange
    idx = np.zeros((4, 1), np.int32); dist = np.zeros((4, 1))
    assert native_lib.b200knn_exchange_query(None, h, x.ctypes.data, 0, 4, 8, 1, 0, idx.ctypes.data, dist.ctypes.data, None) == -1
    native_lib.b200knn_destroy(h)


def test_missing_library_fails_loudly(tmp_path):
    from inclusivegan_b200 import dci as mod
    with pytest.raises(RuntimeError) as ei:
        mod.load_library(str(tmp_path / "nope.so"))
    assert "no CPU or pure-Python fallback" in str(ei.value)


# ---------------------------------------------------------------- Python mirror of dci.py
def test_constructor_and_properties(native_lib):
    from inclusivegan_b200 import DCI
    db = DCI(3072, num_comp_indices=3, num_simp_indices=15)           # training_loop.py:197
    assert (db.dim, db.num_comp_indices, db.num_simp_indices) == (3072, 3, 15)
    assert db.num_points == 0 and db.num_levels == 0
    pv = db.proj_vec
    assert pv.shape == (45, 3072) and pv.dtype == np.float64            # py_dci.c:299-302: (m*L) x dim
    np.testing.assert_allclose(np.linalg.norm(pv[:, :], axis=1), 1.0, rtol=1e-12)   # dci.c:55-71 unit vectors
    pv[0, 0] = 0.5                                                      # writable while empty (dci.py:93-95)
    assert db.proj_vec[0, 0] == 0.5
    with pytest.raises(ValueError):
        db.proj_vec = np.zeros((3, 3))                                  # dci.py:102-103
    db.proj_vec = np.ones((45, 3072))
    assert db.proj_vec[7, 9] == 1.0
    for f in ("dim", "num_points", "num_levels", "num_comp_indices", "num_simp_indices"):
        with pytest.raises(AttributeError):
            setattr(db, f, 1)


def test_add_argument_checks_match_reference(native_lib):
    from inclusivegan_b200 import DCI
    db = DCI(8)
    with pytest.raises(ValueError, match="mismatch between array dimension"):
        db.add(np.zeros((4, 7)))                                        # dci.py:114-115
    with pytest.raises(TypeError, match="double-precision"):
        db.add(np.zeros((4, 8), dtype=np.int32))                        # dci.py:116-117
    with pytest.raises(TypeError, match="double-precision"):
        DCI(8, strict=True).add(np.zeros((4, 8), dtype=np.float32))     # reference is float64-only
    with pytest.raises(ValueError, match="row-major"):
        db.add(np.asfortranarray(np.zeros((4, 8))))                     # dci.py:118-119
    with pytest.raises(ValueError, match="derived from another array"):
        db.add(np.zeros((8, 8))[:4])                                    # dci.py:129-140
    with pytest.raises(ValueError, match="derived from another array"):
        db.add(np.zeros((8, 16))[:, :8].copy()[2:])
    with pytest.raises(TypeError, match="integer"):
        db.add(np.zeros((4, 8)), num_levels=3, field_of_view=10.0)      # dci.py:231-232
    with pytest.raises(ValueError, match="positive"):
        db.add(np.zeros((4, 8)), num_levels=3, field_of_view=0)
    with pytest.raises(IndexError):
        db.add(np.zeros((4, 8)), indices=9)
    with pytest.raises(IndexError):
        db.add(np.zeros((4, 8)), indices=[0, 4])
    with pytest.raises(TypeError):
        db.add(np.zeros((4, 8)), indices=np.array([0.5]))
    with pytest.raises(TypeError):
        db.add(np.zeros((4, 8)), indices="0")


def test_query_argument_checks_match_reference(native_lib):
    from inclusivegan_b200 import DCI
    db = DCI(8)
    with pytest.raises(ValueError, match="mismatch between array dimension"):
        db.query(np.zeros((2, 9)), num_neighbours=1)
    with pytest.raises(TypeError, match="integer"):
        db.query(np.zeros((2, 8)), num_neighbours=1.0)                  # dci.py:107-111,281
    with pytest.raises(ValueError, match="positive"):
        db.query(np.zeros((2, 8)), num_neighbours=0)
    with pytest.raises(ValueError, match="positive"):
        db.query(np.zeros((2, 8)))          # num_neighbours=-1 -> num_points == 0 -> "must be positive" (dci.py:278-281)


def test_select_rows_matches_reference_table(golden_dir, native_lib):
    """Row-selection semantics pinned against the reference's _check_and_fix_indices (dci.py:146-221)."""
    from inclusivegan_b200 import DCI
    z = np.load(os.path.join(golden_dir, "select_rows.npz"))
    data = np.zeros((10, 4))
    selectors = {
        "none": None, "slice_0_10": slice(0, 10), "slice_2_7": slice(2, 7), "slice_neg": slice(-4, -1),
        "slice_step2": slice(1, 9, 2), "slice_over": slice(5, 100), "int_3": 3, "int_neg1": -1,
        "arr_intc": np.array([5, 1, 8], dtype=np.intc), "arr_int64_neg": np.array([0, -1, 4], dtype=np.int64),
        "arr_bool": np.arange(10) % 3 == 0, "list_int": [9, 0, 2], "list_bool": [True, False] * 5,
    }
    for name, sel in selectors.items():
        contig, val = DCI._select_rows(data, sel)
        assert bool(contig) == bool(z[name + "__contig"]), name
        assert np.array_equal(np.asarray(val, dtype=np.int64), z[name + "__val"]), name
    # leniency: the reference crashes on slice(None) (stop is None, dci.py:161); we take it as "all rows"
    assert DCI._select_rows(data, slice(None)) == (True, (0, 10))


def test_protected_array_contract():
    from inclusivegan_b200 import ProtectedArray
    base = np.arange(6.0).reshape(2, 3)
    ro = ProtectedArray(base, when_writable=lambda _: False, write_error=lambda _: AttributeError("locked"))
    assert ro[1, 2] == 5.0 and ro.shape == (2, 3)
    with pytest.raises(AttributeError, match="locked"):
        ro[0, 0] = 1.0
    hidden = ProtectedArray(base, when_readable=lambda _: False)
    with pytest.raises(RuntimeError, match="not currently readable"):
        hidden[0]


def test_dci_code_shim_resolves_to_the_drop_in(native_lib):
    """`sys.path.append('./dci_code'); from dci import DCI` (training_loop.py:21-23) must find our class."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dci_shim_test", os.path.join(ROOT, "dci_code", "dci.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from inclusivegan_b200 import DCI
    assert mod.DCI is DCI


def test_shard_range():
    from inclusivegan_b200.sharding import shard_range, pad_local_topk
    for n in (0, 1, 7, 300000, 1000003):
        for g in (1, 2, 3, 4, 8):
            rs = [shard_range(n, g, r) for r in range(g)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(g - 1))
            assert all(0 <= b - a <= (n + g - 1) // g for a, b in rs)
    i, d = pad_local_topk(np.zeros((3, 2), np.int32), np.ones((3, 2)), 4)
    assert i.shape == (3, 4) and (i[:, 2:] == -1).all() and np.isinf(d[:, 2:]).all()


def test_reference_wrapper_runs_on_the_extension_stand_in(native_lib):
    """INTEGRATION.md option B: the reference's UNMODIFIED dci.py on top of inclusivegan_b200._dci (our stand-in for
    the compiled `_dci` extension).  Needs /root/reference, i.e. runs in the build container only; compute calls are
    exercised on the GPU by tests/test_gpu_parity.py::test_extension_stand_in_functions."""
    import importlib.util
    import sys
    path = "/root/reference/dci_code/src/dci.py"
    if not os.path.exists(path):
        pytest.skip("reference not present")
    if not hasattr(np, "float"):
        np.float = np.float64
    if not hasattr(np, "bool"):
        np.bool = np.bool_
    from inclusivegan_b200 import _dci as stand_in
    saved = sys.modules.get("_dci")
    sys.modules["_dci"] = stand_in
    try:
        spec = importlib.util.spec_from_file_location("_reference_dci_on_b200", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        db = mod.DCI(32, 3, 15)
        assert (db.dim, db.num_points, db.num_levels) == (32, 0, 0)
        assert db.proj_vec.shape == (45, 32)
        db.proj_vec = np.ones((45, 32))
        with pytest.raises(ValueError):
            db.add(np.zeros((4, 31)))                       # the reference's own checks still run first
        if native_lib.b200knn_device_count() == 0:
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                db.add(np.zeros((4, 32)))
        db.clear()
        db.reset()
    finally:
        if saved is None:
            sys.modules.pop("_dci", None)
        else:
            sys.modules["_dci"] = saved


def test_audit_switch_is_parsed_without_a_device(native_lib, monkeypatch):
    """DCI(audit=N) / $B200KNN_AUDIT (run-time cross-check of N rows per call against the exact scan): the switch is host state
    of the Python layer; the handle is created lazily, so this needs no GPU."""
    from inclusivegan_b200 import DCI
    monkeypatch.delenv("B200KNN_AUDIT", raising=False)
    assert DCI(8)._audit == 0
    monkeypatch.setenv("B200KNN_AUDIT", "32")
    assert DCI(8)._audit == 32
    assert DCI(8, audit=5)._audit == 5 and DCI(8, audit=0)._audit == 0
    db = DCI(8, audit=4)
    assert db.audited_queries == 0
    # nothing to audit: empty answers, scan answers, deliberately uncertified answers
    q = np.zeros((3, 8))
    db._audit_answers(q, 1, 4, np.zeros((3, 1), np.int32), np.zeros((3, 1)))       # FLAG_FORCE_SCAN
    db._audit_answers(q, 1, 2, np.zeros((3, 1), np.int32), np.zeros((3, 1)))       # FLAG_NO_CERTIFY
    db._audit_answers(q[:0], 1, 0, np.zeros((0, 1), np.int32), np.zeros((0, 1)))
    assert db.audited_queries == 0
