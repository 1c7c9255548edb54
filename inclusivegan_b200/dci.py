"""Drop-in `dci` module: the reference's `DCI` Python class over the B200 exact-kNN C ABI.

The reference trainer does `sys.path.append('./dci_code'); from dci import DCI`
(training/training_loop.py:21-23) and then uses `DCI(dim, num_comp_indices, num_simp_indices)`,
`reset()`, `add(...)`, `query(...)` (training_loop.py:197,367-368,383,398).  This module keeps that
surface — names, argument meaning, defaults, return types and the exception types/messages of the
reference's dci_code/src/dci.py — but every call lands in inclusivegan_b200/libb200knn.so
(include/b200knn.h) via ctypes over NumPy buffers.  No TensorFlow, Triton or torch on the path.

Differences from the reference, all deliberate:
  * results are EXACT (the reference is approximate); per query exactly min(k, num_points)
    neighbours come back.  The approximation knobs (num_levels, field_of_view, blind,
    num_to_visit, num_to_retrieve, prop_to_visit, prop_to_retrieve) are validated like the
    reference validates them and then ignored.  `blind=True` returns true distances, not
    projection-space priorities (reference: dci.c:446-457).
  * add() copies the rows to the GPU(s); the reference borrows the caller's buffer
    (py_dci.c:118-123).  A Python reference to `data` is still held until clear()/reset(), like
    dci.py:270.
  * float32 arrays are accepted as an extension where the reference raises TypeError; pass
    strict=True to the constructor to get the reference's float64-only behaviour.
  * there is NO CPU fallback: without a usable sm_100 GPU, add()/query() raise RuntimeError.
"""
import ctypes
import os

import numpy as np

__all__ = ["DCI", "DeviceKNN", "PeerExchange", "ProtectedArray", "B200KNNError", "load_library"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libb200knn.so"
_lib = None

F64, F32 = 0, 1
FLAG_SQUARED, FLAG_NO_CERTIFY, FLAG_FORCE_SCAN = 1, 2, 4
PRECISION_TIERS = {"bf16": 0, "bf16x3": 1, "tf32": 2}


class B200KNNError(RuntimeError):
    """A libb200knn call failed (carries the library's status code and message)."""

    def __init__(self, code, message):
        RuntimeError.__init__(self, "libb200knn error %d: %s" % (code, message))
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_int64), ("queries", ctypes.c_int64), ("uncertified", ctypes.c_int64),
                ("ms_convert", ctypes.c_double), ("ms_distance", ctypes.c_double), ("ms_rerank", ctypes.c_double),
                ("ms_scan", ctypes.c_double), ("distance_launches", ctypes.c_int64), ("distance_flops", ctypes.c_double),
                ("exact_scanned", ctypes.c_int64), ("ms_wait", ctypes.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


def load_library(path=None):
    """Load libb200knn.so (in-tree build).  Fails loudly if it is missing: there is no fallback path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("B200KNN_LIBRARY") or os.path.join(_HERE, _LIB_NAME)
    if not os.path.exists(path):
        raise RuntimeError("%s not found — build it with `python -m inclusivegan_b200.build` "
                           "(nvcc, sm_100a); there is no CPU or pure-Python fallback" % path)
    lib = ctypes.CDLL(path)
    vp, i32, i64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint
    lib.b200knn_create.restype = i32
    lib.b200knn_create.argtypes = [i32, i32, ctypes.POINTER(i32), ctypes.POINTER(vp)]
    lib.b200knn_destroy.restype = i32
    lib.b200knn_destroy.argtypes = [vp]
    lib.b200knn_clear.restype = i32
    lib.b200knn_clear.argtypes = [vp]
    lib.b200knn_num_points.restype = i64
    lib.b200knn_num_points.argtypes = [vp]
    lib.b200knn_dim.restype = i32
    lib.b200knn_dim.argtypes = [vp]
    lib.b200knn_add.restype = i32
    lib.b200knn_add.argtypes = [vp, vp, i32, i64, i64]
    lib.b200knn_query.restype = i32
    lib.b200knn_query.argtypes = [vp, vp, i32, i64, i64, i32, u32, vp, vp, ctypes.POINTER(i32)]
    lib.b200knn_set_stream.restype = i32
    lib.b200knn_set_stream.argtypes = [vp, vp]
    lib.b200knn_add_device.restype = i32
    lib.b200knn_add_device.argtypes = [vp, vp, i32, i64, i64, i64]
    lib.b200knn_query_device.restype = i32
    lib.b200knn_query_device.argtypes = [vp, vp, i32, i64, i64, i32, u32, vp, vp, ctypes.POINTER(i32)]
    lib.b200knn_merge_topk_device.restype = i32
    lib.b200knn_merge_topk_device.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp]
    lib.b200knn_set_profiling.restype = i32
    lib.b200knn_set_profiling.argtypes = [vp, i32]
    lib.b200knn_set_precision.restype = i32
    lib.b200knn_set_precision.argtypes = [vp, i32]
    lib.b200knn_get_stats.restype = i32
    lib.b200knn_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    lib.b200knn_reset_stats.restype = i32
    lib.b200knn_reset_stats.argtypes = [vp]
    lib.b200knn_query_self.restype = i32
    lib.b200knn_query_self.argtypes = [vp, i32, u32, vp, vp, ctypes.POINTER(i32)]
    lib.b200knn_exchange_create.restype = i32
    lib.b200knn_exchange_create.argtypes = [i32, i32, i32, i64, i32, ctypes.POINTER(vp)]
    lib.b200knn_exchange_handle.restype = i32
    lib.b200knn_exchange_handle.argtypes = [vp, vp]
    lib.b200knn_exchange_connect.restype = i32
    lib.b200knn_exchange_connect.argtypes = [vp, vp]
    lib.b200knn_exchange_allgather_merge.restype = i32
    lib.b200knn_exchange_allgather_merge.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp]
    lib.b200knn_exchange_destroy.restype = i32
    lib.b200knn_exchange_destroy.argtypes = [vp]
    lib.b200knn_exchange_create_for_queries.restype = i32
    lib.b200knn_exchange_create_for_queries.argtypes = [i32, i32, i32, i32, i64, i32, ctypes.POINTER(vp)]
    lib.b200knn_exchange_connect_local.restype = i32
    lib.b200knn_exchange_connect_local.argtypes = [vp, ctypes.POINTER(vp)]
    lib.b200knn_exchange_add.restype = i32
    lib.b200knn_exchange_add.argtypes = [vp, vp, vp, i32, i64, i64, i64]
    lib.b200knn_exchange_add_device.restype = i32
    lib.b200knn_exchange_add_device.argtypes = [vp, vp, vp, i32, i64, i64, i64]
    lib.b200knn_exchange_query.restype = i32
    lib.b200knn_exchange_query.argtypes = [vp, vp, vp, i32, i64, i64, i32, u32, vp, vp, ctypes.POINTER(i32)]
    lib.b200knn_exchange_query_device.restype = i32
    lib.b200knn_exchange_query_device.argtypes = [vp, vp, vp, i32, i64, i64, i32, u32, vp, vp, ctypes.POINTER(i32)]
    lib.b200knn_ball_membership.restype = i32
    lib.b200knn_ball_membership.argtypes = [vp, vp, i32, i64, i64, vp, vp]
    lib.b200knn_set_projector.restype = i32
    lib.b200knn_set_projector.argtypes = [vp, vp, i64, i64]
    lib.b200knn_add_projected.restype = i32
    lib.b200knn_add_projected.argtypes = [vp, vp, i32, i64, i64]
    lib.b200knn_query_projected.restype = i32
    lib.b200knn_query_projected.argtypes = [vp, vp, i32, i64, i64, i32, u32, vp, vp, ctypes.POINTER(i32)]
    lib.b200knn_project_rows.restype = i32
    lib.b200knn_project_rows.argtypes = [vp, vp, i32, i64, i64, vp]
    lib.b200knn_debug_shortlists.restype = i32
    lib.b200knn_debug_shortlists.argtypes = [vp, vp, vp, i64, ctypes.POINTER(i64), ctypes.POINTER(i32), ctypes.POINTER(i32)]
    lib.b200knn_last_error.restype = ctypes.c_char_p
    lib.b200knn_last_error.argtypes = []
    lib.b200knn_abi_version.restype = i32
    lib.b200knn_abi_version.argtypes = []
    lib.b200knn_device_count.restype = i32
    lib.b200knn_device_count.argtypes = []
    if lib.b200knn_abi_version() != 1:
        raise RuntimeError("libb200knn ABI version mismatch")
    _lib = lib
    return lib


def _check(code):
    if code != 0:
        raise B200KNNError(code, load_library().b200knn_last_error().decode("utf-8", "replace"))


class ProtectedArray(object):
    """Array wrapper whose element reads/writes can be gated (same contract as dci.py:30-59)."""

    def __init__(self, base_array, when_readable=None, read_error=None, when_writable=None, write_error=None):
        self._base = base_array
        self._when_readable = when_readable
        self._read_error = read_error
        self._when_writable = when_writable
        self._write_error = write_error

    def __getitem__(self, key):
        if self._when_readable is not None and not self._when_readable(key):
            raise (RuntimeError("array is not currently readable") if self._read_error is None else self._read_error(key))
        return self._base[key]

    def __setitem__(self, key, value):
        if self._when_writable is not None and not self._when_writable(key):
            raise (RuntimeError("array is not currently writable") if self._write_error is None else self._write_error(key))
        self._base[key] = value

    def __getattr__(self, name):
        return getattr(self._base, name)

    def __repr__(self):
        return repr(self._base)


def _require_positive_int(x):
    # dci.py:107-111
    if not isinstance(x, int):
        raise TypeError("number must be an integer")
    if x <= 0:
        raise ValueError("number must be positive")


def _draw_unit_directions(rows, dim):
    """Shape-/distribution-compatible stand-in for dci_gen_proj_vec (dci.c:55-71): Gaussian unit vectors."""
    v = np.random.standard_normal((rows, dim))
    v /= np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-300)
    return v


class DCI(object):
    """Exact k-nearest-neighbour index with the reference `DCI` interface (dci.py:61-340)."""

    def __init__(self, dim, num_comp_indices=2, num_simp_indices=7, devices=None, strict=False, precision=None, audit=None):
        """dim, num_comp_indices, num_simp_indices: as dci.py:63.  Extensions (keyword-only in spirit):
        devices   — GPU ids to row-shard the pool over (None: $B200KNN_DEVICES or the current device);
        strict    — refuse non-float64 data like the reference (dci.py:116-117);
        precision — tier of the tensor pass: 'bf16' (default), 'bf16x3' (split BF16, three MMAs) or 'tf32'; results are
                    exact in every tier, the tier only decides how much exact re-ranking the tensor scores leave;
        audit     — N > 0 (None: $B200KNN_AUDIT, default 0): after every query, N evenly spaced rows of it are answered
                    again by the exact float64 CUDA-core scan (no tensor scores, no certificate) and compared; a
                    difference raises RuntimeError.  A run-time cross-check of the certificate's one empirical premise
                    (the accumulation model of the tensor core)."""
        self._audit = int(os.environ.get("B200KNN_AUDIT", "0") or 0) if audit is None else int(audit)
        self.audited_queries = 0
        self._dim = int(dim)
        self._num_comp_indices = num_comp_indices
        self._num_simp_indices = num_simp_indices
        self._strict = bool(strict)
        self._lib = load_library()
        if devices is None and os.environ.get("B200KNN_DEVICES"):
            devices = [int(t) for t in os.environ["B200KNN_DEVICES"].split(",") if t.strip() != ""]
        if devices is None:
            n_dev, ids = 0, None
        else:
            devices = [int(d) for d in devices]
            n_dev, ids = len(devices), (ctypes.c_int * len(devices))(*devices)
        handle = ctypes.c_void_p()
        _check(self._lib.b200knn_create(self._dim, n_dev, ids, ctypes.byref(handle)))
        self._handle = handle
        if precision is not None:
            self.set_precision(precision)
        # Exact search uses no projections; the property is kept shape-correct (m*L x dim, float64) and
        # writable-when-empty because callers may read or pin it (dci.py:69,93-105; py_dci.c:299-302).
        self._proj_vec = _draw_unit_directions(num_comp_indices * num_simp_indices, self._dim)
        self._array = None
        self._orig_indices = None
        self._num_levels = 0

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.b200knn_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ---- properties (dci.py:73-105) -------------------------------------------------------------
    @property
    def dim(self):
        return self._dim

    @property
    def num_comp_indices(self):
        return self._num_comp_indices

    @property
    def num_simp_indices(self):
        return self._num_simp_indices

    @property
    def num_points(self):
        return int(self._lib.b200knn_num_points(self._handle))

    @property
    def num_levels(self):
        return self._num_levels

    def _empty_guard_error(self, _=None):
        return AttributeError("can only set projection vectors when the database is empty")

    @property
    def proj_vec(self):
        return ProtectedArray(self._proj_vec, when_writable=lambda _: self.num_points == 0, write_error=self._empty_guard_error)

    @proj_vec.setter
    def proj_vec(self, new_proj_vec):
        if self.num_points != 0:
            raise self._empty_guard_error()
        new_proj_vec = np.asarray(new_proj_vec)     # no broadcasting on direct assignment (dci.py:100-104)
        if new_proj_vec.shape != self._proj_vec.shape:
            raise ValueError("mismatch between the expected shape of projection vectors (%s) and the supplied shape (%s)"
                             % (repr(self._proj_vec.shape), repr(new_proj_vec.shape)))
        self._proj_vec[...] = new_proj_vec

    def stats(self):
        """Counters of the native library (kernel launches, uncertified queries, per-kernel ms when profiling)."""
        s = Stats()
        _check(self._lib.b200knn_get_stats(self._handle, ctypes.byref(s)))
        return s.as_dict()

    def set_precision(self, precision):
        """'bf16' | 'bf16x3' | 'tf32' (or the tier number): b200knn_set_precision."""
        tier = PRECISION_TIERS[precision] if isinstance(precision, str) else int(precision)
        _check(self._lib.b200knn_set_precision(self._handle, tier))

    def set_profiling(self, on):
        _check(self._lib.b200knn_set_profiling(self._handle, int(bool(on))))

    def debug_shortlists(self, capacity=1 << 24):
        """Test hook: (scores float32 [Q, slots, C], rows int32 [Q, slots, C]) of the last tensor pass."""
        sc = np.empty(capacity, dtype=np.float32)
        rw = np.empty(capacity, dtype=np.int32)
        nq, slots, c = ctypes.c_int64(0), ctypes.c_int(0), ctypes.c_int(0)
        _check(self._lib.b200knn_debug_shortlists(self._handle, sc.ctypes.data, rw.ctypes.data, capacity, ctypes.byref(nq),
                                                  ctypes.byref(slots), ctypes.byref(c)))
        n = nq.value * slots.value * c.value
        shape = (nq.value, slots.value, c.value)
        return sc[:n].reshape(shape), rw[:n].reshape(shape)

    # ---- argument checking (dci.py:107-221) -------------------------------------------------------
    def _dtype_code(self, arr):
        if arr.dtype == np.float64:
            return F64
        if arr.dtype == np.float32 and not self._strict:
            return F32
        raise TypeError("array must consist of double-precision floats")

    def _check_dim(self, arr):
        if arr.ndim != 2 or arr.shape[1] != self.dim:
            raise ValueError("mismatch between array dimension (%d) and the declared dimension of this DCI instance (%d)"
                             % (arr.shape[1] if arr.ndim == 2 else -1, self.dim))

    def _check_data(self, data):
        # dci.py:113-144: right width, float64, C-order, and not a view into another array
        if not isinstance(data, np.ndarray):
            raise TypeError("array must consist of double-precision floats")
        self._check_dim(data)
        self._dtype_code(data)
        if not data.flags.c_contiguous:
            raise ValueError("the memory layout of array must be in row-major (C-order)")
        if data.base is not None:
            root = data
            while isinstance(root.base, np.ndarray):
                root = root.base
            same_start = (isinstance(root, np.ndarray) and root.base is None and root.ctypes.data == data.ctypes.data
                          and root.size == data.size)
            if not same_start:
                raise ValueError("array must not be derived from another array, except via the transpose operator. "
                                 "Pass in the original array and specify the indices or make a copy of the derived array.")

    def _fix_query(self, query):
        # dci.py:121-127: queries are silently cast to C-contiguous float64
        query = np.asarray(query)
        if query.ndim != 2 or query.shape[1] != self.dim:
            raise ValueError("mismatch between array dimension (%d) and the declared dimension of this DCI instance (%d)"
                             % (query.shape[1] if query.ndim == 2 else -1, self.dim))
        if query.dtype == np.float64 or (query.dtype == np.float32 and not self._strict):
            return np.ascontiguousarray(query)
        return np.ascontiguousarray(query, dtype=np.float64)

    @staticmethod
    def _select_rows(data, indices):
        """dci.py:146-221 — returns (is_contiguous, (start, stop) | int32 index array)."""
        n = data.shape[0]
        need_bounds_check = False
        if indices is None:
            return True, (0, n)
        if isinstance(indices, slice):
            start = 0 if indices.start is None else indices.start
            stop = n if indices.stop is None else indices.stop
            step = 1 if indices.step is None else indices.step
            if start < 0:
                start += n
            if stop < 0:
                stop += n
            start, stop = max(start, 0), min(stop, n)
            if step == 1:
                return True, (start, stop)
            return False, np.arange(start, stop, step, dtype=np.intc)
        if isinstance(indices, (int, np.integer)) and not isinstance(indices, (bool, np.bool_)):
            i = int(indices) + (n if indices < 0 else 0)
            if i < 0 or i >= n:
                raise IndexError("index out of bounds")
            return True, (i, i + 1)
        if isinstance(indices, np.ndarray):
            if indices.ndim != 1:
                raise IndexError("indices must be in an one-dimensional array")
            if indices.dtype == np.bool_:
                if indices.shape[0] != n:
                    raise IndexError("mismatch between the number of boolean indices (%d) and array dimension (%d)"
                                     % (indices.shape[0], n))
                return False, np.nonzero(indices)[0].astype(np.intc)
            if indices.dtype.kind in "iu":
                sel = indices.astype(np.intc, copy=True)
                sel[sel < 0] += n
                need_bounds_check = True
            else:
                raise TypeError("indices must be integers or booleans")
        elif isinstance(indices, list):
            if len(indices) == 0:
                return False, np.zeros(0, dtype=np.intc)
            first = indices[0]
            if isinstance(first, (bool, np.bool_)):
                return False, np.nonzero(indices)[0].astype(np.intc)
            if isinstance(first, (int, np.integer)):
                sel = np.array(indices, dtype=np.intc)
                sel[sel < 0] += n
                need_bounds_check = True
            elif isinstance(first, list):
                raise IndexError("indices must be in an one-dimensional array")
            else:
                raise TypeError("indices must be integers or booleans")
        else:
            raise TypeError("indices must be None, a slice object, an integer, an array or list of integers")
        if need_bounds_check:
            bad = (sel < 0) | (sel >= n)
            if np.any(bad):
                raise IndexError("some indices (e.g. %d) out of bounds" % int(np.asarray(indices)[bad][0]))
        return False, sel

    # ---- add / query / clear / reset ------------------------------------------------------------------
    def add(self, data, indices=None, num_levels=2, field_of_view=10, blind=False, num_to_visit=-1, num_to_retrieve=-1,
            prop_to_visit=-1.0, prop_to_retrieve=-1.0):
        """Index `data` (N x dim, float64, C-order, a base array) — dci.py:224-270.

        One array per index (RuntimeError on a second add, dci.py:228-229).  `indices` selects rows
        (None / slice / int / int or bool ndarray / list); results of query() refer to row numbers of
        `data`.  The construction knobs are accepted, validated (field_of_view must be a positive int
        when num_levels >= 3, dci.py:231-234) and otherwise unused: the index is exact."""
        if self.num_points > 0:
            raise RuntimeError("DCI class does not support insertion of more than one array. "
                               "Must combine all arrays into one array before inserting")
        if num_levels >= 3:
            _require_positive_int(field_of_view)
        self._check_data(data)
        contiguous, sel = self._select_rows(data, indices)
        code = self._dtype_code(data)
        if contiguous:
            start, stop = sel
            rows = data[start:stop] if stop > start else data[0:0]
            self._orig_indices = None
            self._offset = start
        else:
            rows = np.ascontiguousarray(data[sel])       # gathered copy, results remapped below (dci.py:265-268)
            self._orig_indices = sel
            self._offset = 0
        if rows.shape[0] > 0:
            _check(self._lib.b200knn_add(self._handle, rows.ctypes.data, code, rows.shape[0], self._dim))
        self._array = data                               # keep the caller's array alive like dci.py:270
        self._num_levels = int(num_levels) if rows.shape[0] > 0 else 0

    def query(self, query, num_neighbours=-1, field_of_view=100, blind=False, num_to_visit=-1, num_to_retrieve=-1,
              prop_to_visit=-1.0, prop_to_retrieve=-1.0):
        """k nearest pool rows of every query row — dci.py:273-330.

        Returns (indices, distances): two lists of length Q; element i holds an int32 / float64
        ndarray of min(k, num_points) entries, ascending Euclidean distance (ties: lower index).
        num_neighbours < 0 means all points (dci.py:278-279)."""
        q = self._fix_query(query)
        num_points = self.num_points
        if num_neighbours < 0:
            num_neighbours = num_points
        _require_positive_int(num_neighbours)
        if self.num_levels >= 2:
            _require_positive_int(field_of_view)
        nq = q.shape[0]
        kk = min(num_neighbours, num_points)
        idx = np.empty((nq, kk), dtype=np.int32)
        dist = np.empty((nq, kk), dtype=np.float64)
        out_kk = ctypes.c_int(0)
        _check(self._lib.b200knn_query(self._handle, q.ctypes.data, self._dtype_code(q), nq, self._dim, int(num_neighbours),
                                       0, idx.ctypes.data, dist.ctypes.data, ctypes.byref(out_kk)))
        assert out_kk.value == kk
        self._audit_answers(q, int(num_neighbours), 0, idx, dist)
        if self._orig_indices is not None:
            idx = self._orig_indices[idx].astype(np.int32, copy=False)
        elif self._offset:
            idx += np.int32(self._offset)
        return list(idx), list(dist)          # per-query row views, like dci.py:318-330

    def query_arrays(self, query, num_neighbours, squared=False, flags=0):
        """Extension: same search, rectangular ndarray results (idx int32 [Q,kk], dist float64 [Q,kk])."""
        q = self._fix_query(query)
        _require_positive_int(num_neighbours)
        nq = q.shape[0]
        kk = min(num_neighbours, self.num_points)
        idx = np.empty((nq, kk), dtype=np.int32)
        dist = np.empty((nq, kk), dtype=np.float64)
        _check(self._lib.b200knn_query(self._handle, q.ctypes.data, self._dtype_code(q), nq, self._dim, int(num_neighbours),
                                       int(flags) | (FLAG_SQUARED if squared else 0), idx.ctypes.data, dist.ctypes.data, None))
        self._audit_answers(q, int(num_neighbours), int(flags) | (FLAG_SQUARED if squared else 0), idx, dist)
        if self._orig_indices is not None:
            idx = self._orig_indices[idx].astype(np.int32, copy=False)
        elif self._offset:
            idx += np.int32(self._offset)
        return idx, dist

    def _audit_answers(self, q, k, flags, idx, dist):
        """audit > 0: re-answer a sample of the call's rows by the exact scan (B200KNN_FLAG_FORCE_SCAN) and compare the
        library's raw output.  The scan sums a distance sequentially, the tensor paths in the canonical tree order:
        distances agree to float64 rounding, indices exactly unless two rows tie within that rounding."""
        if self._audit <= 0 or q.shape[0] == 0 or idx.shape[1] == 0 or (flags & (FLAG_FORCE_SCAN | FLAG_NO_CERTIFY)) or idx.shape[1] > 32:
            return                                   # (scan answers and deliberately uncertified ones: nothing to cross-check)
        rows = np.unique(np.linspace(0, q.shape[0] - 1, min(self._audit, q.shape[0])).astype(np.int64))
        sub = np.ascontiguousarray(q[rows])
        ai = np.empty((len(rows), idx.shape[1]), dtype=np.int32)
        ad = np.empty((len(rows), idx.shape[1]), dtype=np.float64)
        _check(self._lib.b200knn_query(self._handle, sub.ctypes.data, self._dtype_code(sub), len(rows), self._dim, k,
                                       flags | FLAG_FORCE_SCAN, ai.ctypes.data, ad.ctypes.data, None))
        self.audited_queries += len(rows)
        gi, gd = idx[rows], dist[rows]
        bad = np.abs(gd - ad) > 1e-9 * np.maximum(np.abs(ad), 1e-300)      # same distance, other row: a tie within rounding
        if bad.any():
            r, c = np.argwhere(bad)[0]
            raise RuntimeError("b200knn audit: query row %d rank %d: tensor path returned (%d, %r), exact scan (%d, %r)"
                               % (rows[r], c, gi[r, c], gd[r, c], ai[r, c], ad[r, c]))

    def query_self_arrays(self, num_neighbours, squared=False):
        """Extension: kNN of every indexed row among the indexed rows (itself first), without re-uploading them
        (b200knn_query_self; multi-device handles answer k <= 32 natively, larger k by a plain query of the original array)."""
        _require_positive_int(num_neighbours)
        n = self.num_points
        kk = min(num_neighbours, n)
        if self._orig_indices is not None or self._offset:
            raise ValueError("query_self_arrays needs an index built from a whole array (indices=None)")
        idx = np.empty((n, kk), dtype=np.int32)
        dist = np.empty((n, kk), dtype=np.float64)
        rc = self._lib.b200knn_query_self(self._handle, int(num_neighbours), FLAG_SQUARED if squared else 0, idx.ctypes.data,
                                          dist.ctypes.data, None)
        if rc == -1 and b"single-device" in self._lib.b200knn_last_error():
            return self.query_arrays(self._array, num_neighbours, squared=squared)
        _check(rc)
        return idx, dist

    def ball_membership(self, query, radius2):
        """Extension for the k-NN precision/recall metric: uint8 [Q], 1 where the query row lies inside at least one
        ball of squared radius radius2[j] around pool row j (metrics/precision_recall.py:119-120).  Exact."""
        if self._orig_indices is not None or self._offset:
            raise ValueError("ball_membership needs an index built from a whole array (indices=None)")
        q = self._fix_query(query)
        r2 = np.ascontiguousarray(radius2, dtype=np.float64)
        if r2.shape != (self.num_points,):
            raise ValueError("radius2 must have one entry per indexed point")
        out = np.zeros(q.shape[0], dtype=np.uint8)
        _check(self._lib.b200knn_ball_membership(self._handle, q.ctypes.data, self._dtype_code(q), q.shape[0], self._dim,
                                                 r2.ctypes.data, out.ctypes.data))
        return out

    # ---- random projection on the device (training_loop.py:205-212,362-365,379-381) ------------------
    def set_projector(self, projector):
        """Extension: register the trainer's random projector (in_dim x dim, float64).  add_projected() /
        query_projected() then take UNPROJECTED rows (in_dim wide, float32 or float64) and compute
        `rows.astype(float64) @ projector` on the device, in float64, before indexing / searching."""
        p = np.ascontiguousarray(projector, dtype=np.float64)
        if p.ndim != 2 or p.shape[1] != self.dim or p.shape[0] < 1:
            raise ValueError("projector must have shape (in_dim, %d), got %r" % (self.dim, p.shape))
        _check(self._lib.b200knn_set_projector(self._handle, p.ctypes.data, p.shape[0], p.shape[1]))
        self._proj_in_dim = int(p.shape[0])

    def _fix_unprojected(self, rows):
        in_dim = getattr(self, "_proj_in_dim", 0)
        if not in_dim:
            raise RuntimeError("no projector set: call set_projector() first")
        rows = np.asarray(rows)
        if rows.ndim > 2:                                   # images: flatten like np.reshape(x, (-1, prod(shape[1:])))
            rows = rows.reshape(rows.shape[0], -1)
        if rows.ndim != 2 or rows.shape[1] != in_dim:
            raise ValueError("mismatch between row width (%d) and the projector's input dimension (%d)"
                             % (rows.shape[1] if rows.ndim == 2 else -1, in_dim))
        if rows.dtype not in (np.float32, np.float64):
            rows = rows.astype(np.float64)
        return np.ascontiguousarray(rows)

    def project_rows(self, rows):
        """Extension (parity checks): the float64 projected rows the device computes."""
        r = self._fix_unprojected(rows)
        out = np.empty((r.shape[0], self.dim), dtype=np.float64)
        _check(self._lib.b200knn_project_rows(self._handle, r.ctypes.data, F64 if r.dtype == np.float64 else F32, r.shape[0],
                                              r.shape[1], out.ctypes.data))
        return out

    def add_projected(self, rows, num_levels=2, field_of_view=10, **_ignored):
        """Extension: add(rows.astype(float64) @ projector) without the host matmul (one array per index, like add)."""
        if self.num_points > 0:
            raise RuntimeError("DCI class does not support insertion of more than one array. "
                               "Must combine all arrays into one array before inserting")
        r = self._fix_unprojected(rows)
        if r.shape[0] > 0:
            _check(self._lib.b200knn_add_projected(self._handle, r.ctypes.data, F64 if r.dtype == np.float64 else F32,
                                                   r.shape[0], r.shape[1]))
        self._orig_indices = None
        self._offset = 0
        self._array = None
        self._num_levels = int(num_levels) if r.shape[0] > 0 else 0

    def query_projected_arrays(self, rows, num_neighbours, squared=False, flags=0):
        """Extension: query_arrays(rows.astype(float64) @ projector) without the host matmul."""
        r = self._fix_unprojected(rows)
        _require_positive_int(num_neighbours)
        kk = min(num_neighbours, self.num_points)
        idx = np.empty((r.shape[0], kk), dtype=np.int32)
        dist = np.empty((r.shape[0], kk), dtype=np.float64)
        _check(self._lib.b200knn_query_projected(self._handle, r.ctypes.data, F64 if r.dtype == np.float64 else F32, r.shape[0],
                                                 r.shape[1], int(num_neighbours), int(flags) | (FLAG_SQUARED if squared else 0),
                                                 idx.ctypes.data, dist.ctypes.data, None))
        return idx, dist

    def query_projected(self, rows, num_neighbours=-1, **_ignored):
        """Extension: query(rows.astype(float64) @ projector); same return convention as query()."""
        if num_neighbours < 0:
            num_neighbours = self.num_points
        idx, dist = self.query_projected_arrays(rows, num_neighbours)
        return list(idx), list(dist)

    def clear(self):
        """Drop the pool (dci.py:332-335)."""
        _check(self._lib.b200knn_clear(self._handle))
        self._array = None
        self._orig_indices = None
        self._offset = 0
        self._num_levels = 0

    def reset(self):
        """Drop the pool and redraw the (inert) projection directions (dci.py:337-340, dci.c:859-863)."""
        self.clear()
        self._proj_vec[...] = _draw_unit_directions(*self._proj_vec.shape)

    _offset = 0


class DeviceKNN(object):
    """Device-pointer face of the C ABI (b200knn_add_device / query_device / merge_topk_device).

    For hosts that already own device memory and streams (bench.py uses torch for that plumbing):
    every buffer is passed as a raw device address (int), nothing is copied, and all work is issued on
    the stream given to set_stream().  One handle = one GPU = one row shard of the pool; `index_base`
    is the shard's first global row (multi-process sharding: one rank per GPU)."""

    def __init__(self, dim, device=None):
        self._lib = load_library()
        self.dim = int(dim)
        handle = ctypes.c_void_p()
        if device is None:
            _check(self._lib.b200knn_create(self.dim, 0, None, ctypes.byref(handle)))
        else:
            ids = (ctypes.c_int * 1)(int(device))
            _check(self._lib.b200knn_create(self.dim, 1, ids, ctypes.byref(handle)))
        self._handle = handle

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.b200knn_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    @property
    def num_points(self):
        return int(self._lib.b200knn_num_points(self._handle))

    def set_stream(self, stream_ptr):
        _check(self._lib.b200knn_set_stream(self._handle, ctypes.c_void_p(stream_ptr or 0)))

    def set_profiling(self, on):
        _check(self._lib.b200knn_set_profiling(self._handle, int(bool(on))))

    def add(self, data_ptr, dtype, n, ld=None, index_base=0):
        _check(self._lib.b200knn_add_device(self._handle, ctypes.c_void_p(data_ptr), int(dtype), int(n),
                                            int(ld if ld is not None else self.dim), int(index_base)))

    def query(self, query_ptr, dtype, nq, k, out_idx_ptr, out_dist_ptr, ld=None, flags=0):
        kk = ctypes.c_int(0)
        _check(self._lib.b200knn_query_device(self._handle, ctypes.c_void_p(query_ptr), int(dtype), int(nq),
                                              int(ld if ld is not None else self.dim), int(k), int(flags),
                                              ctypes.c_void_p(out_idx_ptr), ctypes.c_void_p(out_dist_ptr), ctypes.byref(kk)))
        return kk.value

    def merge(self, idx_ptr, dist_ptr, n_lists, nq, kk, out_idx_ptr, out_dist_ptr, stream_ptr=0):
        _check(self._lib.b200knn_merge_topk_device(ctypes.c_void_p(idx_ptr), ctypes.c_void_p(dist_ptr), int(n_lists), int(nq),
                                                   int(kk), ctypes.c_void_p(out_idx_ptr), ctypes.c_void_p(out_dist_ptr),
                                                   ctypes.c_void_p(stream_ptr or 0)))

    def clear(self):
        _check(self._lib.b200knn_clear(self._handle))

    def stats(self):
        s = Stats()
        _check(self._lib.b200knn_get_stats(self._handle, ctypes.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        _check(self._lib.b200knn_reset_stats(self._handle))


class PeerExchange(object):
    """NVLink all-gather + merge of per-shard results over CUDA-IPC peer memory (b200knn_exchange_*), one process per
    GPU.  `handle()` bytes of every rank, in rank order, go to `connect()` (any out-of-band channel will do; bench.py
    uses torch.distributed.all_gather_object once at start-up)."""

    IPC_BYTES = 64

    def __init__(self, device, rank, world, max_nq, max_kk, dim=0):
        """dim > 0: also allocate the query-side buffers of the collective protocol (add / query / query_device below)."""
        self._lib = load_library()
        h = ctypes.c_void_p()
        if dim:
            _check(self._lib.b200knn_exchange_create_for_queries(int(device), int(rank), int(world), int(dim), int(max_nq), int(max_kk),
                                                                 ctypes.byref(h)))
        else:
            _check(self._lib.b200knn_exchange_create(int(device), int(rank), int(world), int(max_nq), int(max_kk), ctypes.byref(h)))
        self._h = h
        self.world = world

    # ---- collective protocol: every rank makes the same calls (see include/b200knn.h) ----
    def add_device(self, knn, data_ptr, dtype, n, ld=None, index_base=0):
        _check(self._lib.b200knn_exchange_add_device(self._h, knn._handle, ctypes.c_void_p(data_ptr), int(dtype), int(n),
                                                     int(ld if ld is not None else knn.dim), int(index_base)))

    def add(self, knn, rows, index_base=0):
        rows = np.ascontiguousarray(rows)
        _check(self._lib.b200knn_exchange_add(self._h, knn._handle, rows.ctypes.data, F64 if rows.dtype == np.float64 else F32,
                                              rows.shape[0], rows.shape[1], int(index_base)))

    def query_device(self, knn, query_ptr, dtype, nq, k, out_idx_ptr, out_dist_ptr, ld=None, flags=0):
        kk = ctypes.c_int(0)
        _check(self._lib.b200knn_exchange_query_device(self._h, knn._handle, ctypes.c_void_p(query_ptr), int(dtype), int(nq),
                                                       int(ld if ld is not None else knn.dim), int(k), int(flags),
                                                       ctypes.c_void_p(out_idx_ptr), ctypes.c_void_p(out_dist_ptr), ctypes.byref(kk)))
        return kk.value

    def query_host(self, knn, query_ptr, dtype, nq, k, out_idx_ptr, out_dist_ptr, ld=None, flags=0):
        kk = ctypes.c_int(0)
        _check(self._lib.b200knn_exchange_query(self._h, knn._handle, ctypes.c_void_p(query_ptr), int(dtype), int(nq),
                                                int(ld if ld is not None else knn.dim), int(k), int(flags),
                                                ctypes.c_void_p(out_idx_ptr), ctypes.c_void_p(out_dist_ptr), ctypes.byref(kk)))
        return kk.value

    def query(self, knn, queries, k, flags=0):
        """NumPy face of query_host: (idx int32 [Q, kk], dist float64 [Q, kk]) on every rank."""
        q = np.ascontiguousarray(queries)
        if q.dtype not in (np.float32, np.float64):
            q = q.astype(np.float64)
        idx = np.empty((q.shape[0], k), dtype=np.int32)
        dist = np.empty((q.shape[0], k), dtype=np.float64)
        kk = self.query_host(knn, q.ctypes.data, F64 if q.dtype == np.float64 else F32, q.shape[0], k, idx.ctypes.data, dist.ctypes.data,
                             flags=flags)
        if kk != k:          # fewer rows in the whole pool than k: the library wrote [Q][kk] densely
            idx = idx.reshape(-1)[:q.shape[0] * kk].reshape(q.shape[0], kk)
            dist = dist.reshape(-1)[:q.shape[0] * kk].reshape(q.shape[0], kk)
        return idx, dist

    def handle(self):
        buf = ctypes.create_string_buffer(self.IPC_BYTES)
        _check(self._lib.b200knn_exchange_handle(self._h, buf))
        return bytes(buf.raw)

    def connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == self.world * self.IPC_BYTES
        _check(self._lib.b200knn_exchange_connect(self._h, ctypes.c_char_p(blob)))

    def allgather_merge(self, idx_ptr, dist_ptr, nq, kk, out_idx_ptr, out_dist_ptr, stream_ptr=0):
        _check(self._lib.b200knn_exchange_allgather_merge(self._h, ctypes.c_void_p(idx_ptr), ctypes.c_void_p(dist_ptr), int(nq), int(kk),
                                                          ctypes.c_void_p(out_idx_ptr), ctypes.c_void_p(out_dist_ptr),
                                                          ctypes.c_void_p(stream_ptr or 0)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200knn_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
