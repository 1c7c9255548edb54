"""Stand-in for the reference's compiled extension module `_dci` (dci_code/src/py_dci.c:311-321) over the C ABI.

For maintainers who keep the reference's own Python wrapper (dci_code/src/dci.py) and only swap the native layer:
put this module on the path as `_dci` (e.g. `sys.modules['_dci'] = inclusivegan_b200._dci`) and the unmodified
`dci.py` runs on the B200 engine.  Function names, argument lists and return shapes are exactly what dci.py calls
(dci.py:68-69,82,86,263-267,313,333,338).  INTEGRATION.md, option B.
"""
import ctypes

import numpy as np

from .dci import load_library, _check, F64


class _Inst(object):
    """Plays the role of the PyCapsule "py_dci_inst" (py_dci.c:46-64)."""

    def __init__(self, dim, num_comp_indices, num_simp_indices):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        self.dim = dim
        self.offset = 0
        self.levels = 0
        self.data = None                                   # like Py_INCREF(py_data), py_dci.c:123
        v = np.random.standard_normal((num_comp_indices * num_simp_indices, dim))
        self.proj = v / np.linalg.norm(v, axis=1, keepdims=True)     # inert (dci.c:55-71 draws unit vectors)
        _check(self.lib.b200knn_create(dim, 0, None, ctypes.byref(self.h)))

    def __del__(self):
        try:
            self.lib.b200knn_destroy(self.h)
        except Exception:
            pass


def new(dim, num_comp_indices, num_simp_indices):                                   # py_dci.c:66-83
    return _Inst(dim, num_comp_indices, num_simp_indices)


def add(inst, data, start, end, num_levels, blind, num_to_visit, num_to_retrieve, prop_to_visit, prop_to_retrieve,
        field_of_view):                                                             # py_dci.c:86-128
    rows = data[start:end]                         # float64, C-contiguous, checked by dci.py:113-119
    if rows.shape[0] > 0:
        _check(inst.lib.b200knn_add(inst.h, rows.ctypes.data, F64, rows.shape[0], data.shape[1]))
        inst.offset, inst.levels, inst.data = start, num_levels, data


def query(inst, q, num_neighbours, blind, num_to_visit, num_to_retrieve, prop_to_visit, prop_to_retrieve,
          field_of_view):                                                           # py_dci.c:130-211
    n = int(inst.lib.b200knn_num_points(inst.h))
    nq = q.shape[0]
    kk = min(num_neighbours, n)
    idx = np.empty((nq, kk), np.int32)
    dist = np.empty((nq, kk), np.float64)
    _check(inst.lib.b200knn_query(inst.h, q.ctypes.data, F64, nq, q.shape[1], num_neighbours, 0, idx.ctypes.data,
                                  dist.ctypes.data, None))
    idx += np.int32(inst.offset)                                                   # py_dci.c:185 data_idx_offset
    return idx.ravel(), dist.ravel(), np.full(nq, kk, np.int32)                    # flat idx, flat dist, counts


def clear(inst):                                                                    # py_dci.c:214-236
    _check(inst.lib.b200knn_clear(inst.h))
    inst.levels, inst.data = 0, None


def reset(inst):                                                                    # py_dci.c:238-259
    clear(inst)


def get_num_points(inst):                                                           # py_dci.c:262-272
    return int(inst.lib.b200knn_num_points(inst.h))


def get_num_levels(inst):                                                           # py_dci.c:275-285
    return inst.levels


def get_proj_vec(inst):                                                             # py_dci.c:288-305
    return inst.proj
