"""TEST INFRASTRUCTURE — not product code.

Driver for the UNMODIFIED reference DCI, compiled by oracle/Makefile into oracle/_ref/_dci.so
(CPython extension `_dci`, reference dci_code/src/py_dci.c:311-321).  Used

  * by tests/golden/make_golden.py to generate fixtures (reference in exhaustive mode is exact),
  * by bench.py's `cpu_baseline` leg and `--impl reference` arm as the CPU baseline.

The GPU box has no /root/reference, so this module talks to the extension's functions directly
(`_dci.new / add / query / clear`), with the very argument lists the reference's own Python wrapper
passes (dci_code/src/dci.py:68,263,313); the wrapper's defaulting rules that matter for timing are
restated in `RefDCI.add/query` with file:line citations.  When /root/reference IS present (this
container), `import_reference_wrapper()` imports the real dci.py for cross-checks.
"""
import importlib.util
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "_dci.so")
_ext = None


def available():
    return os.path.exists(_REF_SO)


def ext():
    """The compiled reference extension module `_dci`."""
    global _ext
    if _ext is None:
        if not available():
            raise RuntimeError("oracle/_ref/_dci.so missing — run `make -C oracle ref` where /root/reference exists")
        os.environ.setdefault("OMP_STACKSIZE", "256M")   # dci.c:572-573 keeps ~1.3 MB VLAs on worker stacks
        spec = importlib.util.spec_from_file_location("_dci", _REF_SO)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ext = mod
    return _ext


def import_reference_wrapper(ref_root="/root/reference"):
    """Import the reference's own dci.py (only possible where /root/reference exists)."""
    path = os.path.join(ref_root, "dci_code", "src", "dci.py")
    if not os.path.exists(path):
        raise RuntimeError("reference not present at %s" % ref_root)
    # dci.py:116,124,127,189 use aliases NumPy >= 1.24 removed; set them in the harness, not the reference
    if not hasattr(np, "float"):
        np.float = np.float64
    if not hasattr(np, "bool"):
        np.bool = np.bool_
    sys.modules["_dci"] = ext()
    spec = importlib.util.spec_from_file_location("_reference_dci_wrapper", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class RefDCI(object):
    """Thin driver over the reference extension: same calls as dci_code/src/dci.py makes."""

    def __init__(self, dim, num_comp_indices=2, num_simp_indices=7):
        self._e = ext()
        self.dim = dim
        self._h = self._e.new(dim, num_comp_indices, num_simp_indices)           # dci.py:68
        self.proj_vec = self._e.get_proj_vec(self._h)                            # dci.py:69
        self._data = None

    @property
    def num_points(self):
        return self._e.get_num_points(self._h)

    def add(self, data, num_levels=2, field_of_view=10, prop_to_retrieve=0.002, prop_to_visit=1.0, blind=False):
        assert data.dtype == np.float64 and data.flags.c_contiguous and data.shape[1] == self.dim
        if num_levels < 3:
            field_of_view = -1                                                   # dci.py:231-234
        self._data = data                                                        # index borrows the buffer
        # dci.py:263: add(inst, data, start, end, num_levels, blind, n_visit, n_retrieve, p_visit, p_retrieve, fov)
        self._e.add(self._h, data, 0, data.shape[0], num_levels, blind, -1, -1, float(prop_to_visit),
                    float(prop_to_retrieve), field_of_view)

    def query(self, query, num_neighbours, field_of_view=100, prop_to_retrieve=0.05, prop_to_visit=1.0, blind=False):
        q = np.ascontiguousarray(query, dtype=np.float64)                         # dci.py:121-127,274
        if self._e.get_num_levels(self._h) < 2:
            field_of_view = -1                                                   # dci.py:283-286
        # dci.py:313: query(inst, q, k, blind, n_visit, n_retrieve, p_visit, p_retrieve, fov)
        flat_idx, flat_dist, counts = self._e.query(self._h, q, int(num_neighbours), blind, -1, -1,
                                                    float(prop_to_visit), float(prop_to_retrieve), field_of_view)
        return flat_idx, flat_dist, counts

    def clear(self):
        self._e.clear(self._h)
        self._data = None


def split(flat_idx, flat_dist, counts):
    """Flat ragged results -> per-query lists (what dci.py:318-330 does)."""
    off = np.concatenate([[0], np.cumsum(counts)])
    return ([flat_idx[off[i]:off[i + 1]] for i in range(len(counts))],
            [flat_dist[off[i]:off[i + 1]] for i in range(len(counts))])
