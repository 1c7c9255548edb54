"""inclusivegan_b200 — B200-native exact kNN for InclusiveGAN's IMLE matching step.

Only what the hot path needs: csrc/ (CUDA kernels + C ABI, built into libb200knn.so by build.py)
and dci.py (the host-side mirror of the reference's `DCI` Python API over that C ABI).
"""
from .dci import DCI, DeviceKNN, ProtectedArray, B200KNNError, load_library  # noqa: F401

__all__ = ["DCI", "DeviceKNN", "ProtectedArray", "B200KNNError", "load_library"]
