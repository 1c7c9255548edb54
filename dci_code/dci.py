"""Drop-in for the reference's dci_code/dci.py (a symlink to dci_code/src/dci.py there).

The reference trainer does `sys.path.append('./dci_code'); from dci import DCI`
(training/training_loop.py:21-23).  Copy or symlink this directory over the reference's
`dci_code/` (or put it first on sys.path) and the same import resolves to the B200 engine.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inclusivegan_b200.dci import DCI, ProtectedArray  # noqa: E402,F401
