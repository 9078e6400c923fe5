"""Host bookkeeping kept from the reference's contract (core/resources.py):
``assignable`` (does an array of this shape fit in RAM?) and ``pickleable``
(the probe tests/test_concurrency.py uses on producers)."""

import pickle

import numpy as np

try:
    import psutil
except ImportError:      # pragma: no cover
    psutil = None


def assignable(shape, dtype=float, limit=None, tol=50e6):
    required = float(np.prod(shape, dtype=np.float64)) * np.dtype(dtype).itemsize
    if limit is None:
        limit = psutil.virtual_memory().available if psutil else float("inf")
    if limit - required > tol:
        return True
    print("openseize_b200: array of {:.1f} MB exceeds the {:.1f} MB available"
          .format(required / 1e6, limit / 1e6))
    return False


def pickleable(obj):
    try:
        pickle.dumps(obj)
        return True
    except Exception:
        return False
