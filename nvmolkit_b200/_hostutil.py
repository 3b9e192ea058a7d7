"""Vectorised host-side index bookkeeping (no per-conformer Python loops: batches reach 1e6 conformers)."""

from __future__ import annotations

import numpy as np


def rows_of(starts: np.ndarray, order: np.ndarray) -> np.ndarray:
    """Concatenation of arange(starts[c], starts[c+1]) for c in `order` (CSR row gather), int64."""
    from nvmolkit_b200 import _lib

    mod = _lib.core()
    if mod is not None:
        return mod.rows_of(np.ascontiguousarray(starts, dtype=np.int64), np.ascontiguousarray(order, dtype=np.int64))
    starts = np.asarray(starts, dtype=np.int64)
    order = np.asarray(order, dtype=np.int64)
    sizes = starts[order + 1] - starts[order]
    total = int(sizes.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    out_start = np.cumsum(sizes) - sizes
    return np.repeat(starts[order] - out_start, sizes) + np.arange(total, dtype=np.int64)


def running_index(keys: np.ndarray) -> np.ndarray:
    """k-th occurrence number of each key, in order of appearance (conformer index within its molecule), int32."""
    from nvmolkit_b200 import _lib

    mod = _lib.core()
    if mod is not None:
        return mod.running_index(np.ascontiguousarray(keys, dtype=np.int64))
    keys = np.asarray(keys)
    n = len(keys)
    if n == 0:
        return np.zeros(0, dtype=np.int32)
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    first = np.concatenate([[True], sk[1:] != sk[:-1]])
    group_start = np.maximum.accumulate(np.where(first, np.arange(n), 0))
    out = np.empty(n, dtype=np.int32)
    out[order] = (np.arange(n) - group_start).astype(np.int32)
    return out
