"""N x M fingerprint similarity on the GPU. API of ``nvmolkit/similarity.py`` (reference :34-184)."""

from __future__ import annotations

import numpy as np
import torch

from nvmolkit_b200 import _lib
from nvmolkit_b200._interop import fingerprint_matrix, require_cuda, stream_ctx, stream_ptr
from nvmolkit_b200.types import AsyncGpuResult


def _cross(fn_name: str, one, two, stream) -> AsyncGpuResult:
    sptr = stream_ptr(stream)  # validates `stream` first, like the reference (TypeError)
    require_cuda()
    a = fingerprint_matrix(one, "fingerprint_group_one")
    b = a if two is None else fingerprint_matrix(two, "fingerprint_group_two")
    if a.shape[1] != b.shape[1]:
        raise ValueError(f"fingerprint widths differ: {a.shape[1]} vs {b.shape[1]} words")
    with stream_ctx(stream):
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device=a.device)
        _lib.call(fn_name, a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], a.shape[1], out.data_ptr(), sptr)
        # keep inputs alive until the stream has consumed them
        a.record_stream(torch.cuda.current_stream())
        b.record_stream(torch.cuda.current_stream())
    return AsyncGpuResult(out)


def crossTanimotoSimilarity(fingerprint_group_one, fingerprint_group_two=None, stream=None) -> AsyncGpuResult:
    """(n, m) fp64 Tanimoto similarities; all-to-all within group one when group two is None. Asynchronous."""
    return _cross("b200mol_tanimoto_cross", fingerprint_group_one, fingerprint_group_two, stream)


def crossCosineSimilarity(fingerprint_group_one, fingerprint_group_two=None, stream=None) -> AsyncGpuResult:
    return _cross("b200mol_cosine_cross", fingerprint_group_one, fingerprint_group_two, stream)


def _cross_host(metric: str, one, two, max_device_bytes: int = 0) -> np.ndarray:
    require_cuda()

    def host(x):
        if isinstance(x, AsyncGpuResult):
            x = x.torch()
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x)
        if x.dtype not in (np.int32, np.uint32) or x.ndim != 2:
            raise ValueError("fingerprints must be a 2D int32 array of packed 32-bit words")
        return x

    a = host(one)
    b = a if two is None else host(two)
    if a.shape[1] != b.shape[1]:
        raise ValueError(f"fingerprint widths differ: {a.shape[1]} vs {b.shape[1]} words")
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    _lib.call("b200mol_similarity_cross_host", a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], a.shape[1],
              _lib.METRIC[metric], out.ctypes.data, int(max_device_bytes))
    return out


def crossTanimotoSimilarityMemoryConstrained(fingerprint_group_one, fingerprint_group_two=None) -> np.ndarray:
    """Host-in / host-out variant: result as a NumPy array, computed in row blocks with overlapped D2H."""
    return _cross_host("tanimoto", fingerprint_group_one, fingerprint_group_two)


def crossCosineSimilarityMemoryConstrained(fingerprint_group_one, fingerprint_group_two=None) -> np.ndarray:
    return _cross_host("cosine", fingerprint_group_one, fingerprint_group_two)
