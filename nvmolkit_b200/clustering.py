"""Butina clustering on the GPU. API of ``nvmolkit/clustering.py`` (reference :41-189).

Both entry points implement RDKit's ``ClusterData(reordering=True)`` definition exactly (ties -> highest index),
so the cluster assignment is deterministic and equal to the CPU result.
"""

from __future__ import annotations

import torch

from nvmolkit_b200 import _lib
from nvmolkit_b200._interop import as_tensor, require_cuda, stream_ctx, stream_ptr
from nvmolkit_b200.types import AsyncGpuResult

_VALID_NEIGHBORLIST_SIZES = frozenset({8, 16, 24, 32, 64, 128})


def butina(distance_matrix, cutoff: float, neighborlist_max_size: int = 64, return_centroids: bool = False,
           stream=None):
    """Cluster from a dense (N, N) fp64 distance matrix. Returns cluster ids (N,), cluster 0 the largest.

    ``neighborlist_max_size`` is validated for API compatibility; the CSR design has no such limit.
    """
    if neighborlist_max_size not in _VALID_NEIGHBORLIST_SIZES:
        raise ValueError(
            f"neighborlist_max_size must be one of {sorted(_VALID_NEIGHBORLIST_SIZES)}, got {neighborlist_max_size}")
    sptr = stream_ptr(stream)
    require_cuda()
    d = as_tensor(distance_matrix)
    if d.ndim != 2 or d.shape[0] != d.shape[1]:
        raise ValueError(f"distance_matrix must be square, got shape={tuple(d.shape)}")
    if d.dtype != torch.float64:
        raise ValueError(f"distance_matrix must be float64, got {d.dtype}")
    d = d.contiguous()
    n = d.shape[0]
    with stream_ctx(stream):
        ids = torch.empty(n, dtype=torch.int32, device=d.device)
        cen = torch.empty(max(n, 1), dtype=torch.int32, device=d.device)
        ncl = torch.zeros(1, dtype=torch.int32, device=d.device)
        _lib.call("b200mol_butina_dense", d.data_ptr(), n, float(cutoff), ids.data_ptr(), cen.data_ptr(),
                  ncl.data_ptr(), None, sptr)
        d.record_stream(torch.cuda.current_stream())
        if return_centroids:
            k = int(ncl.item())
            return AsyncGpuResult(ids), AsyncGpuResult(cen[:k])
    return AsyncGpuResult(ids)


def fused_butina_device(x, cutoff: float, stream=None, metric: str = "tanimoto"):
    """Device-resident result of the fused path: (ids int32[N], centroids int32[nClusters])."""
    if not isinstance(x, torch.Tensor):
        raise TypeError("x must be a torch.Tensor")
    if not x.is_cuda:
        raise ValueError("x must be a CUDA tensor")
    if x.dtype != torch.int32:
        raise ValueError("x must have dtype int32")
    if x.ndim != 2:
        raise ValueError(f"x must be 2D, got shape={tuple(x.shape)}")
    if metric not in ["tanimoto", "cosine"]:
        raise ValueError(f"metric must be one of ['tanimoto', 'cosine'], got {metric}")
    sptr = stream_ptr(stream)
    if cutoff < 0 or cutoff > 1:
        raise ValueError(f"cutoff must be in [0, 1], got {cutoff}")
    require_cuda()
    x = x.contiguous()
    n = x.shape[0]
    with stream_ctx(stream):
        ids = torch.empty(n, dtype=torch.int32, device=x.device)
        cen = torch.empty(max(n, 1), dtype=torch.int32, device=x.device)
        ncl = torch.zeros(1, dtype=torch.int32, device=x.device)
        _lib.call("b200mol_butina_fused", x.data_ptr(), n, x.shape[1], _lib.METRIC[metric], float(cutoff),
                  ids.data_ptr(), cen.data_ptr(), ncl.data_ptr(), None, sptr)
        x.record_stream(torch.cuda.current_stream())
        k = int(ncl.item())
    return ids, cen[:k]


def fused_butina_sharded(x, cutoff: float, stream=None, metric: str = "tanimoto", group=None):
    """Multi-GPU fused Butina: every rank holds all fingerprints, evaluates the pair tiles of its tile-row groups
    (rank, rank + world, ...), then the ranks all-reduce the neighbour counts and all-gather their edge lists (NCCL)
    and each clusters the full graph. Returns the same (ids, centroids) on every rank."""
    import ctypes as C

    import torch.distributed as dist

    from nvmolkit_b200.distributed import all_gather_v, rank_world

    rank, world = rank_world(group)
    if world == 1:
        return fused_butina_device(x, cutoff, stream=stream, metric=metric)
    sptr = stream_ptr(stream)
    x = x.contiguous()
    n, words = x.shape
    with stream_ctx(stream):
        cap = max(4096, (n * 64) // world)
        while True:
            counts = torch.zeros(n, dtype=torch.int32, device=x.device)
            edges = torch.empty((cap, 2), dtype=torch.int32, device=x.device)
            found = C.c_uint64(0)
            _lib.call("b200mol_neighbor_edges", x.data_ptr(), n, words, _lib.METRIC[metric], float(cutoff), rank, world,
                      counts.data_ptr(), edges.data_ptr(), cap, C.byref(found), sptr)
            if found.value <= cap:
                break
            cap = int(found.value)
        dist.all_reduce(counts, group=group)
        all_edges, _ = all_gather_v(edges[: found.value], group=group)
        all_edges = all_edges.contiguous()
        ids = torch.empty(n, dtype=torch.int32, device=x.device)
        cen = torch.empty(max(n, 1), dtype=torch.int32, device=x.device)
        ncl = torch.zeros(1, dtype=torch.int32, device=x.device)
        _lib.call("b200mol_butina_from_edges", n, counts.data_ptr(), all_edges.data_ptr(), all_edges.shape[0],
                  ids.data_ptr(), cen.data_ptr(), ncl.data_ptr(), None, sptr)
        k = int(ncl.item())
    return ids, cen[:k]


def fused_butina(x, cutoff: float, return_centroids: bool = False, stream=None, metric: str = "tanimoto"):
    """Butina straight from fingerprints, O(N + edges) memory (the similarity matrix is never materialised).

    Returns ``(clusters, cluster_sizes[, centroids])`` like the reference: clusters is a list of tuples with the
    centroid first, cluster_sizes the cumulative sizes starting at 0.

    Neighbour predicate: fp64 ``1.0 - c/u <= cutoff`` evaluated exactly (an integer threshold table), the SAME predicate as
    :func:`butina` on the fp64 distance matrix and as RDKit's ``ClusterData`` on ``1 - BulkTanimotoSimilarity``. The
    reference's fused path tests ``sim >= float32(1 - cutoff)`` in fp32 (``nvmolkit/_fusedButina.py:172-173``): pairs that
    sit exactly on the boundary as a ratio of small integers (cutoff 0.3, sim = 7/10: 1 - 0.7 = 0.30000000000000004 > 0.3)
    are neighbours there and not here. Deliberate: this path agrees with the dense path and with RDKit bit for bit.
    """
    ids, cen = fused_butina_device(x, cutoff, stream=stream, metric=metric)
    clusters, sizes = clusters_from_ids(ids.cpu().numpy(), cen.cpu().numpy())
    if return_centroids:
        return clusters, sizes, cen.cpu().numpy().tolist()
    return clusters, sizes


def clusters_from_ids(ids_h, cen_h):
    """(cluster ids [N], centroids [K]) -> the reference's output format (nvmolkit/clustering.py:171-189): list of
    tuples, centroid first then the other members in ascending index order, and the cumulative sizes starting at 0.
    NumPy only: one stable argsort + one split, no per-point Python loop (1M points: tens of milliseconds)."""
    import numpy as np

    ids_h = np.asarray(ids_h)
    cen_h = np.asarray(cen_h)
    k = len(cen_h)
    counts = np.bincount(ids_h, minlength=k)
    sizes = np.concatenate([[0], np.cumsum(counts)])
    # sort key: (cluster id, not-the-centroid, index) -> the centroid leads its cluster, members follow ascending
    n = len(ids_h)
    idx = np.arange(n, dtype=np.int64)
    is_member = (cen_h[ids_h] != idx).astype(np.int64) if k else np.zeros(0, np.int64)
    order = np.argsort((ids_h.astype(np.int64) * 2 + is_member) * max(n, 1) + idx)
    clusters = [tuple(part.tolist()) for part in np.split(order, sizes[1:-1])] if k else []
    return clusters, sizes.tolist()
