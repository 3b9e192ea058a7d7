"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on the box, gloo in CPU tests).

The hot path shards by independent units (tile-row groups of the pair matrix, molecule ranges of a conformer batch);
the only collectives are the result exchanges at the end of a step.
"""

from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def rank_world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def molecule_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Plain contiguous molecule-range split [lo, hi) (north_star: 'plain molecule-range split')."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_v(t: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather of tensors whose first dimension differs per rank. Returns (concatenated, sizes[world]).

    One all-gather of the sizes, one all-gather of max-padded payloads (NCCL has no native all-gather-v)."""
    rank, world = rank_world(group)
    if world == 1:
        return t, torch.tensor([t.shape[0]], dtype=torch.int64)
    size = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = torch.empty(world, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(sizes, size, group=group)
    sizes_h = sizes.cpu()
    mx = int(sizes_h.max().item())
    padded = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    padded[: t.shape[0]] = t
    gathered = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = [gathered[r * mx: r * mx + int(sizes_h[r])] for r in range(world)]
    return torch.cat(parts, dim=0), sizes_h


def sharded_upload(h: torch.Tensor, device, group=None) -> torch.Tensor:
    """Host rows that EVERY rank needs on its device (the fingerprints of a sharded Butina pass): each rank copies only
    its 1/world of the rows over PCIe and the ranks all-gather the slices over NVLink, instead of every rank pulling the
    whole array through the host at the same time (8 x 256 MB at 1M fingerprints). `h` is the same host tensor on every
    rank (pinned for an asynchronous copy). Returns the full [n, ...] device tensor; the copies are stream-ordered."""
    rank, world = rank_world(group)
    if world == 1:
        return h.to(device, non_blocking=True)
    n = h.shape[0]
    chunk = (n + world - 1) // world  # equal slices for all_gather_into_tensor; the last one is padded
    full = torch.empty((world * chunk,) + tuple(h.shape[1:]), dtype=h.dtype, device=device)
    lo, hi = min(n, rank * chunk), min(n, (rank + 1) * chunk)
    mine = full[rank * chunk:(rank + 1) * chunk]
    mine[: hi - lo].copy_(h[lo:hi], non_blocking=True)
    if hi - lo < chunk:
        mine[hi - lo:].zero_()
    dist.all_gather_into_tensor(full, mine.clone() if full.device.type == "cpu" else mine, group=group)
    return full[:n]


def map_molecule_range(n_items: int, compute, group=None) -> torch.Tensor:
    """Run `compute(lo, hi) -> tensor[hi - lo, ...]` on this rank's contiguous share of `n_items` independent items
    (molecules, fingerprint rows) and return the results of all ranks, concatenated in item order, on every rank.
    The one collective is the all-gather-v of the results (Morgan fingerprints, conformer results)."""
    rank, world = rank_world(group)
    lo, hi = molecule_range(n_items, rank, world)
    part = compute(lo, hi)
    if part.shape[0] != hi - lo:
        raise ValueError(f"compute({lo}, {hi}) returned {part.shape[0]} rows")
    out, _ = all_gather_v(part, group)
    return out


def sharded_fingerprints(generator, graphs, group=None) -> torch.Tensor:
    """Morgan fingerprints of a MolGraphBatch computed by molecule range on every rank, all-gathered: int32 [n, words]."""
    import numpy as np

    def compute(lo, hi):
        if hi == lo:
            return torch.empty((0, generator.fpSize // 32), dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        return generator.GetFingerprints(graphs.select(np.arange(lo, hi))).torch()

    return map_molecule_range(len(graphs), compute, group)


def sharded_cross_similarity(fp_a: torch.Tensor, fp_b: torch.Tensor, metric: str = "tanimoto", group=None):
    """Rows of A sharded, B replicated; the matrix stays distributed (no collective): returns (rows [hi-lo, m], lo, hi)."""
    from nvmolkit_b200.similarity import crossCosineSimilarity, crossTanimotoSimilarity

    rank, world = rank_world(group)
    lo, hi = molecule_range(fp_a.shape[0], rank, world)
    fn = crossTanimotoSimilarity if metric == "tanimoto" else crossCosineSimilarity
    return fn(fp_a[lo:hi].contiguous(), fp_b).torch(), lo, hi
