"""Batched MMFF94 optimisation on the GPU. API of ``nvmolkit/mmffOptimization.py`` (reference :60-201)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from nvmolkit_b200._hostutil import rows_of, running_index
from nvmolkit_b200.forcefield import ConformerBatch, FlatSystem
from nvmolkit_b200.minimizer import minimize
from nvmolkit_b200.types import AsyncGpuResult, CoordinateOutput, Device3DResult, HardwareOptions


@dataclass
class FlatMMFFMolecules:
    """Pre-flattened input: MMFF term tables per molecule + the conformers to optimise (the seam below RDKit)."""

    system: FlatSystem
    batch: ConformerBatch


def _device_result(system: FlatSystem, batch: ConformerBatch, res, gpu: int) -> Device3DResult:
    dev = res.positions.device
    conf_idx = running_index(batch.conf_mol)
    t = lambda a: AsyncGpuResult(torch.from_numpy(np.ascontiguousarray(a)).to(dev))  # noqa: E731
    return Device3DResult(AsyncGpuResult(res.positions.reshape(-1, 3)), t(batch.atom_starts), t(batch.conf_mol), t(conf_idx),
                          gpu, system.n_mols, energies=AsyncGpuResult(res.energies),
                          converged=AsyncGpuResult((res.status == 0).to(torch.int8)))


def _optimize(kind_system: FlatSystem, batch: ConformerBatch, max_iters: int, hardwareOptions, output, targetGpu,
              grad_tol: float = 1e-4):
    """Minimise every conformer. With several ``hardwareOptions.gpuIds`` the size-sorted conformer queue is dealt round-robin
    to the devices (the reference spreads its batches over all listed GPUs in one process, src/minimizer/bfgs_mmff.cpp:
    139-157); every device runs its share asynchronously and the results are collected on the target device."""
    if hardwareOptions is None:
        hardwareOptions = HardwareOptions()
    gpus = list(hardwareOptions.gpuIds) if hardwareOptions.gpuIds else [torch.cuda.current_device()]
    gpu = int(targetGpu) if targetGpu >= 0 else gpus[0]
    order = np.argsort(-np.diff(batch.atom_starts), kind="stable")  # largest conformers first: evens out the persistent CTAs' tail
    shares = [order[d::len(gpus)] for d in range(len(gpus))]
    parts = []
    for dev_id, share in zip(gpus, shares):
        if len(share) == 0:
            continue
        with torch.cuda.device(dev_id):
            sizes = np.diff(batch.atom_starts)[share]
            starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
            sub = ConformerBatch(batch.conf_mol[share], starts, batch.positions[rows_of(batch.atom_starts, share)])
            parts.append((share, starts, minimize(kind_system, sub, max_iters, grad_tol)))  # asynchronous on that device
    with torch.cuda.device(gpu):
        dev = torch.device("cuda", gpu)
        n_conf, n_atoms = batch.n_conf, int(batch.atom_starts[-1])
        positions = torch.empty((n_atoms, 3), dtype=torch.float64, device=dev)
        energies = torch.empty(n_conf, dtype=torch.float64, device=dev)
        status = torch.empty(n_conf, dtype=torch.int8, device=dev)
        iters = torch.empty(n_conf, dtype=torch.int32, device=dev)
        for share, starts, r in parts:  # back to input order (peer copies when the part ran on another device)
            idx = torch.from_numpy(share.astype(np.int64)).to(dev)
            rows = torch.from_numpy(rows_of(batch.atom_starts, share)).to(dev)
            positions[rows] = r.positions.reshape(-1, 3).to(dev)
            energies[idx], status[idx], iters[idx] = r.energies.to(dev), r.status.to(dev), r.iters.to(dev)
        from nvmolkit_b200.minimizer import MinimizeResult

        res = MinimizeResult(positions, energies, status, iters)
        if output == CoordinateOutput.DEVICE:
            return _device_result(kind_system, batch, res, gpu)
        energies = res.energies.cpu().numpy()
        positions = res.positions.cpu().numpy()
    out: List[List[float]] = [[] for _ in range(kind_system.n_mols)]
    coords: List[List[np.ndarray]] = [[] for _ in range(kind_system.n_mols)]
    for c, m in enumerate(batch.conf_mol):
        out[int(m)].append(float(energies[c]))
        coords[int(m)].append(positions[batch.atom_starts[c]:batch.atom_starts[c + 1]])
    return out, coords


def MMFFOptimizeMoleculesConfs(molecules, maxIters: int = 200, properties=None, nonBondedThreshold=100.0,
                               ignoreInterfragInteractions=True, hardwareOptions: Optional[HardwareOptions] = None,
                               output: CoordinateOutput = CoordinateOutput.RDKIT_CONFORMERS, targetGpu: int = -1):
    """Optimise every conformer of every molecule with MMFF94 + BFGS (gradTol 1e-4, like the reference).

    RDKit molecules: conformers are updated in place and a list of per-molecule energy lists is returned
    (``DEVICE``: a :class:`Device3DResult` with coordinates, energies and converged flags). With a pre-flattened
    :class:`FlatMMFFMolecules` the optimised coordinates cannot be written into RDKit objects; RDKIT_CONFORMERS mode then
    returns ``(energies, coordinates)`` as nested lists.
    """
    if isinstance(molecules, FlatMMFFMolecules):
        return _optimize(molecules.system, molecules.batch, maxIters, hardwareOptions, output, targetGpu)
    if not molecules:
        if output == CoordinateOutput.DEVICE:
            raise ValueError("MMFFOptimizeMoleculesConfs(output=DEVICE) requires at least one molecule")
        return []
    from nvmolkit_b200.rdkit_adapter import mmff_from_rdkit, write_back_conformers

    flat = mmff_from_rdkit(molecules, properties, nonBondedThreshold, ignoreInterfragInteractions)
    result = _optimize(flat.system, flat.batch, maxIters, hardwareOptions, output, targetGpu)
    if output == CoordinateOutput.DEVICE:
        return result
    energies, coords = result
    write_back_conformers(molecules, coords)
    return energies
