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
    if hardwareOptions is None:
        hardwareOptions = HardwareOptions()
    gpu = int(targetGpu) if targetGpu >= 0 else (hardwareOptions.gpuIds[0] if hardwareOptions.gpuIds else torch.cuda.current_device())
    with torch.cuda.device(gpu):
        # size-sorted queue: largest conformers first
        order = np.argsort(-np.diff(batch.atom_starts), kind="stable")
        sizes = np.diff(batch.atom_starts)[order]
        starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        rows = rows_of(batch.atom_starts, order)
        sorted_batch = ConformerBatch(batch.conf_mol[order], starts, batch.positions[rows])
        res = minimize(kind_system, sorted_batch, max_iters, grad_tol)
        # back to input order
        inv = np.argsort(order, kind="stable")
        pos_sorted = res.positions
        dev = pos_sorted.device
        back_rows = rows_of(starts, inv)
        res.positions = pos_sorted[torch.from_numpy(back_rows).to(dev)]
        inv_t = torch.from_numpy(inv.astype(np.int64)).to(dev)
        res.energies, res.status, res.iters = res.energies[inv_t], res.status[inv_t], res.iters[inv_t]
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
