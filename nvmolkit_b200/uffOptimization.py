"""Batched UFF optimisation on the GPU. API of ``nvmolkit/uffOptimization.py`` (reference :36-142)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from nvmolkit_b200.forcefield import ConformerBatch, FlatSystem
from nvmolkit_b200.mmffOptimization import _optimize
from nvmolkit_b200.types import CoordinateOutput, HardwareOptions


@dataclass
class FlatUFFMolecules:
    """Pre-flattened input: UFF term tables per molecule + the conformers to optimise."""

    system: FlatSystem
    batch: ConformerBatch


def UFFOptimizeMoleculesConfs(molecules, maxIters: int = 1000, vdwThreshold=10.0, ignoreInterfragInteractions=True,
                              hardwareOptions: Optional[HardwareOptions] = None,
                              output: CoordinateOutput = CoordinateOutput.RDKIT_CONFORMERS, targetGpu: int = -1):
    """Optimise every conformer of every molecule with UFF + BFGS (gradTol 1e-4). Same conventions as
    :func:`nvmolkit_b200.mmffOptimization.MMFFOptimizeMoleculesConfs`."""
    if isinstance(molecules, FlatUFFMolecules):
        return _optimize(molecules.system, molecules.batch, maxIters, hardwareOptions, output, targetGpu)
    if not molecules:
        if output == CoordinateOutput.DEVICE:
            raise ValueError("UFFOptimizeMoleculesConfs(output=DEVICE) requires at least one molecule")
        return []
    from nvmolkit_b200.rdkit_adapter import uff_from_rdkit, write_back_conformers

    flat = uff_from_rdkit(molecules, vdwThreshold, ignoreInterfragInteractions)
    result = _optimize(flat.system, flat.batch, maxIters, hardwareOptions, output, targetGpu)
    if output == CoordinateOutput.DEVICE:
        return result
    energies, coords = result
    write_back_conformers(molecules, coords)
    return energies
