"""Batched force-field objects with restraints. API of ``nvmolkit/batchedForcefield.py`` (reference :97-714):
``MMFFBatchedForcefield`` / ``UFFBatchedForcefield`` over a list of molecules, ``ff[i].add_*_constraint(...)``,
``compute_energy()``, ``compute_gradients()``, ``minimize()``.

Molecules are RDKit molecules (flattened through ``rdkit_adapter``) or the pre-flattened ``FlatMMFFMolecules`` /
``FlatUFFMolecules``. Restraint terms (distance, position, angle, torsion; reference term math
``src/forcefields/mmff_kernels_device.cuh:673-1036``, specs ``src/forcefields/forcefield_constraints.{h,cpp}``) are extra
term tables of the same flattened system, evaluated by the same kernels. A ``relative`` restraint and every position
restraint is anchored on a conformer's own starting geometry, so the object keeps ONE table entry per conformer (the plain
optimisers share one entry among a molecule's conformers).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from nvmolkit_b200.forcefield import LAYOUT, ConformerBatch, FlatSystem
from nvmolkit_b200.minimizer import energy_and_grad, minimize
from nvmolkit_b200.types import AsyncGpuResult, CoordinateOutput, Device3DResult, HardwareOptions

_RAD2DEG = 180.0 / np.pi


@dataclass
class _DistanceConstraint:
    idx1: int
    idx2: int
    relative: bool
    min_len: float
    max_len: float
    force_constant: float


@dataclass
class _PositionConstraint:
    idx: int
    max_displ: float
    force_constant: float


@dataclass
class _AngleConstraint:
    idx1: int
    idx2: int
    idx3: int
    relative: bool
    min_angle_deg: float
    max_angle_deg: float
    force_constant: float


@dataclass
class _TorsionConstraint:
    idx1: int
    idx2: int
    idx3: int
    idx4: int
    relative: bool
    min_dihedral_deg: float
    max_dihedral_deg: float
    force_constant: float


def normalize_angle_deg(a: float) -> float:
    """forcefield_constraints.cpp:79-87."""
    a = float(np.fmod(a, 360.0))
    if a < -180.0:
        a += 360.0
    elif a > 180.0:
        a -= 360.0
    return a


def angle_deg(xyz: np.ndarray, i: int, j: int, k: int) -> float:
    """forcefield_constraints.cpp:65-77."""
    r1, r2 = xyz[i] - xyz[j], xyz[k] - xyz[j]
    l1, l2 = max(1.0e-5, float(r1 @ r1)), max(1.0e-5, float(r2 @ r2))
    return _RAD2DEG * float(np.arccos(np.clip(float(r1 @ r2) / np.sqrt(l1 * l2), -1.0, 1.0)))


def dihedral_deg(xyz: np.ndarray, i: int, j: int, k: int, l: int) -> float:
    """Signed dihedral, forcefield_constraints.cpp:89-127."""
    r0, r1 = xyz[i] - xyz[j], xyz[k] - xyz[j]
    r2, r3 = -r1, xyz[l] - xyz[k]
    t0, t1 = np.cross(r0, r1), np.cross(r2, r3)
    t0 = t0 / max(float(np.linalg.norm(t0)), 1.0e-5)
    t1 = t1 / max(float(np.linalg.norm(t1)), 1.0e-5)
    cos_phi = float(np.clip(t0 @ t1, -1.0, 1.0))
    m = np.cross(t0, r1)
    return _RAD2DEG * -float(np.arctan2(float(m @ t1) / max(float(np.linalg.norm(m)), 1.0e-5), cos_phi))


class _BatchElement:
    """Per-molecule view for adding restraints (reference: _BatchElementBase, :171-289). They apply to all conformers of
    the molecule."""

    def __init__(self, parent, idx: int):
        self._parent, self._idx = parent, idx

    @property
    def num_atoms(self) -> int:
        return int(self._parent._atom_counts[self._idx])

    def add_distance_constraint(self, idx1: int, idx2: int, relative: bool, min_len: float, max_len: float, force_constant: float) -> None:
        self._parent._validate_atom_indices(self._idx, idx1, idx2)
        self._parent._constraints[self._idx].append(_DistanceConstraint(idx1, idx2, bool(relative), min_len, max_len, force_constant))
        self._parent._dirty = True

    def add_position_constraint(self, idx: int, max_displ: float, force_constant: float) -> None:
        self._parent._validate_atom_indices(self._idx, idx)
        self._parent._constraints[self._idx].append(_PositionConstraint(idx, max_displ, force_constant))
        self._parent._dirty = True

    def add_angle_constraint(self, idx1: int, idx2: int, idx3: int, relative: bool, min_angle_deg: float, max_angle_deg: float,
                             force_constant: float) -> None:
        self._parent._validate_atom_indices(self._idx, idx1, idx2, idx3)
        self._parent._constraints[self._idx].append(
            _AngleConstraint(idx1, idx2, idx3, bool(relative), min_angle_deg, max_angle_deg, force_constant))
        self._parent._dirty = True

    def add_torsion_constraint(self, idx1: int, idx2: int, idx3: int, idx4: int, relative: bool, min_dihedral_deg: float,
                               max_dihedral_deg: float, force_constant: float) -> None:
        self._parent._validate_atom_indices(self._idx, idx1, idx2, idx3, idx4)
        self._parent._constraints[self._idx].append(
            _TorsionConstraint(idx1, idx2, idx3, idx4, bool(relative), min_dihedral_deg, max_dihedral_deg, force_constant))
        self._parent._dirty = True


def restraint_tables(constraints, xyz: np.ndarray):
    """The four restraint tables of ONE conformer from the molecule's restraint specs and that conformer's coordinates
    (append*ConstraintImpl, forcefield_constraints.cpp:129-226: relative windows are offset by the current value, the
    dihedral window is normalised to (-180, 180], position restraints are anchored on the current position)."""
    out = {"distc": ([], []), "posc": ([], []), "anglec": ([], []), "torsc": ([], [])}
    for c in constraints:
        if isinstance(c, _DistanceConstraint):
            mn, mx = float(c.min_len), float(c.max_len)
            if mx < mn:
                raise ValueError("Distance constraint maxLen must be >= minLen")
            if c.relative:
                d = float(np.linalg.norm(xyz[c.idx1] - xyz[c.idx2]))
                mn, mx = max(mn + d, 0.0), max(mx + d, 0.0)
            out["distc"][0].append((c.idx1, c.idx2))
            out["distc"][1].append((mn, mx, c.force_constant))
        elif isinstance(c, _PositionConstraint):
            out["posc"][0].append((c.idx,))
            out["posc"][1].append((*xyz[c.idx].tolist(), c.max_displ, c.force_constant))
        elif isinstance(c, _AngleConstraint):
            mn, mx = float(c.min_angle_deg), float(c.max_angle_deg)
            if mx < mn:
                raise ValueError("Angle constraint maxAngleDeg must be >= minAngleDeg")
            if c.relative:
                a = angle_deg(xyz, c.idx1, c.idx2, c.idx3)
                mn, mx = mn + a, mx + a
            if not (0.0 <= mn <= 180.0 and 0.0 <= mx <= 180.0):
                raise ValueError("Angle constraint bounds must be within [0, 180]")
            out["anglec"][0].append((c.idx1, c.idx2, c.idx3))
            out["anglec"][1].append((mn, mx, c.force_constant))
        else:
            mn, mx = float(c.min_dihedral_deg), float(c.max_dihedral_deg)
            if mx < mn:
                raise ValueError("Torsion constraint maxDihedralDeg must be >= minDihedralDeg")
            if c.relative:
                d = dihedral_deg(xyz, c.idx1, c.idx2, c.idx3, c.idx4)
                mn, mx = mn + d, mx + d
            out["torsc"][0].append((c.idx1, c.idx2, c.idx3, c.idx4))
            out["torsc"][1].append((normalize_angle_deg(mn), normalize_angle_deg(mx), c.force_constant))
    return out


class _BatchedForcefieldBase:
    kind = ""

    def _init_common(self, flat, hardwareOptions: Optional[HardwareOptions]):
        self._base: FlatSystem = flat.system
        self._batch: ConformerBatch = flat.batch
        self._atom_counts = self._base.atom_counts
        self._hardware_options = hardwareOptions if hardwareOptions is not None else HardwareOptions()
        self._constraints: List[list] = [[] for _ in range(self._base.n_mols)]
        self._system: Optional[FlatSystem] = None
        self._dirty = True
        self.num_molecules = self._base.n_mols
        self.data_dim = 3

    def __len__(self) -> int:
        return self.num_molecules

    def __getitem__(self, idx: int) -> _BatchElement:
        if idx < 0 or idx >= self.num_molecules:
            raise IndexError(f"Batch element index {idx} out of range")
        return _BatchElement(self, idx)

    def _validate_atom_indices(self, batch_idx: int, *indices: int) -> None:
        n = int(self._atom_counts[batch_idx])
        for idx in indices:
            if idx < 0 or idx >= n:
                raise IndexError(f"Atom index {idx} out of range for molecule {batch_idx} with {n} atoms")

    def _build(self) -> None:
        """One table entry per conformer: the molecule's terms + that conformer's restraint tables."""
        b, base = self._batch, self._base
        per_conf, counts = [], []
        for c in range(b.n_conf):
            m = int(b.conf_mol[c])
            xyz = b.positions[b.atom_starts[c]:b.atom_starts[c + 1]]
            terms = {name: (ix[st[m]:st[m + 1]], pr[st[m]:st[m + 1]]) for name, (st, ix, pr) in base.tables.items()
                     if name not in ("distc", "posc", "anglec", "torsc")}
            for name, (ix, pr) in restraint_tables(self._constraints[m], xyz).items():
                k, p = next((kk, pp) for n2, kk, pp in LAYOUT[self.kind] if n2 == name)
                terms[name] = (np.array(ix, dtype=np.int16).reshape(-1, k), np.array(pr, dtype=np.float64).reshape(-1, p))
            per_conf.append(terms)
            counts.append(int(self._atom_counts[m]))
        self._system = FlatSystem.from_molecules(self.kind, counts, per_conf)
        self._conf_batch = ConformerBatch(np.arange(b.n_conf, dtype=np.int32), b.atom_starts, b.positions)
        self._dirty = False

    def _ensure_built(self) -> None:
        if self._dirty or self._system is None:
            self._build()

    def rebuild(self) -> None:
        self._build()

    def _per_molecule(self, values) -> list:
        out: List[list] = [[] for _ in range(self.num_molecules)]
        for c, m in enumerate(self._batch.conf_mol):
            out[int(m)].append(values[c])
        return out

    def compute_energy(self) -> List[List[float]]:
        """``result[mol][conf]`` (reference :402-411)."""
        if self.num_molecules == 0:
            return []
        self._ensure_built()
        e, _ = energy_and_grad(self._system, self._conf_batch, want_grad=False)
        return self._per_molecule([float(v) for v in e.cpu().numpy()])

    def compute_gradients(self) -> List[List[List[float]]]:
        """``result[mol][conf]`` = flattened [x0, y0, z0, ...] gradient (reference :413-423)."""
        if self.num_molecules == 0:
            return []
        self._ensure_built()
        _, g = energy_and_grad(self._system, self._conf_batch, want_grad=True)
        g = g.cpu().numpy()
        st = self._batch.atom_starts
        return self._per_molecule([g[st[c]:st[c + 1]].ravel().tolist() for c in range(self._batch.n_conf)])

    def _minimize(self, maxIters: int, forceTol: float, output: CoordinateOutput, target_gpu: Optional[int]):
        if self.num_molecules == 0:
            if output == CoordinateOutput.DEVICE:
                raise ValueError("minimize(output=DEVICE) requires at least one molecule")
            return [], []
        self._ensure_built()
        gpu = int(target_gpu) if target_gpu is not None and target_gpu >= 0 else (
            self._hardware_options.gpuIds[0] if self._hardware_options.gpuIds else torch.cuda.current_device())
        with torch.cuda.device(gpu):
            res = minimize(self._system, self._conf_batch, int(maxIters), float(forceTol))
            if output == CoordinateOutput.DEVICE:
                from nvmolkit_b200._hostutil import running_index

                dev = res.positions.device
                t = lambda a: AsyncGpuResult(torch.from_numpy(np.ascontiguousarray(a)).to(dev))  # noqa: E731
                return Device3DResult(AsyncGpuResult(res.positions.reshape(-1, 3)), t(self._batch.atom_starts), t(self._batch.conf_mol),
                                      t(running_index(self._batch.conf_mol)), gpu, self.num_molecules,
                                      energies=AsyncGpuResult(res.energies), converged=AsyncGpuResult((res.status == 0).to(torch.int8)))
            energies = res.energies.cpu().numpy()
            converged = (res.status.cpu().numpy() == 0)
            pos = res.positions.cpu().numpy()
        # the optimised coordinates become the conformers' coordinates (the reference writes them into the RDKit molecules)
        self._batch = ConformerBatch(self._batch.conf_mol, self._batch.atom_starts, pos)
        self._conf_batch = ConformerBatch(self._conf_batch.conf_mol, self._conf_batch.atom_starts, pos)
        if self._rdkit_molecules is not None:
            from nvmolkit_b200.rdkit_adapter import write_back_conformers

            st = self._batch.atom_starts
            write_back_conformers(self._rdkit_molecules, self._per_molecule([pos[st[c]:st[c + 1]] for c in range(self._batch.n_conf)]))
        return (self._per_molecule([float(v) for v in energies]), self._per_molecule([bool(v) for v in converged]))

    def positions(self) -> List[List[np.ndarray]]:
        """Current coordinates ``[mol][conf] -> (nAtoms, 3)`` (after minimize(): the optimised ones)."""
        st = self._batch.atom_starts
        return self._per_molecule([self._batch.positions[st[c]:st[c + 1]].copy() for c in range(self._batch.n_conf)])


class MMFFBatchedForcefield(_BatchedForcefieldBase):
    """MMFF94 over a list of molecules (reference :443-598)."""

    kind = "mmff"

    def __init__(self, molecules, properties=None, nonBondedThreshold: float = 100.0, ignoreInterfragInteractions=True,
                 hardwareOptions: Optional[HardwareOptions] = None):
        from nvmolkit_b200.mmffOptimization import FlatMMFFMolecules

        self._rdkit_molecules = None
        if isinstance(molecules, FlatMMFFMolecules):
            flat = molecules
        else:
            from nvmolkit_b200.rdkit_adapter import mmff_from_rdkit

            self._rdkit_molecules = list(molecules)
            flat = mmff_from_rdkit(self._rdkit_molecules, properties, nonBondedThreshold, bool(np.all(ignoreInterfragInteractions)))
        self._init_common(flat, hardwareOptions)

    def minimize(self, maxIters: int = 200, forceTol: float = 1e-4, output: CoordinateOutput = CoordinateOutput.RDKIT_CONFORMERS,
                 targetGpu: Optional[int] = None):
        """BFGS minimisation of every conformer; returns ``(energies[mol][conf], converged[mol][conf])`` or, with
        ``output=DEVICE``, a :class:`Device3DResult`."""
        return self._minimize(maxIters, forceTol, output, targetGpu)


class UFFBatchedForcefield(_BatchedForcefieldBase):
    """UFF over a list of molecules (reference :601-714)."""

    kind = "uff"

    def __init__(self, molecules, vdwThreshold: float = 10.0, ignoreInterfragInteractions=True,
                 hardwareOptions: Optional[HardwareOptions] = None):
        from nvmolkit_b200.uffOptimization import FlatUFFMolecules

        self._rdkit_molecules = None
        if isinstance(molecules, FlatUFFMolecules):
            flat = molecules
        else:
            from nvmolkit_b200.rdkit_adapter import uff_from_rdkit

            self._rdkit_molecules = list(molecules)
            flat = uff_from_rdkit(self._rdkit_molecules, vdwThreshold, bool(np.all(ignoreInterfragInteractions)))
        self._init_common(flat, hardwareOptions)

    def minimize(self, maxIters: int = 1000, forceTol: float = 1e-4, output: CoordinateOutput = CoordinateOutput.RDKIT_CONFORMERS,
                 targetGpu: Optional[int] = None):
        return self._minimize(maxIters, forceTol, output, targetGpu)
