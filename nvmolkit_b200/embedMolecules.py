"""ETKDG conformer generation on the GPU. API of ``nvmolkit/embedMolecules.py`` (reference :55-158).

``molecules`` is either a list of RDKit molecules (flattened through ``nvmolkit_b200.rdkit_adapter`` when RDKit is
importable) or a pre-flattened :class:`FlatEmbedMolecules` — the seam below RDKit (bounds matrices already turned into
DG terms, experimental-torsion terms, chiral sets). The whole attempt pipeline runs in one persistent kernel.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from nvmolkit_b200 import _lib
from nvmolkit_b200._hostutil import rows_of, running_index
from nvmolkit_b200._interop import require_cuda, stream_ctx, stream_ptr
from nvmolkit_b200.forcefield import CheckTables, FlatSystem
from nvmolkit_b200.types import AsyncGpuResult, CoordinateOutput, Device3DResult, HardwareOptions

STAGES = ("coordgen", "first_minimize_energy", "tetrahedral", "first_chirality", "fourth_dim_minimize", "etk_planarity",
          "double_bond_geometry", "final_chirality", "chiral_dist_matrix", "chiral_centre_volume", "double_bond_stereo")


class EmbedParamsC(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("boxSize", C.c_double), ("optimizerForceTol", C.c_double),
                ("enforceChirality", C.c_int32), ("useExpTorsions", C.c_int32), ("useBasicKnowledge", C.c_int32),
                ("maxAttempts", C.c_int32), ("dgIters", C.c_int32), ("fourthIters", C.c_int32), ("etkIters", C.c_int32),
                ("maxRestarts", C.c_int32), ("useMetricStart", C.c_int32)]


class EmbedParameters:
    """The RDKit ``EmbedParameters`` fields this path reads (ETKDGv3 defaults). An RDKit object works as well."""

    def __init__(self, useRandomCoords: bool = True, randomSeed: int = -1, boxSizeMult: float = 2.0,
                 optimizerForceTol: float = 1e-3, enforceChirality: bool = True, useExpTorsionAnglePrefs: bool = True,
                 useBasicKnowledge: bool = True, pruneRmsThresh: float = -1.0):
        self.useRandomCoords = useRandomCoords
        self.randomSeed = randomSeed
        self.boxSizeMult = boxSizeMult
        self.optimizerForceTol = optimizerForceTol
        self.enforceChirality = enforceChirality
        self.useExpTorsionAnglePrefs = useExpTorsionAnglePrefs
        self.useBasicKnowledge = useBasicKnowledge
        self.pruneRmsThresh = pruneRmsThresh


@dataclass
class FlatEmbedMolecules:
    dg: FlatSystem
    etk: FlatSystem
    checks: CheckTables
    prune_matches: Optional[list] = None  # per molecule: [K, L] atom-index lists for RMS pruning (None = all atoms)

    def __len__(self) -> int:
        return self.dg.n_mols

    @property
    def atom_counts(self) -> np.ndarray:
        return self.dg.atom_counts

    @classmethod
    def concat(cls, parts: "List[FlatEmbedMolecules]") -> "FlatEmbedMolecules":
        return cls(FlatSystem.concat([p.dg for p in parts]), FlatSystem.concat([p.etk for p in parts]),
                   CheckTables.concat([p.checks for p in parts]))

    def nbytes(self) -> int:
        return self.dg.nbytes() + self.etk.nbytes() + self.checks.nbytes()

    def drop_device(self) -> None:
        """Forget the uploaded copies (the next call uploads the tables again: bench.py's end-to-end leg)."""
        self.dg._device.clear()
        self.etk._device.clear()
        self.checks._device.clear()


def _params_struct(params, max_attempts: int, seed_fallback: int = 0xB200) -> EmbedParamsC:
    box = float(params.boxSizeMult)
    seed = int(getattr(params, "randomSeed", -1))
    return EmbedParamsC(seed=seed if seed >= 0 else seed_fallback, boxSize=5.0 * box if box > 0 else -box,
                        optimizerForceTol=float(params.optimizerForceTol),
                        enforceChirality=int(bool(params.enforceChirality)),
                        useExpTorsions=int(bool(params.useExpTorsionAnglePrefs)),
                        useBasicKnowledge=int(bool(params.useBasicKnowledge)), maxAttempts=int(max_attempts),
                        dgIters=400, fourthIters=200, etkIters=300, maxRestarts=20,
                        useMetricStart=0 if bool(getattr(params, "useRandomCoords", True)) else 1)


@dataclass
class EmbedRaw:
    coords: torch.Tensor  # float64 [nSlots atoms, 3] (rows of failed slots undefined)
    ok: torch.Tensor  # int8 [nSlots]
    attempts: torch.Tensor  # int32 [nSlots]
    energy: torch.Tensor  # float64 [nSlots]
    stage_failures: torch.Tensor  # int64 [11]
    slot_mol: np.ndarray
    slot_atom_start: np.ndarray


def embed_slots(flat: FlatEmbedMolecules, params, confs_per_molecule: int, max_iterations: int = -1, stream=None,
                mol_indices: Optional[np.ndarray] = None) -> EmbedRaw:
    """Run the embedding kernel: one slot per requested conformer of each molecule in `mol_indices` (default: all)."""
    sptr = stream_ptr(stream)
    require_cuda()
    mols = np.arange(len(flat), dtype=np.int32) if mol_indices is None else np.asarray(mol_indices, dtype=np.int32)
    counts = flat.atom_counts[mols]
    # largest molecules first (reference sorts by size, src/etkdg.cpp:151-154): evens out the persistent CTAs' tail
    order = np.argsort(-counts, kind="stable")
    slot_mol = np.repeat(mols[order], confs_per_molecule).astype(np.int32)
    slot_atoms = flat.atom_counts[slot_mol]
    slot_start = np.concatenate([[0], np.cumsum(slot_atoms)]).astype(np.int32)
    max_atoms = int(counts.max(initial=1))
    if max_iterations is None or max_iterations < 0:
        max_iterations = 10 * max_atoms  # src/etkdg.cpp:71-85
    pc = _params_struct(params, max_iterations)
    dev = torch.device("cuda", torch.cuda.current_device())
    dg, _k1 = flat.dg.to_device(dev)
    etk, _k2 = flat.etk.to_device(dev)
    chk, _k3 = flat.checks.to_device(dev)
    n = len(slot_mol)
    with stream_ctx(stream):
        d_mol = torch.from_numpy(slot_mol).to(dev)
        d_start = torch.from_numpy(slot_start).to(dev)
        coords = torch.zeros((int(slot_start[-1]), 3), dtype=torch.float64, device=dev)
        ok = torch.zeros(n, dtype=torch.int8, device=dev)
        attempts = torch.zeros(n, dtype=torch.int32, device=dev)
        energy = torch.zeros(n, dtype=torch.float64, device=dev)
        fails = torch.zeros(len(STAGES), dtype=torch.int64, device=dev)
        _lib.call("b200mol_etkdg_embed", C.byref(dg), C.byref(etk), C.byref(chk), C.byref(pc), n, d_mol.data_ptr(),
                  d_start.data_ptr(), max_atoms, coords.data_ptr(), ok.data_ptr(), attempts.data_ptr(), energy.data_ptr(),
                  fails.data_ptr(), sptr)
    return EmbedRaw(coords, ok, attempts, energy, fails, slot_mol, slot_start)


def _prune(raw: EmbedRaw, flat: "FlatEmbedMolecules", rms_thresh: float) -> torch.Tensor:
    """RDKit's pruneRmsThresh on the device: the slots of a molecule are contiguous and in embedding order."""
    from nvmolkit_b200.pruning import rms_prune

    change = np.nonzero(np.diff(raw.slot_mol) != 0)[0] + 1
    mol_conf_start = np.concatenate([[0], change, [len(raw.slot_mol)]]).astype(np.int32)
    mols = raw.slot_mol[mol_conf_start[:-1]]
    matches = None
    if getattr(flat, "prune_matches", None) is not None:
        matches = [flat.prune_matches[int(m)] for m in mols]
    return rms_prune(raw.coords, raw.slot_atom_start, mol_conf_start, rms_thresh, matches, flat.atom_counts[mols], valid=raw.ok)


def _to_device_result(raw: EmbedRaw, n_mols: int, gpu_id: int, keep: Optional[torch.Tensor] = None) -> Device3DResult:
    """Compact the successful slots into the reference's CSR result (src/conformer/device_coord_result.h:58-67)."""
    ok_h = (raw.ok if keep is None else keep).cpu().numpy().astype(bool)
    sizes = np.diff(raw.slot_atom_start)
    keep = np.nonzero(ok_h)[0]
    # present results in input molecule order, conformer order within a molecule
    keep = keep[np.argsort(raw.slot_mol[keep], kind="stable")]
    dev = raw.coords.device
    rows = rows_of(raw.slot_atom_start, keep)
    values = raw.coords[torch.from_numpy(rows).to(dev)]
    starts = np.concatenate([[0], np.cumsum(sizes[keep])]).astype(np.int32)
    mol_idx = raw.slot_mol[keep].astype(np.int32)
    conf_idx = running_index(mol_idx)
    t = lambda a: AsyncGpuResult(torch.from_numpy(a).to(dev))  # noqa: E731
    return Device3DResult(AsyncGpuResult(values), t(starts), t(mol_idx), t(conf_idx), gpu_id, n_mols)


def EmbedMolecules(molecules, params, confsPerMolecule: int = 1, maxIterations: int = -1,
                   hardwareOptions: Optional[HardwareOptions] = None,
                   output: CoordinateOutput = CoordinateOutput.RDKIT_CONFORMERS, targetGpu: int = -1):
    """Embed `confsPerMolecule` conformers of every molecule.

    RDKit molecules: RDKIT_CONFORMERS writes conformers into the molecules and returns None; DEVICE returns a
    :class:`Device3DResult`. A :class:`FlatEmbedMolecules` input returns, in RDKIT_CONFORMERS mode, a list (per
    molecule) of lists of ``(nAtoms, 3)`` NumPy coordinate arrays instead of mutating RDKit objects.
    """
    flat_input = isinstance(molecules, FlatEmbedMolecules)
    if not flat_input and not molecules:
        if output == CoordinateOutput.DEVICE:
            raise ValueError("EmbedMolecules(output=DEVICE) requires at least one molecule")
        return None
    if not flat_input:
        for i, mol in enumerate(molecules):
            if mol is None:
                raise ValueError(f"Molecule at index {i} is None")
    # useRandomCoords=False (refused by the reference, src/etkdg.cpp:99-101) selects the on-device metric-matrix start
    # (the reference refuses pruneRmsThresh > 0 with DEVICE output, src/etkdg.cpp:106-110: its pruning is host-only;
    #  here the pruning runs on the device, so both outputs take it)
    if hardwareOptions is None:
        hardwareOptions = HardwareOptions()
    gpu = int(targetGpu) if targetGpu >= 0 else (hardwareOptions.gpuIds[0] if hardwareOptions.gpuIds else torch.cuda.current_device())
    if len(hardwareOptions.gpuIds) > 1:
        import warnings

        warnings.warn("EmbedMolecules runs on one GPU per process (the first of hardwareOptions.gpuIds, or targetGpu); shard "
                      "molecule ranges over ranks with torch.distributed (nvmolkit_b200.distributed.molecule_range) to use "
                      f"all of {list(hardwareOptions.gpuIds)}", RuntimeWarning, stacklevel=2)
    if flat_input:
        flat = molecules
    else:
        from nvmolkit_b200.rdkit_adapter import embed_molecules_from_rdkit

        flat = embed_molecules_from_rdkit(molecules, params)
    with torch.cuda.device(gpu):
        raw = embed_slots(flat, params, int(confsPerMolecule), int(maxIterations))
        prune = float(getattr(params, "pruneRmsThresh", -1.0))
        keep = _prune(raw, flat, prune) if prune > 0.0 else None
        if output == CoordinateOutput.DEVICE:
            return _to_device_result(raw, len(flat), gpu, keep)
        res = _to_device_result(raw, len(flat), gpu, keep)
        per_mol: List[List[np.ndarray]] = [[c.cpu().numpy() for c in confs] for confs in res.per_molecule()]
    if flat_input:
        return per_mol
    from nvmolkit_b200.rdkit_adapter import add_conformers

    add_conformers(molecules, per_mol, -1.0)  # (already pruned on the device)
    return None
