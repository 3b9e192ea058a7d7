"""RDKit -> flat-table adapters (the CPU seam above the C-ABI).

Imported only when RDKit molecules are handed to the public API; RDKit itself is imported lazily, so the package works
without it on pre-flattened inputs. RDKit is absent from the container this repository was built in: this module is
written against RDKit's documented Python API (2025.03 .. 2026.03: ``rdForceFieldHelpers.GetMMFF*Params``,
``MMFFGetMoleculeProperties``) and has NOT been executed here — DESIGN.md §2 / §6 list it as unpinned.

What it mirrors: ``MMFF::constructForcefieldContribs`` (rdkit_extensions/mmff_flattened_builder.cpp:453-556), i.e.
RDKit's own MMFF builder: bond, angle, stretch-bend, out-of-plane, torsion terms by topology; van der Waals and
electrostatic pairs for every atom pair three or more bonds apart (1-4 pairs flagged) within ``nonBondedThreshold`` of
the conformer geometry. One term block per molecule is shared by all its conformers (include/b200mol.h), so the pair
list is the UNION over the molecule's conformers of the pairs within the threshold (the reference builds one list per
conformer; with the default threshold of 100 A the two coincide).

UFF and ETKDG (bounds matrix, experimental torsions, chiral sets) need RDKit internals that its Python API does not
expose (UFF angle orders / inversion coefficients, `findChiralSets`); those entry points raise NotImplementedError and
name the pre-flattened input classes to use instead.
"""

from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from nvmolkit_b200.forcefield import ConformerBatch, FlatSystem

_LINEAR_MMFF_TYPES = frozenset({4, 53, 61})  # MMFFPROP.PAR `linh` = 1: =C=/-C#, =N=, isonitrile N
_TORSION_BOND_SMARTS = "[!$([D1]);!$(*#*)]~[!$([D1]);!$(*#*)]"  # RDKit DefaultTorsionBondSmarts


def _rdkit():
    try:
        from rdkit import Chem
        from rdkit.Chem import rdForceFieldHelpers as FFH
    except ImportError as e:  # pragma: no cover - RDKit is optional
        raise ImportError("RDKit molecules were passed but RDKit is not importable; pass pre-flattened "
                          "Flat*Molecules instead") from e
    return Chem, FFH


def _conformer_coords(mol) -> List[np.ndarray]:
    return [np.asarray(c.GetPositions(), dtype=np.float64) for c in mol.GetConformers()]


def _mmff_terms(Chem, FFH, mol, props, coords: Sequence[np.ndarray], non_bonded_thresh: float, ignore_interfrag: bool
                ) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    n = mol.GetNumAtoms()
    nbrs = [[a.GetIdx() for a in atom.GetNeighbors()] for atom in mol.GetAtoms()]
    types = [props.GetMMFFAtomType(i) for i in range(n)]
    t: Dict[str, Tuple[list, list]] = {k: ([], []) for k in ("bond", "angle", "strbend", "oop", "torsion", "vdw", "ele")}

    bond_r0: Dict[Tuple[int, int], float] = {}
    for b in mol.GetBonds():
        i, j = b.GetBeginAtomIdx(), b.GetEndAtomIdx()
        p = FFH.GetMMFFBondStretchParams(mol, i, j)  # (bondType, kb, r0)
        if p:
            t["bond"][0].append((i, j))
            t["bond"][1].append((p[2], p[1]))
            bond_r0[(i, j)] = bond_r0[(j, i)] = p[2]

    for j in range(n):
        if len(nbrs[j]) < 2:
            continue
        linear = types[j] in _LINEAR_MMFF_TYPES
        for a, i in enumerate(nbrs[j]):
            for k in nbrs[j][a + 1:]:
                pa = FFH.GetMMFFAngleBendParams(mol, i, j, k)  # (angleType, ka, theta0)
                if pa:
                    t["angle"][0].append((i, j, k))
                    t["angle"][1].append((pa[2], pa[1], 1.0 if linear else 0.0))
                if linear or not pa:
                    continue
                ps = FFH.GetMMFFStretchBendParams(mol, i, j, k)  # (stretchBendType, kbaIJK, kbaKJI)
                if ps and (i, j) in bond_r0 and (k, j) in bond_r0:
                    t["strbend"][0].append((i, j, k))
                    t["strbend"][1].append((pa[2], bond_r0[(i, j)], bond_r0[(k, j)], ps[1], ps[2]))

    for j in range(n):
        if len(nbrs[j]) != 3:
            continue
        i, k, l = nbrs[j]
        koop = FFH.GetMMFFOopBendParams(mol, i, j, k, l)
        if koop is None:
            continue
        for quad in ((i, j, k, l), (i, j, l, k), (k, j, l, i)):  # the three Wilson angles around the centre j
            t["oop"][0].append(quad)
            t["oop"][1].append((float(koop),))

    sp23 = (Chem.HybridizationType.SP2, Chem.HybridizationType.SP3)
    query = Chem.MolFromSmarts(_TORSION_BOND_SMARTS)
    for j, k in mol.GetSubstructMatches(query):
        if mol.GetAtomWithIdx(j).GetHybridization() not in sp23 or mol.GetAtomWithIdx(k).GetHybridization() not in sp23:
            continue
        for i in nbrs[j]:
            if i == k:
                continue
            for l in nbrs[k]:
                if l == j or l == i:
                    continue
                p = FFH.GetMMFFTorsionParams(mol, i, j, k, l)  # (torsionType, V1, V2, V3)
                if p:
                    t["torsion"][0].append((i, j, k, l))
                    t["torsion"][1].append((p[1], p[2], p[3]))

    topo = Chem.GetDistanceMatrix(mol)
    frags = None
    if ignore_interfrag:
        frags = np.zeros(n, dtype=np.int64)
        for f, atoms in enumerate(Chem.GetMolFrags(mol)):
            frags[list(atoms)] = f
    within = np.zeros((n, n), dtype=bool)
    for xyz in coords:
        d = np.sqrt(((xyz[:, None, :] - xyz[None, :, :]) ** 2).sum(-1))
        within |= d <= non_bonded_thresh
    charges = [props.GetMMFFPartialCharge(i) for i in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            if topo[i, j] < 3 or not within[i, j] or (frags is not None and frags[i] != frags[j]):
                continue
            pv = FFH.GetMMFFVdWParams(mol, i, j)  # (R_ij_starUnscaled, epsilonUnscaled, R_ij_star, epsilon)
            if pv:
                t["vdw"][0].append((i, j))
                t["vdw"][1].append((pv[2], pv[3]))
            qq = charges[i] * charges[j]  # constant dielectric, D = 1 (RDKit defaults; MMFFMolProperties has no getters)
            if abs(qq) > 1.0e-10:
                t["ele"][0].append((i, j))
                t["ele"][1].append((qq, 1.0, 1.0 if topo[i, j] == 3 else 0.0))
    return {k: (np.array(v[0], dtype=np.int16).reshape(len(v[0]), -1), np.array(v[1], dtype=np.float64).reshape(len(v[1]), -1))
            for k, v in t.items()}


class _FlatWithMap:
    """FlatSystem + ConformerBatch + the (molecule, conformer id) each batch entry came from."""

    def __init__(self, system: FlatSystem, batch: ConformerBatch, conf_ids: List[List[int]]):
        self.system, self.batch, self.conf_ids = system, batch, conf_ids


def mmff_from_rdkit(molecules, properties=None, nonBondedThreshold: float = 100.0, ignoreInterfragInteractions: bool = True):
    """Flatten RDKit molecules (all conformers) into MMFF term tables. Error contract of the reference
    (nvmolkit/mmffOptimization.py:145-162): ValueError(message, {"none": [...], "no_params": [...]})."""
    Chem, FFH = _rdkit()
    molecules = list(molecules)
    none = [i for i, m in enumerate(molecules) if m is None]
    props_list = []
    if properties is not None and not isinstance(properties, (list, tuple)):
        properties = [properties] * len(molecules)
    no_params = []
    for i, m in enumerate(molecules):
        if m is None:
            props_list.append(None)
            continue
        p = properties[i] if properties is not None and properties[i] is not None else FFH.MMFFGetMoleculeProperties(m)
        if p is None:
            no_params.append(i)
        props_list.append(p)
    if none or no_params:
        raise ValueError("MMFF cannot be set up for some molecules (None entries or missing MMFF parameters)",
                         {"none": none, "no_params": no_params})
    per_mol, counts, coords_per_mol, conf_ids = [], [], [], []
    for m, p in zip(molecules, props_list):
        coords = _conformer_coords(m)
        per_mol.append(_mmff_terms(Chem, FFH, m, p, coords, float(nonBondedThreshold), bool(ignoreInterfragInteractions)))
        counts.append(m.GetNumAtoms())
        coords_per_mol.append(coords)
        conf_ids.append([c.GetId() for c in m.GetConformers()])
    system = FlatSystem.from_molecules("mmff", counts, per_mol)
    return _FlatWithMap(system, ConformerBatch.from_coords(system, coords_per_mol), conf_ids)


def write_back_conformers(molecules, coords_per_mol) -> None:
    """Overwrite the conformers of each RDKit molecule, in conformer order, with the optimised coordinates."""
    from rdkit.Geometry import Point3D

    for mol, confs in zip(molecules, coords_per_mol):
        for conf, xyz in zip(mol.GetConformers(), confs):
            for a, (x, y, z) in enumerate(np.asarray(xyz, dtype=np.float64)):
                conf.SetAtomPosition(a, Point3D(float(x), float(y), float(z)))


def add_conformers(molecules, coords_per_mol, prune_rms_thresh: float = -1.0) -> None:
    """Append embedded conformers to the RDKit molecules (reference: addConformersToMoleculeWithPruning,
    src/etkdg.cpp:430-484). With prune_rms_thresh > 0 a conformer is kept only if its heavy-atom RMSD (after alignment)
    to every conformer kept before it exceeds the threshold."""
    from rdkit import Chem
    from rdkit.Chem import rdMolAlign
    from rdkit.Geometry import Point3D

    for mol, confs in zip(molecules, coords_per_mol):
        heavy = [a.GetIdx() for a in mol.GetAtoms() if a.GetAtomicNum() > 1]
        amap = list(zip(heavy, heavy))
        for xyz in confs:
            conf = Chem.Conformer(mol.GetNumAtoms())
            for a, (x, y, z) in enumerate(np.asarray(xyz, dtype=np.float64)):
                conf.SetAtomPosition(a, Point3D(float(x), float(y), float(z)))
            cid = mol.AddConformer(conf, assignId=True)
            if prune_rms_thresh > 0.0:
                for other in [c.GetId() for c in mol.GetConformers() if c.GetId() != cid]:
                    if rdMolAlign.GetBestRMS(mol, mol, other, cid, map=[amap]) < prune_rms_thresh:
                        mol.RemoveConformer(cid)
                        break


def uff_from_rdkit(molecules, vdwThreshold: float = 10.0, ignoreInterfragInteractions: bool = True):
    raise NotImplementedError(
        "UFF term construction from RDKit molecules needs RDKit's UFF builder internals (angle orders, inversion "
        "coefficients; rdkit_extensions/uff_flattened_builder.cpp) that the Python API does not expose. Pass a "
        "pre-flattened nvmolkit_b200.uffOptimization.FlatUFFMolecules (layout: include/b200mol.h).")


def embed_molecules_from_rdkit(molecules, params):
    raise NotImplementedError(
        "ETKDG set-up from RDKit molecules (bounds matrix, experimental torsions, chiral sets: "
        "src/embedder_utils.cpp:671-708) is not reachable through RDKit's Python API. Pass a pre-flattened "
        "nvmolkit_b200.embedMolecules.FlatEmbedMolecules (layout: include/b200mol.h).")
