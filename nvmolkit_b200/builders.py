"""Term construction from smoothed bounds matrices — the seam just below RDKit (reference:
``rdkit_extensions/dist_geom_flattened_builder.cpp:472-541`` ``constructForceFieldContribs`` /
``construct3DForceFieldContribs``), through the host-side C-ABI builders ``b200mol_dg_terms_from_bounds`` and
``b200mol_etk_terms_from_details`` (``csrc/builders.cu``). No RDKit needed: inputs are plain arrays.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from nvmolkit_b200 import _lib
from nvmolkit_b200.forcefield import CheckTables, FlatSystem


@dataclass
class CrystalFFDetails:
    """RDKit ``ForceFields::CrystalFF::CrystalFFDetails`` as arrays (what ``getExperimentalTorsions`` +
    ``setTopolBounds`` fill, ``src/embedder_utils.cpp:235-287``)."""

    torsion_atoms: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))  # expTorsionAtoms
    torsion_v: np.ndarray = field(default_factory=lambda: np.zeros((0, 6)))  # force constants
    torsion_signs: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.int32))
    improper_atoms: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.int32))  # a0, centre, a2, a3, Z, isCBoundToO
    bonds: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    angles: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))  # a, centre, b, isTripleBond
    bounds_mat_force_scaling: float = 1.0

    def __post_init__(self):
        self.torsion_atoms = np.ascontiguousarray(self.torsion_atoms, np.int32).reshape(-1, 4)
        v = np.zeros((len(self.torsion_atoms), 6))
        sg = np.zeros((len(self.torsion_atoms), 6), np.int32)
        if len(self.torsion_atoms):  # zero-padded to six terms like the reference (:160-174)
            tv = np.asarray(self.torsion_v, np.float64).reshape(len(self.torsion_atoms), -1)[:, :6]
            ts = np.asarray(self.torsion_signs, np.int32).reshape(len(self.torsion_atoms), -1)[:, :6]
            v[:, : tv.shape[1]] = tv
            sg[:, : ts.shape[1]] = ts
        self.torsion_v, self.torsion_signs = v, sg
        self.improper_atoms = np.ascontiguousarray(self.improper_atoms, np.int32).reshape(-1, 6)
        self.bonds = np.ascontiguousarray(self.bonds, np.int32).reshape(-1, 2)
        self.angles = np.ascontiguousarray(self.angles, np.int32).reshape(-1, 4)


class _DetailsC(C.Structure):
    _fields_ = [("nTorsions", C.c_int32), ("torsionAtoms", C.c_void_p), ("torsionV", C.c_void_p), ("torsionSigns", C.c_void_p),
                ("nImpropers", C.c_int32), ("improperAtoms", C.c_void_p), ("nBonds", C.c_int32), ("bonds", C.c_void_p),
                ("nAngles", C.c_int32), ("angles", C.c_void_p), ("boundsMatForceScaling", C.c_double)]


class _EtkBuffersC(C.Structure):
    _fields_ = [(f"{t}_{k}", C.c_void_p) for t in ("torsion", "improper", "dist12", "dist13", "angle13", "longrange")
                for k in ("idx", "par")]


def inversion_coefficients(z: int, c_bound_to_o: bool):
    """(k / 3, C0, C1, C2) of an inversion centre of atomic number z - the formula of csrc/builders.cu
    (dist_geom_flattened_builder.cpp:178-235; the UFF inversion shares it, uff_flattened_builder.cpp:454-530)."""
    if z in (6, 7, 8):
        return (50.0 if c_bound_to_o else 6.0) / 3.0, 1.0, -1.0, 0.0
    w = np.pi / 180.0 * {15: 84.4339, 33: 86.9735, 51: 87.7047, 83: 90.0}.get(z, 1.0)
    c2 = 1.0
    c1 = -4.0 * np.cos(w)
    c0 = -(c1 * np.cos(w) + c2 * np.cos(2.0 * w))
    return 22.0 / (c0 + c1 + c2) / 3.0, c0, c1, c2


def dg_terms_from_bounds(bounds: np.ndarray, chiral_atoms=None, chiral_bounds=None, dim: int = 4,
                         basin_size_tol: float = 1e8) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """DG tables {dist, chiral, fourth} of one molecule from its (smoothed) bounds matrix and chiral sets
    (``chiral_atoms [n,4]``, ``chiral_bounds [n,2]`` = lower, upper volume)."""
    b = np.ascontiguousarray(bounds, np.float64)
    n = b.shape[0]
    if b.shape != (n, n):
        raise ValueError(f"bounds matrix must be square, got {b.shape}")
    ca = np.ascontiguousarray(chiral_atoms if chiral_atoms is not None else np.zeros((0, 4)), np.int32).reshape(-1, 4)
    cb = np.ascontiguousarray(chiral_bounds if chiral_bounds is not None else np.zeros((0, 2)), np.float64).reshape(-1, 2)
    if len(ca) != len(cb):
        raise ValueError("chiral_atoms and chiral_bounds differ in length")
    npair = n * (n - 1) // 2
    d_idx, d_par = np.empty((npair, 2), np.int16), np.empty((npair, 3))
    c_idx, c_par = np.empty((len(ca), 4), np.int16), np.empty((len(ca), 2))
    f_idx = np.empty((n, 1), np.int16)
    counts = (C.c_int32 * 3)()
    _lib.call("b200mol_dg_terms_from_bounds", n, b.ctypes.data, len(ca), ca.ctypes.data, cb.ctypes.data, int(dim),
              float(basin_size_tol), d_idx.ctypes.data, d_par.ctypes.data, c_idx.ctypes.data, c_par.ctypes.data,
              f_idx.ctypes.data, counts)
    return {"dist": (d_idx[: counts[0]], d_par[: counts[0]]), "chiral": (c_idx[: counts[1]], c_par[: counts[1]]),
            "fourth": (f_idx[: counts[2]], np.zeros((counts[2], 0)))}


def etk_terms_from_details(bounds: np.ndarray, details: CrystalFFDetails, use_basic_knowledge: bool = True):
    """ETK tables of one molecule ({torsion, improper, dist12, dist13, angle13, longrange}) and the planarity count."""
    b = np.ascontiguousarray(bounds, np.float64)
    n = b.shape[0]
    d = details
    st = _DetailsC(len(d.torsion_atoms), d.torsion_atoms.ctypes.data, d.torsion_v.ctypes.data, d.torsion_signs.ctypes.data,
                   len(d.improper_atoms), d.improper_atoms.ctypes.data, len(d.bonds), d.bonds.ctypes.data, len(d.angles),
                   d.angles.ctypes.data, float(d.bounds_mat_force_scaling))
    sizes = {"torsion": (len(d.torsion_atoms), 4, 12), "improper": (3 * len(d.improper_atoms), 4, 4),
             "dist12": (len(d.bonds), 2, 4), "dist13": (len(d.angles), 2, 4), "angle13": (len(d.angles), 3, 2),
             "longrange": (n * (n - 1) // 2, 2, 3)}
    arrays = {t: (np.empty((m, k), np.int16), np.empty((m, p))) for t, (m, k, p) in sizes.items()}
    buf = _EtkBuffersC(*[a.ctypes.data for t in sizes for a in arrays[t]])
    counts = (C.c_int32 * 6)()
    n_imp = C.c_int32(0)
    _lib.call("b200mol_etk_terms_from_details", n, b.ctypes.data, C.byref(st), 1 if use_basic_knowledge else 0,
              C.byref(buf), counts, C.byref(n_imp))
    return {t: (arrays[t][0][: counts[i]], arrays[t][1][: counts[i]]) for i, t in enumerate(sizes)}, int(n_imp.value)


def flat_embed_molecules(bounds_list: Sequence[np.ndarray], details_list: Sequence[CrystalFFDetails],
                         chiral_list: Optional[Sequence[Tuple[np.ndarray, np.ndarray]]] = None,
                         checks_list: Optional[Sequence[dict]] = None, use_basic_knowledge: bool = True):
    """Everything the embedding kernel needs for a list of molecules, from their smoothed bounds matrices, CrystalFF
    details, chiral sets ((atoms [n,4], bounds [n,2]) per molecule) and check tables (dicts keyed like
    ``forcefield.CHECK_LAYOUT``; default: none) — the replacement of the per-molecule part of prepareEmbedderArgs that
    follows RDKit's own calls (``src/embedder_utils.cpp:671-708``)."""
    from nvmolkit_b200.embedMolecules import FlatEmbedMolecules

    dgs, etks, nimp = [], [], []
    for m, b in enumerate(bounds_list):
        ca, cb = chiral_list[m] if chiral_list is not None else (None, None)
        dgs.append(dg_terms_from_bounds(b, ca, cb))
        etk, k = etk_terms_from_details(b, details_list[m], use_basic_knowledge)
        etks.append(etk)
        nimp.append(k)
    counts = [np.asarray(b).shape[0] for b in bounds_list]
    checks = checks_list if checks_list is not None else [{} for _ in counts]
    return FlatEmbedMolecules(FlatSystem.from_molecules("dg", counts, dgs), FlatSystem.from_molecules("etk", counts, etks),
                              CheckTables.from_molecules(counts, checks, nimp))


@dataclass
class StereoInfo:
    """What the reference's ``findChiralSets`` / ``findDoubleBonds`` extract from a molecule
    (``src/embedder_utils.cpp:117-206, 617-664``), as plain arrays: the structural input of the acceptance checks."""

    chiral_centers: np.ndarray = field(default_factory=lambda: np.zeros((0, 5), np.int32))  # centre, n1..n4 (n4 = centre if 3-coord.)
    chiral_bounds: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))  # volume lower, upper
    tetrahedral: np.ndarray = field(default_factory=lambda: np.zeros((0, 5), np.int32))  # centre, n1..n4
    tetrahedral_fused: np.ndarray = field(default_factory=lambda: np.zeros(0))  # 1 = in two or more small (< 5) rings
    double_bond_ends: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int32))  # neighbour, atom, other end
    stereo_double_bonds: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))  # stereo atom, begin, end, stereo atom
    stereo_signs: np.ndarray = field(default_factory=lambda: np.zeros(0))  # -1 cis / Z, +1 trans / E

    def __post_init__(self):
        self.chiral_centers = np.ascontiguousarray(self.chiral_centers, np.int32).reshape(-1, 5)
        self.chiral_bounds = np.ascontiguousarray(self.chiral_bounds, np.float64).reshape(-1, 2)
        self.tetrahedral = np.ascontiguousarray(self.tetrahedral, np.int32).reshape(-1, 5)
        self.tetrahedral_fused = np.ascontiguousarray(self.tetrahedral_fused, np.float64).reshape(-1)
        self.double_bond_ends = np.ascontiguousarray(self.double_bond_ends, np.int32).reshape(-1, 3)
        self.stereo_double_bonds = np.ascontiguousarray(self.stereo_double_bonds, np.int32).reshape(-1, 4)
        self.stereo_signs = np.ascontiguousarray(self.stereo_signs, np.float64).reshape(-1)

    def dg_chiral_terms(self) -> Tuple[np.ndarray, np.ndarray]:
        """(atoms [n,4], bounds [n,2] lower/upper) of the DG chiral-volume terms: the four neighbours of every chiral set
        (RDKit ChiralSet d_idx1..4; dist_geom_flattened_builder.cpp:88-109)."""
        return self.chiral_centers[:, 1:5], self.chiral_bounds


def check_tables(bounds: np.ndarray, stereo: StereoInfo) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """Acceptance-check tables of one molecule (layout ``forcefield.CHECK_LAYOUT`` / b200mol_etkdg_checks):
    tetrahedral and chiral sets as they are; the chiral distance-matrix check over every pair of atoms that belong to a
    four-coordinate chiral set, with that pair's bounds (ETKDGChiralDistMatrixCheckStage::loadDataset,
    src/etkdg_stage_stereochem_checks.cu:614-655); double-bond geometry triples and stereo quadruples."""
    b = np.asarray(bounds, np.float64)
    atoms = sorted({int(a) for row in stereo.chiral_centers if row[0] != row[4] for a in row})
    pairs = [(atoms[j], atoms[k]) for j in range(len(atoms)) for k in range(j + 1, len(atoms))]
    cd_par = [[b[max(i, j), min(i, j)], b[min(i, j), max(i, j)]] for i, j in pairs]
    return {"tetrahedral": (stereo.tetrahedral.astype(np.int16), stereo.tetrahedral_fused.reshape(-1, 1)),
            "chiral": (stereo.chiral_centers.astype(np.int16), stereo.chiral_bounds),
            "chiralDist": (np.array(pairs, np.int16).reshape(-1, 2), np.array(cd_par, np.float64).reshape(-1, 2)),
            "dbStereo": (stereo.stereo_double_bonds.astype(np.int16), stereo.stereo_signs.reshape(-1, 1)),
            "dbGeom": (stereo.double_bond_ends.astype(np.int16), np.zeros((len(stereo.double_bond_ends), 0)))}


def flat_embed_from_parts(bounds_list, details_list, stereo_list, use_basic_knowledge: bool = True):
    """FlatEmbedMolecules from per-molecule (smoothed bounds, CrystalFFDetails, StereoInfo): what prepareEmbedderArgs
    leaves behind after RDKit's own calls (``src/embedder_utils.cpp:671-708``), flattened for the embedding kernel."""
    chiral = [s.dg_chiral_terms() for s in stereo_list]
    checks = [check_tables(b, s) for b, s in zip(bounds_list, stereo_list)]
    return flat_embed_molecules(bounds_list, details_list, chiral, checks, use_basic_knowledge)
