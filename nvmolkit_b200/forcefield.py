"""Flattened force-field systems — the seam data format of the conformer path (include/b200mol.h).

A ``FlatSystem`` is a MOLECULE table: per term type a CSR ``starts[nMols+1]``, molecule-local int16 atom indices
``idx[n][K]`` and fp64 parameter records ``par[n][P]``. Conformer batches point into it, so all conformers of a
molecule share one term block (the reference re-flattens the terms per conformer, src/forcefields/mmff.cu
addMoleculeToBatch). ``to_device()`` uploads it once and yields the ctypes struct the C-ABI takes.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

# kind -> ordered (table name, K indices, P parameters); order = field order of the C structs in include/b200mol.h
LAYOUT: Dict[str, Tuple[Tuple[str, int, int], ...]] = {
    "mmff": (("bond", 2, 2), ("angle", 3, 3), ("strbend", 3, 5), ("oop", 4, 1), ("torsion", 4, 3), ("vdw", 2, 2),
             ("ele", 2, 3), ("distc", 2, 3), ("posc", 1, 5), ("anglec", 3, 3), ("torsc", 4, 3)),
    "uff": (("bond", 2, 2), ("angle", 3, 6), ("torsion", 4, 3), ("inversion", 4, 4), ("vdw", 2, 3), ("distc", 2, 3),
            ("posc", 1, 5), ("anglec", 3, 3), ("torsc", 4, 3)),
    "dg": (("dist", 2, 3), ("chiral", 4, 2), ("fourth", 1, 0)),
    "etk": (("torsion", 4, 12), ("improper", 4, 4), ("dist12", 2, 4), ("dist13", 2, 4), ("angle13", 3, 2),
            ("longrange", 2, 3)),
}
DIM = {"mmff": 3, "uff": 3, "dg": 4, "etk": 4}
CHECK_LAYOUT = (("tetrahedral", 5, 1), ("chiral", 5, 2), ("chiralDist", 2, 2), ("dbStereo", 4, 1), ("dbGeom", 3, 0))


class TermTableC(C.Structure):
    _fields_ = [("starts", C.c_void_p), ("idx", C.c_void_p), ("par", C.c_void_p), ("molWaves", C.c_void_p),
                ("waves", C.c_void_p)]


def _system_struct(kind: str):
    fields = [("nMols", C.c_int32), ("atomCounts", C.c_void_p)] + [(name, TermTableC) for name, _, _ in LAYOUT[kind]]
    return type(f"{kind.capitalize()}SystemC", (C.Structure,), {"_fields_": fields})


SYSTEM_STRUCT = {k: _system_struct(k) for k in LAYOUT}


def schedule_waves(starts: np.ndarray, idx: np.ndarray, par: np.ndarray):
    """Order every molecule's terms into atom-disjoint waves of <= 32 (the gradient schedule of the kernels: one warp
    per wave, lane = term, plain shared-memory adds instead of atomics; include/b200mol.h b200mol_schedule_waves).
    Returns (idx, par, mol_waves [nMols+1], waves [nWaves+1]). Any order of the terms is correct; energies change
    only by summation order."""
    from nvmolkit_b200 import _lib

    starts = np.ascontiguousarray(starts, dtype=np.int32)
    idx = np.ascontiguousarray(idx, dtype=np.int16)
    n, k = idx.shape
    perm = np.empty(n, dtype=np.int32)
    mol_waves = np.empty(len(starts), dtype=np.int32)
    waves = np.empty(n + 1, dtype=np.int32)
    n_waves = C.c_int64(0)
    _lib.call("b200mol_schedule_waves", len(starts) - 1, starts.ctypes.data, idx.ctypes.data if n else None, int(k),
              perm.ctypes.data, mol_waves.ctypes.data, waves.ctypes.data, C.byref(n_waves))
    return (np.ascontiguousarray(idx[perm]), np.ascontiguousarray(par[perm]), mol_waves,
            np.ascontiguousarray(waves[: n_waves.value + 1]))


@dataclass
class FlatSystem:
    kind: str
    atom_counts: np.ndarray  # int32 [nMols]
    tables: Dict[str, Tuple[np.ndarray, np.ndarray, np.ndarray]]  # name -> (starts, idx [n,K], par [n,P])
    _device: dict = field(default_factory=dict, repr=False)
    waves: Dict[str, Tuple[np.ndarray, np.ndarray]] = field(default_factory=dict, repr=False)  # name -> (molWaves, waves)

    def __post_init__(self):
        self.atom_counts = np.ascontiguousarray(self.atom_counts, dtype=np.int32)
        if self.waves:  # already scheduled (concat of scheduled systems)
            return
        n_mols = len(self.atom_counts)
        fixed = {}
        for name, k, p in LAYOUT[self.kind]:
            if name not in self.tables:  # (restraint tables are optional: none)
                self.tables[name] = (np.zeros(n_mols + 1, np.int32), np.zeros((0, k), np.int16), np.zeros((0, p)))
            starts, idx, par = self.tables[name]
            starts = np.ascontiguousarray(starts, dtype=np.int32)
            idx = np.ascontiguousarray(idx, dtype=np.int16).reshape(-1, k)
            par = np.ascontiguousarray(par, dtype=np.float64).reshape(-1, p) if p else np.zeros((len(idx), 0))
            if len(starts) != n_mols + 1 or starts[-1] != len(idx) or (p and len(par) != len(idx)):
                raise ValueError(f"inconsistent term table '{name}'")
            idx, par, mol_waves, waves = schedule_waves(starts, idx, par)
            self.waves[name] = (mol_waves, waves)
            fixed[name] = (starts, idx, par)
        self.tables = fixed

    @property
    def n_mols(self) -> int:
        return len(self.atom_counts)

    @classmethod
    def from_molecules(cls, kind: str, atom_counts: Sequence[int], mols: Sequence[Dict[str, Tuple]]) -> "FlatSystem":
        """mols[m][table] = (idx [n,K], par [n,P]) with molecule-local indices."""
        tables = {}
        for name, k, p in LAYOUT[kind]:
            idxs, pars, starts = [], [], [0]
            for m in mols:
                idx, par = m.get(name, (np.zeros((0, k), np.int16), np.zeros((0, p))))
                idx = np.asarray(idx, dtype=np.int16).reshape(-1, k)
                idxs.append(idx)
                pars.append(np.asarray(par, dtype=np.float64).reshape(len(idx), p))
                starts.append(starts[-1] + len(idx))
            tables[name] = (np.array(starts, dtype=np.int32),
                            np.concatenate(idxs) if idxs else np.zeros((0, k), np.int16),
                            np.concatenate(pars) if pars else np.zeros((0, p)))
        return cls(kind, np.asarray(atom_counts, dtype=np.int32), tables)

    @classmethod
    def concat(cls, parts: "Sequence[FlatSystem]") -> "FlatSystem":
        """The molecules of several systems of one kind, in order; the wave schedules are spliced, not recomputed."""
        kind = parts[0].kind
        tables, waves = {}, {}
        for name, _k, _p in LAYOUT[kind]:
            starts, mol_waves, wave_list = [np.zeros(1, np.int32)], [np.zeros(1, np.int32)], []
            t_off = w_off = 0
            for part in parts:
                st, _ix, _pr = part.tables[name]
                mw, wv = part.waves[name]
                starts.append(st[1:] + t_off)
                mol_waves.append(mw[1:] + w_off)
                wave_list.append(wv[:-1] + t_off)
                t_off += int(st[-1])
                w_off += len(wv) - 1
            wave_list.append(np.array([t_off], np.int32))
            tables[name] = (np.concatenate(starts).astype(np.int32), np.concatenate([p.tables[name][1] for p in parts]),
                            np.concatenate([p.tables[name][2] for p in parts]))
            waves[name] = (np.concatenate(mol_waves).astype(np.int32), np.concatenate(wave_list).astype(np.int32))
        return cls(kind, np.concatenate([p.atom_counts for p in parts]), tables, waves=waves)

    def nbytes(self) -> int:
        """Bytes of the tables as uploaded (starts, indices, parameters, wave schedule)."""
        total = self.atom_counts.nbytes
        for name, (st, ix, pr) in self.tables.items():
            total += st.nbytes + ix.nbytes + pr.nbytes + sum(a.nbytes for a in self.waves[name])
        return int(total)

    def tile(self, reps: int) -> "FlatSystem":
        """The same molecules repeated `reps` times as distinct table entries (bench workloads)."""
        tables = {}
        for name, (starts, idx, par) in self.tables.items():
            counts = np.tile(np.diff(starts), reps)
            tables[name] = (np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), np.tile(idx, (reps, 1)),
                            np.tile(par, (reps, 1)))
        return FlatSystem(self.kind, np.tile(self.atom_counts, reps), tables)

    def host_struct(self):
        """ctypes struct with HOST pointers (what the CPU oracle consumes)."""
        st = SYSTEM_STRUCT[self.kind]()
        st.nMols = self.n_mols
        st.atomCounts = self.atom_counts.ctypes.data
        for name, _, p in LAYOUT[self.kind]:
            starts, idx, par = self.tables[name]
            mw, wv = self.waves[name]
            setattr(st, name, TermTableC(starts.ctypes.data, idx.ctypes.data, par.ctypes.data if p else None,
                                         mw.ctypes.data, wv.ctypes.data))
        return st

    def to_device(self, device=None):
        """Upload once per device; returns (ctypes struct of device pointers, keep-alive list)."""
        import torch

        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        key = str(dev)
        if key not in self._device:
            keep = []

            def up(a):
                t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                keep.append(t)
                return t.data_ptr()

            st = SYSTEM_STRUCT[self.kind]()
            st.nMols = self.n_mols
            st.atomCounts = up(self.atom_counts)
            for name, _, p in LAYOUT[self.kind]:
                starts, idx, par = self.tables[name]
                mw, wv = self.waves[name]
                setattr(st, name, TermTableC(up(starts), up(idx) if len(idx) else None, up(par) if p and len(par) else None,
                                             up(mw), up(wv)))
            self._device[key] = (st, keep)
        return self._device[key]


@dataclass
class ConformerBatch:
    """Conformers of the molecules of a FlatSystem: conformer c is molecule conf_mol[c]; coordinates are
    positions[atom_starts[c]:atom_starts[c+1]] (fp64, `dim` columns)."""

    conf_mol: np.ndarray  # int32 [nConf]
    atom_starts: np.ndarray  # int32 [nConf+1]
    positions: np.ndarray  # float64 [totalAtoms, dim]

    def __post_init__(self):
        self.conf_mol = np.ascontiguousarray(self.conf_mol, dtype=np.int32)
        self.atom_starts = np.ascontiguousarray(self.atom_starts, dtype=np.int32)
        self.positions = np.ascontiguousarray(self.positions, dtype=np.float64)

    @property
    def n_conf(self) -> int:
        return len(self.conf_mol)

    @property
    def max_atoms(self) -> int:
        return int(np.diff(self.atom_starts).max(initial=0))

    @classmethod
    def from_coords(cls, system: FlatSystem, coords_per_mol: Sequence[Sequence[np.ndarray]]) -> "ConformerBatch":
        conf_mol, starts, pos = [], [0], []
        for m, confs in enumerate(coords_per_mol):
            for xyz in confs:
                xyz = np.asarray(xyz, dtype=np.float64)
                if xyz.shape[0] != system.atom_counts[m]:
                    raise ValueError(f"conformer of molecule {m} has {xyz.shape[0]} atoms, expected {system.atom_counts[m]}")
                conf_mol.append(m)
                starts.append(starts[-1] + xyz.shape[0])
                pos.append(xyz)
        dim = pos[0].shape[1] if pos else 3
        return cls(np.array(conf_mol, dtype=np.int32), np.array(starts, dtype=np.int32),
                   np.concatenate(pos) if pos else np.zeros((0, dim)))


class ChecksC(C.Structure):
    _fields_ = [(name, TermTableC) for name, _, _ in CHECK_LAYOUT] + [("numImpropers", C.c_void_p)]


@dataclass
class CheckTables:
    """ETKDG stereo / geometry check tables (include/b200mol.h b200mol_etkdg_checks), CSR by molecule."""

    tables: Dict[str, Tuple[np.ndarray, np.ndarray, np.ndarray]]
    num_impropers: np.ndarray  # int32 [nMols]
    _device: dict = field(default_factory=dict, repr=False)

    @classmethod
    def from_molecules(cls, atom_counts, mols: Sequence[Dict[str, Tuple]], num_impropers) -> "CheckTables":
        tables = {}
        for name, k, p in CHECK_LAYOUT:
            idxs, pars, starts = [], [], [0]
            for m in mols:
                idx, par = m.get(name, (np.zeros((0, k), np.int16), np.zeros((0, p))))
                idx = np.asarray(idx, dtype=np.int16).reshape(-1, k)
                idxs.append(idx)
                pars.append(np.asarray(par, dtype=np.float64).reshape(len(idx), p))
                starts.append(starts[-1] + len(idx))
            tables[name] = (np.array(starts, dtype=np.int32), np.ascontiguousarray(np.concatenate(idxs)),
                            np.ascontiguousarray(np.concatenate(pars)))
        return cls(tables, np.ascontiguousarray(num_impropers, dtype=np.int32))

    @classmethod
    def concat(cls, parts: "Sequence[CheckTables]") -> "CheckTables":
        tables = {}
        for name, _k, _p in CHECK_LAYOUT:
            starts, off = [np.zeros(1, np.int32)], 0
            for part in parts:
                st = part.tables[name][0]
                starts.append(st[1:] + off)
                off += int(st[-1])
            tables[name] = (np.concatenate(starts).astype(np.int32), np.concatenate([p.tables[name][1] for p in parts]),
                            np.concatenate([p.tables[name][2] for p in parts]))
        return cls(tables, np.concatenate([p.num_impropers for p in parts]))

    def nbytes(self) -> int:
        return int(self.num_impropers.nbytes + sum(a.nbytes for t in self.tables.values() for a in t))

    def host_struct(self):
        st = ChecksC()
        for name, _, p in CHECK_LAYOUT:
            starts, idx, par = self.tables[name]
            setattr(st, name, TermTableC(starts.ctypes.data, idx.ctypes.data, par.ctypes.data if p else None, None, None))
        st.numImpropers = self.num_impropers.ctypes.data
        return st

    def to_device(self, device=None):
        import torch

        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        key = str(dev)
        if key not in self._device:
            keep = []

            def up(a):
                t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                keep.append(t)
                return t.data_ptr()

            st = ChecksC()
            for name, _, p in CHECK_LAYOUT:
                starts, idx, par = self.tables[name]
                setattr(st, name, TermTableC(up(starts), up(idx) if len(idx) else None, up(par) if p and len(par) else None,
                                             None, None))
            st.numImpropers = up(self.num_impropers)
            self._device[key] = (st, keep)
        return self._device[key]
