"""MMFF / UFF parameter getters answered from the synthetic term tables attached to the fake molecule."""
import numpy as np


def _lookup(mol, kind, name, atoms):
    tab = mol.d[kind][name]
    key = tuple(int(a) for a in atoms)
    idx, par = tab
    for row, p in zip(np.asarray(idx).tolist(), np.asarray(par)):
        if tuple(row) == key or tuple(row) == key[::-1]:
            return p
    return None


class _Props:
    def __init__(self, mol):
        self._m = mol

    def GetMMFFAtomType(self, i):
        return 1

    def GetMMFFPartialCharge(self, i):
        return float(self._m.d["charges"][i])


def MMFFGetMoleculeProperties(mol):
    return None if mol.d.get("no_mmff") else _Props(mol)


def GetMMFFBondStretchParams(mol, i, j):
    p = _lookup(mol, "terms", "bond", (i, j))
    return None if p is None else (0, p[1], p[0])


def GetMMFFAngleBendParams(mol, i, j, k):
    p = _lookup(mol, "terms", "angle", (i, j, k))
    return None if p is None else (0, p[1], p[0])


def GetMMFFStretchBendParams(mol, i, j, k):
    idx, par = mol.d["terms"]["strbend"]
    for row, p in zip(np.asarray(idx).tolist(), np.asarray(par)):
        if tuple(row) == (i, j, k):
            return (0, p[3], p[4])
        if tuple(row) == (k, j, i):
            return (0, p[4], p[3])
    return None


def GetMMFFOopBendParams(mol, i, j, k, l):
    idx, par = mol.d["terms"]["oop"]
    for row, p in zip(np.asarray(idx).tolist(), np.asarray(par)):
        if row[1] == j and set(row) == {i, j, k, l}:
            return float(p[0])
    return None


def GetMMFFTorsionParams(mol, i, j, k, l):
    p = _lookup(mol, "terms", "torsion", (i, j, k, l))
    return None if p is None else (0, p[0], p[1], p[2])


def GetMMFFVdWParams(mol, i, j):
    p = _lookup(mol, "terms", "vdw", (i, j))
    return None if p is None else (p[0], p[1], p[0], p[1])


def UFFHasAllMoleculeParams(mol):
    return not mol.d.get("no_uff")


def GetUFFBondStretchParams(mol, i, j):
    p = _lookup(mol, "uff", "bond", (i, j))
    return None if p is None else (p[1], p[0])


def GetUFFAngleBendParams(mol, i, j, k):
    p = _lookup(mol, "uff", "angle", (i, j, k))
    return None if p is None else (p[1], float(np.rad2deg(p[0])))


def GetUFFTorsionParams(mol, i, j, k, l):
    p = _lookup(mol, "uff", "torsion", (i, j, k, l))
    return None if p is None else float(p[0])


def GetUFFInversionParams(mol, i, j, k, l):
    idx, par = mol.d["uff"]["inversion"]
    for row, p in zip(np.asarray(idx).tolist(), np.asarray(par)):
        if row[1] == j and set(row) == {i, j, k, l}:
            return float(p[0])
    return None


def GetUFFVdWParams(mol, i, j):
    p = _lookup(mol, "uff", "vdw", (i, j))
    return None if p is None else (p[0], p[1])
