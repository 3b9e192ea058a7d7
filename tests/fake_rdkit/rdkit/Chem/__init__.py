import enum

import numpy as np


class HybridizationType(enum.IntEnum):
    UNSPECIFIED = 0
    S = 1
    SP = 2
    SP2 = 3
    SP3 = 4
    SP3D = 6
    SP3D2 = 7


class ChiralType(enum.IntEnum):
    CHI_UNSPECIFIED = 0
    CHI_TETRAHEDRAL_CW = 1
    CHI_TETRAHEDRAL_CCW = 2
    CHI_OTHER = 3


class BondType(enum.IntEnum):
    SINGLE = 1
    DOUBLE = 2
    TRIPLE = 3
    AROMATIC = 12


class BondStereo(enum.IntEnum):
    STEREONONE = 0
    STEREOANY = 1
    STEREOZ = 2
    STEREOE = 3
    STEREOCIS = 4
    STEREOTRANS = 5


class Conformer:
    def __init__(self, n=0):
        self._xyz = np.zeros((n, 3))
        self._id = 0

    def GetPositions(self):
        return self._xyz

    def SetAtomPosition(self, a, p):
        self._xyz[a] = (p.x, p.y, p.z)

    def GetId(self):
        return self._id


class Atom:
    def __init__(self, mol, idx):
        self._m, self._i = mol, idx

    def GetIdx(self):
        return self._i

    def GetAtomicNum(self):
        return int(self._m.d["z"][self._i])

    def GetDegree(self):
        return len(self._m.d["nbrs"][self._i])

    def GetNeighbors(self):
        return [Atom(self._m, j) for j in self._m.d["nbrs"][self._i]]

    def GetBonds(self):
        return [b for b in self._m.GetBonds() if self._i in (b.GetBeginAtomIdx(), b.GetEndAtomIdx())]

    def GetHybridization(self):
        return self._m.hyb[self._i]

    # what the Morgan invariant generator asks for (src/morgan_fingerprint_common.cpp:43-124); hydrogens are explicit
    # atoms in the synthetic molecules, charges / isotopes come from optional per-atom arrays of the dict
    def GetNumExplicitHs(self):
        return 0

    def GetNumImplicitHs(self):
        return int(self._m.d.get("implicit_hs", [0] * self._m.GetNumAtoms())[self._i])

    def GetFormalCharge(self):
        return int(self._m.d.get("formal_charge", [0] * self._m.GetNumAtoms())[self._i])

    def GetMass(self):
        return GetPeriodicTable().GetAtomicWeight(self.GetAtomicNum()) + float(self._m.d.get("isotope_delta", [0] * self._m.GetNumAtoms())[self._i])

    def GetChiralTag(self):
        return self._m.chiral.get(self._i, ChiralType.CHI_UNSPECIFIED)


class Bond:
    def __init__(self, mol, k):
        self._m, self._k = mol, k

    def GetBeginAtomIdx(self):
        return int(self._m.d["bonds"][self._k][0])

    def GetEndAtomIdx(self):
        return int(self._m.d["bonds"][self._k][1])

    def GetOtherAtomIdx(self, a):
        i, j = self._m.d["bonds"][self._k]
        return int(j if a == i else i)

    def GetBondType(self):
        return self._m.bond_types[self._k]

    def GetBondTypeAsDouble(self):
        return {BondType.SINGLE: 1.0, BondType.DOUBLE: 2.0, BondType.TRIPLE: 3.0, BondType.AROMATIC: 1.5}[self.GetBondType()]

    def GetStereo(self):
        return self._m.bond_stereo.get(self._k, (BondStereo.STEREONONE, ()))[0]

    def GetStereoAtoms(self):
        return self._m.bond_stereo.get(self._k, (BondStereo.STEREONONE, ()))[1]


class RingInfo:
    def __init__(self, rings, n):
        self._r = rings
        self._n = n

    def NumAtomRings(self, i):
        return sum(1 for r in self._r if i in r)

    def IsAtomInRingOfSize(self, i, size):
        return any(i in r and len(r) == size for r in self._r)

    def AtomRingSizes(self, i):
        return [len(r) for r in self._r if i in r]


class Mol:
    """Wraps a synthetic molecule dict; `planar` atoms are sp2, every other heavy atom sp3."""

    def __init__(self, d, conformers=(), chiral=None, rings=(), bond_types=None, bond_stereo=None, hyb=None):
        self.d = d
        n = len(d["z"])
        planar = d.get("planar", set())
        self.hyb = hyb or [HybridizationType.SP2 if a in planar else (HybridizationType.S if d["z"][a] == 1 else HybridizationType.SP3)
                           for a in range(n)]
        self.chiral = chiral or {}
        self.rings = [tuple(r) for r in rings]
        self.bond_types = bond_types or [BondType.SINGLE] * len(d["bonds"])
        self.bond_stereo = bond_stereo or {}
        self._confs = []
        for xyz in conformers:
            c = Conformer(n)
            c._xyz = np.array(xyz, dtype=np.float64)
            self.AddConformer(c, assignId=True)

    def GetNumAtoms(self):
        return len(self.d["z"])

    def GetAtoms(self):
        return [Atom(self, i) for i in range(self.GetNumAtoms())]

    def GetAtomWithIdx(self, i):
        return Atom(self, i)

    def GetBonds(self):
        return [Bond(self, k) for k in range(len(self.d["bonds"]))]

    def GetBondBetweenAtoms(self, i, j):
        for k, (a, b) in enumerate(self.d["bonds"]):
            if (a, b) == (i, j) or (a, b) == (j, i):
                return Bond(self, k)
        return None

    def GetRingInfo(self):
        return RingInfo(self.rings, self.GetNumAtoms())

    def GetConformers(self):
        return list(self._confs)

    def AddConformer(self, conf, assignId=False):
        if assignId:
            conf._id = max([c._id for c in self._confs], default=-1) + 1
        self._confs.append(conf)
        return conf._id

    def RemoveConformer(self, cid):
        self._confs = [c for c in self._confs if c._id != cid]

    def GetSubstructMatches(self, query):  # the default torsion-bond SMARTS: bonds whose two atoms both have degree > 1
        return [(i, j) for i, j in self.d["bonds"] if len(self.d["nbrs"][i]) > 1 and len(self.d["nbrs"][j]) > 1]


class _PeriodicTable:
    _W = {1: 1.008, 6: 12.011, 7: 14.007, 8: 15.999, 9: 18.998, 15: 30.974, 16: 32.06, 17: 35.45, 35: 79.904}

    def GetAtomicWeight(self, z):
        return self._W.get(int(z), 2.0 * int(z))


def GetPeriodicTable():
    return _PeriodicTable()


def MolFromSmarts(s):
    return s


def GetDistanceMatrix(mol):
    return np.asarray(mol.d["topo"], dtype=np.float64)


def GetMolFrags(mol):
    return (tuple(range(mol.GetNumAtoms())),)


def AssignStereochemistry(mol):
    return None
