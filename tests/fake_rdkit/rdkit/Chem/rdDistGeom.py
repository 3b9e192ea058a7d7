import numpy as np


def GetMoleculeBoundsMatrix(mol, set15bounds=True, scaleVDW=False, doTriangleSmoothing=True, useMacrocycle14config=False):
    if doTriangleSmoothing:
        raise NotImplementedError("the adapter smooths on the GPU")
    b = np.array(mol.d["bounds_raw"], dtype=np.float64)
    if not set15bounds and mol.d.get("bounds_relaxed") is not None:
        b = np.array(mol.d["bounds_relaxed"], dtype=np.float64)
    return b


def GetExperimentalTorsions(mol, useExpTorsionAnglePrefs=True, useSmallRingTorsions=False, useMacrocycleTorsions=True,
                            useBasicKnowledge=True, ETversion=2, printExpTorsionAngles=False):
    return tuple(mol.d.get("exp_torsions", ()))
