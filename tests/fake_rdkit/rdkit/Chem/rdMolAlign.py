import numpy as np


def GetBestRMS(prb, ref, prbId=-1, refId=-1, map=None):
    a = next(c for c in prb.GetConformers() if c.GetId() == prbId).GetPositions()
    b = next(c for c in ref.GetConformers() if c.GetId() == refId).GetPositions()
    idx = [i for i, _ in map[0]] if map else list(range(len(a)))
    a, b = a[idx] - a[idx].mean(0), b[idx] - b[idx].mean(0)
    u, s, vt = np.linalg.svd(a.T @ b)
    d = np.sign(np.linalg.det(u @ vt))
    e0 = (a * a).sum() + (b * b).sum()
    return float(np.sqrt(max(e0 - 2.0 * (s[0] + s[1] + d * s[2]), 0.0) / len(a)))
