"""Seeded synthetic workloads of the BASELINE.json shapes (no network, no RDKit on the box): fingerprints,
molecular graphs. Pure NumPy; used by tests/ and bench.py to feed both the CUDA path and the CPU oracle."""

from __future__ import annotations

import numpy as np

from nvmolkit_b200.molgraph import MolGraphBatch, atom_invariant

SEED = 20260924


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """bool [n][nbits] -> uint32 [n][nbits/32], bit j -> word j>>5, mask 1<<(j&31)."""
    n, nb = bits.shape
    b = bits.reshape(n, nb // 32, 32).astype(np.uint64)
    return (b << np.arange(32, dtype=np.uint64)).sum(axis=2).astype(np.uint32)


def random_fingerprints(n: int, bits: int = 2048, p: float = 0.025, seed: int = SEED, near_dups: int = 0) -> np.ndarray:
    """ECFP-like density (p=0.025 -> ~51 on bits of 2048); the last `near_dups` rows are copies of earlier rows with
    0-8 flipped bits so that similarities span [0, 1]."""
    rng = np.random.default_rng(seed)
    m = rng.random((n, bits)) < p
    for k in range(min(near_dups, n // 2)):
        src = rng.integers(0, n - near_dups)
        row = m[src].copy()
        flips = rng.integers(0, 9)
        row[rng.integers(0, bits, size=flips)] ^= True
        m[n - 1 - k] = row
    return pack_bits(m)


def clustered_fingerprints(n_centres: int, members: int, bits: int = 2048, p: float = 0.025, max_flips: int = 12,
                           seed: int = SEED, chunk: int = 4096) -> np.ndarray:
    """BASELINE config 2 generator: n_centres random centres x `members` copies with 0..max_flips random bit flips,
    shuffled (style of nvmolkit/tests/test_clustering.py:154-163)."""
    rng = np.random.default_rng(seed)
    words = bits // 32
    out = np.empty((n_centres * members, words), dtype=np.uint32)
    for c0 in range(0, n_centres, chunk):
        c1 = min(n_centres, c0 + chunk)
        centres = pack_bits(rng.random((c1 - c0, bits)) < p)
        block = np.repeat(centres, members, axis=0)
        nrows = block.shape[0]
        nflips = rng.integers(0, max_flips + 1, size=nrows)
        for f in range(max_flips):
            active = nflips > f
            pos = rng.integers(0, bits, size=nrows)
            w, b = pos >> 5, (pos & 31).astype(np.uint32)
            rows = np.nonzero(active)[0]
            block[rows, w[rows]] ^= (np.uint32(1) << b[rows])
        out[c0 * members:c1 * members] = block
    rng.shuffle(out, axis=0)
    return out


_ELEMENTS = np.array([6, 6, 6, 6, 7, 8, 16, 9], dtype=np.int64)
_MAXVAL = {6: 4, 7: 3, 8: 2, 16: 2, 9: 1}


def random_molgraphs(n: int, min_atoms: int = 8, max_atoms: int = 50, seed: int = SEED) -> MolGraphBatch:
    """Drug-like pseudo molecules as heavy-atom graphs: a random tree with valence-respecting degrees plus 0-3 ring
    closures, bond types 1/2/12 (RDKit SINGLE/DOUBLE/AROMATIC codes), atom invariants hashed exactly like RDKit's
    Morgan atom invariants from (Z, degree+Hs, Hs, charge=0, deltaMass=0, inRing)."""
    rng = np.random.default_rng(seed)
    a_starts, b_starts = [0], [0]
    ainv, binv, ba, bb = [], [], [], []
    for _ in range(n):
        na = int(rng.integers(min_atoms, max_atoms + 1))
        z = _ELEMENTS[rng.integers(0, len(_ELEMENTS), size=na)]
        z[0] = 6
        deg = np.zeros(na, dtype=np.int64)
        bonds = []
        for a in range(1, na):
            cand = [q for q in range(a) if deg[q] < _MAXVAL[int(z[q])] - (1 if q else 0)] or \
                   [q for q in range(a) if deg[q] < 4]
            q = int(cand[rng.integers(0, len(cand))])
            if deg[q] >= _MAXVAL[int(z[q])]:
                z[q] = 6
            bonds.append((q, a))
            deg[q] += 1
            deg[a] += 1
        in_ring = np.zeros(na, dtype=bool)
        adj = {i: set() for i in range(na)}
        for u, v in bonds:
            adj[u].add(v)
            adj[v].add(u)
        for _r in range(int(rng.integers(0, 4))):
            u, v = (int(x) for x in rng.integers(0, na, size=2))
            if u == v or v in adj[u] or deg[u] >= _MAXVAL[int(z[u])] or deg[v] >= _MAXVAL[int(z[v])]:
                continue
            # mark the tree path u..v as ring atoms (BFS parents)
            prev, frontier = {u: -1}, [u]
            while frontier and v not in prev:
                nxt = []
                for x in frontier:
                    for y in adj[x]:
                        if y not in prev:
                            prev[y] = x
                            nxt.append(y)
                frontier = nxt
            x = v
            while x != -1:
                in_ring[x] = True
                x = prev[x]
            bonds.append((u, v))
            adj[u].add(v)
            adj[v].add(u)
            deg[u] += 1
            deg[v] += 1
        types = np.where(rng.random(len(bonds)) < 0.2, 2, 1)
        for k, (u, v) in enumerate(bonds):
            if in_ring[u] and in_ring[v] and rng.random() < 0.5:
                types[k] = 12
        for a in range(na):
            hs = max(0, _MAXVAL[int(z[a])] - int(deg[a]))
            ainv.append(atom_invariant(int(z[a]), int(deg[a]) + hs, hs, 0, 0, bool(in_ring[a])))
        for k, (u, v) in enumerate(bonds):
            binv.append(int(types[k]))
            ba.append(u)
            bb.append(v)
        a_starts.append(len(ainv))
        b_starts.append(len(binv))
    return MolGraphBatch(np.array(a_starts), np.array(b_starts), np.array(ainv, dtype=np.uint32),
                         np.array(binv, dtype=np.uint32), np.array(ba, dtype=np.uint16), np.array(bb, dtype=np.uint16))


# ------------------------------------------------------------------------------------------------------------------
# Path B: pseudo drug-like molecules with MMFF-shaped term tables and 3-D coordinates (no RDKit on the box).
# ------------------------------------------------------------------------------------------------------------------
_VALENCE = {6: 4, 7: 3, 8: 2, 16: 2, 9: 1, 1: 1}
_VDW_R = {1: 2.6, 6: 3.9, 7: 3.7, 8: 3.5, 9: 3.3, 16: 4.2}
_VDW_E = {1: 0.02, 6: 0.07, 7: 0.08, 8: 0.09, 9: 0.06, 16: 0.12}


def _grow_molecule(rng, n_heavy):
    """Heavy-atom tree + hydrogens to fill valences; 3-D coordinates by greedy placement. Returns (z, bonds, xyz)."""
    z = list(_ELEMENTS[rng.integers(0, len(_ELEMENTS), size=n_heavy)])
    z[0] = 6
    deg = [0] * n_heavy
    bonds = []
    for a in range(1, n_heavy):
        cand = [q for q in range(a) if deg[q] < _VALENCE[int(z[q])] - (1 if q else 0)] or [q for q in range(a) if deg[q] < 3]
        q = int(cand[rng.integers(0, len(cand))])
        if deg[q] >= _VALENCE[int(z[q])]:
            z[q] = 6
        bonds.append((q, a))
        deg[q] += 1
        deg[a] += 1
    for a in range(n_heavy):  # hydrogens
        for _ in range(max(0, _VALENCE[int(z[a])] - deg[a])):
            z.append(1)
            bonds.append((a, len(z) - 1))
    n = len(z)
    nbrs = [[] for _ in range(n)]
    for u, v in bonds:
        nbrs[u].append(v)
        nbrs[v].append(u)
    xyz = np.zeros((n, 3))
    placed = np.zeros(n, dtype=bool)
    placed[0] = True
    order = [0]
    for a in order:
        for b in nbrs[a]:
            if placed[b]:
                continue
            r0 = 1.09 if (z[a] == 1 or z[b] == 1) else 1.5
            dirs = rng.normal(size=(24, 3))
            dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
            cand = xyz[a] + r0 * dirs
            others = xyz[placed]
            dmin = np.linalg.norm(cand[:, None, :] - others[None, :, :], axis=2)
            dmin[:, np.nonzero(np.nonzero(placed)[0] == a)[0]] = 9.0
            xyz[b] = cand[np.argmax(dmin.min(axis=1))]
            placed[b] = True
            order.append(b)
    return np.array(z, dtype=np.int64), bonds, nbrs, xyz


def _topological_distances(n, nbrs):
    dist = np.full((n, n), 99, dtype=np.int16)
    for s in range(n):
        dist[s, s] = 0
        frontier, d = [s], 0
        while frontier:
            d += 1
            nxt = []
            for x in frontier:
                for y in nbrs[x]:
                    if dist[s, y] == 99:
                        dist[s, y] = d
                        nxt.append(y)
            frontier = nxt
    return dist


def random_mmff_molecule(rng, n_heavy):
    """One pseudo molecule: dict of MMFF term tables (molecule-local indices) + coordinates + graph."""
    z, bonds, nbrs, xyz = _grow_molecule(rng, n_heavy)
    n = len(z)
    topo = _topological_distances(n, nbrs)
    t = {}
    b_idx = np.array(bonds, dtype=np.int16).reshape(-1, 2)
    is_h = (z[b_idx[:, 0]] == 1) | (z[b_idx[:, 1]] == 1)
    t["bond"] = (b_idx, np.stack([np.where(is_h, 1.09, 1.5) + rng.normal(0, 0.01, len(b_idx)),
                                  np.where(is_h, 4.8, 4.3) + rng.normal(0, 0.2, len(b_idx))], axis=1))
    ang, oop, tor = [], [], []
    for j in range(n):
        nb = nbrs[j]
        for x in range(len(nb)):
            for y in range(x + 1, len(nb)):
                ang.append((nb[x], j, nb[y]))
        if len(nb) == 3 and z[j] in (6, 7) and rng.random() < 0.3:
            a, b, c = nb
            oop += [(a, j, b, c), (a, j, c, b), (b, j, c, a)]
    for j, k in bonds:
        for i in nbrs[j]:
            if i == k:
                continue
            for l in nbrs[k]:
                if l == j or l == i:
                    continue
                tor.append((i, j, k, l))
    ang = np.array(ang, dtype=np.int16).reshape(-1, 3)
    theta0 = np.where(rng.random(len(ang)) < 0.25, 120.0, 109.5) + rng.normal(0, 1.0, len(ang))
    t["angle"] = (ang, np.stack([theta0, 0.5 + 0.4 * rng.random(len(ang)), np.zeros(len(ang))], axis=1))
    r0 = lambda a, b: np.where((z[a] == 1) | (z[b] == 1), 1.09, 1.5)  # noqa: E731
    t["strbend"] = (ang, np.stack([theta0, r0(ang[:, 0], ang[:, 1]), r0(ang[:, 2], ang[:, 1]),
                                   0.3 * rng.random(len(ang)), 0.3 * rng.random(len(ang))], axis=1))
    oop = np.array(oop, dtype=np.int16).reshape(-1, 4)
    t["oop"] = (oop, 0.02 + 0.05 * rng.random((len(oop), 1)))
    tor = np.array(tor, dtype=np.int16).reshape(-1, 4)
    t["torsion"] = (tor, rng.normal(0, 0.3, (len(tor), 3)))
    iu, ju = np.nonzero(np.triu(topo >= 3, k=1))
    pairs = np.stack([iu, ju], axis=1).astype(np.int16)
    rr = np.array([_VDW_R[int(e)] for e in z])
    ee = np.array([_VDW_E[int(e)] for e in z])
    t["vdw"] = (pairs, np.stack([0.5 * (rr[iu] + rr[ju]), np.sqrt(ee[iu] * ee[ju])], axis=1))
    q = rng.normal(0, 0.15, n)
    q -= q.mean()
    t["ele"] = (pairs, np.stack([q[iu] * q[ju], np.ones(len(iu)), (topo[iu, ju] == 3).astype(np.float64)], axis=1))
    return {"z": z, "bonds": bonds, "nbrs": nbrs, "topo": topo, "xyz": xyz, "terms": t}


def random_mmff_system(n_mols: int, min_heavy: int = 10, max_heavy: int = 50, seed: int = SEED):
    """(FlatSystem kind 'mmff', list of start coordinates [nAtoms,3], list of raw molecule dicts)."""
    from nvmolkit_b200.forcefield import FlatSystem

    rng = np.random.default_rng(seed)
    mols = [random_mmff_molecule(rng, int(rng.integers(min_heavy, max_heavy + 1))) for _ in range(n_mols)]
    system = FlatSystem.from_molecules("mmff", [len(m["z"]) for m in mols], [m["terms"] for m in mols])
    return system, [m["xyz"] for m in mols], mols


# ------------------------------------------------------------------------------------------------------------------
# ETKDG inputs for the pseudo molecules: smoothed bounds matrix -> DG terms, ETK terms, stereo/geometry check tables.
# ------------------------------------------------------------------------------------------------------------------
_VDW_RADIUS = {1: 1.1, 6: 1.7, 7: 1.55, 8: 1.52, 9: 1.47, 16: 1.8}


def smooth_bounds_numpy(ub: np.ndarray, lb: np.ndarray):
    """Floyd-style triangle smoothing on separate symmetric upper / lower matrices (data preparation only)."""
    n = len(ub)
    for k in range(n):
        ub = np.minimum(ub, ub[:, k:k + 1] + ub[k:k + 1, :])
        lb = np.maximum(lb, np.maximum(lb[:, k:k + 1] - ub[k:k + 1, :], lb[k:k + 1, :] - ub[:, k:k + 1]))
    np.fill_diagonal(ub, 0.0)
    np.fill_diagonal(lb, 0.0)
    return ub, lb


def bounds_matrix(mol) -> np.ndarray:
    """RDKit-layout bounds matrix ([i][j], i<j upper; [j][i] lower) from the molecule's topology, UNSMOOTHED."""
    z, nbrs, topo = mol["z"], mol["nbrs"], mol["topo"]
    n = len(z)
    planar = mol.get("planar", set())
    r0 = lambda a, b: 1.09 if (z[a] == 1 or z[b] == 1) else 1.5  # noqa: E731
    ub = np.full((n, n), 1000.0)
    lb = np.zeros((n, n))
    rv = np.array([_VDW_RADIUS[int(e)] for e in z])
    scale = np.where(topo == 4, 0.7, np.where(topo == 5, 0.85, 1.0))
    lb[:] = scale * (rv[:, None] + rv[None, :])
    for i in range(n):
        for j in nbrs[i]:
            ub[i, j], lb[i, j] = r0(i, j) + 0.01, r0(i, j) - 0.01
    for j in range(n):
        theta = np.deg2rad(120.0 if j in planar else 109.5)
        nb = nbrs[j]
        for x in range(len(nb)):
            for y in range(x + 1, len(nb)):
                a, b = nb[x], nb[y]
                d = np.sqrt(r0(a, j) ** 2 + r0(b, j) ** 2 - 2 * r0(a, j) * r0(b, j) * np.cos(theta))
                ub[a, b] = ub[b, a] = d + 0.04
                lb[a, b] = lb[b, a] = d - 0.04
    for (j, k) in mol["bonds"]:
        for i in nbrs[j]:
            if i == k:
                continue
            for l in nbrs[k]:
                if l == j or l == i or topo[i, l] != 3:
                    continue
                tj = np.deg2rad(120.0 if j in planar else 109.5)
                tk = np.deg2rad(120.0 if k in planar else 109.5)
                r1, r2, r3 = r0(i, j), r0(j, k), r0(k, l)
                # cis (torsion 0) and trans (180) 1-4 distances
                xi, yi = -r1 * np.cos(tj), r1 * np.sin(tj)
                xl, yl = r2 + r3 * np.cos(tk) * -1.0, r3 * np.sin(tk)
                cis = np.hypot(xl - xi, yl - yi)
                trans = np.hypot(xl - xi, yl + yi)
                lo, hi = min(cis, trans) - 0.06, max(cis, trans) + 0.06
                ub[i, l] = ub[l, i] = min(ub[i, l], hi) if ub[i, l] < 999 else hi
                lb[i, l] = lb[l, i] = lo
    m = np.zeros((n, n))
    iu = np.triu_indices(n, 1)
    m[iu] = ub[iu]
    m.T[iu] = lb[iu]
    return m


def etkdg_tables(rng, mol, smoothed: np.ndarray, strict_checks: bool = True):
    """DG / ETK / check term tables of one pseudo molecule from its SMOOTHED bounds matrix."""
    z, nbrs, topo = mol["z"], mol["nbrs"], mol["topo"]
    n = len(z)
    planar = mol.get("planar", set())
    iu, ju = np.triu_indices(n, 1)
    ubv, lbv = smoothed[iu, ju], smoothed[ju, iu]
    dg = {"dist": (np.stack([iu, ju], 1), np.stack([lbv ** 2, ubv ** 2, np.ones(len(iu))], 1)),
          "fourth": (np.arange(n).reshape(-1, 1), np.zeros((n, 0)))}
    quat = [a for a in range(n) if len(nbrs[a]) == 4 and z[a] == 6]
    rng.shuffle(quat)
    chiral_idx, chiral_par, chk_chiral = [], [], []
    for c in quat[:2]:
        nb = list(nbrs[c])
        if rng.random() < 0.5:
            lo, hi = 5.0, 100.0
        else:
            lo, hi = -100.0, -5.0
        chiral_idx.append(nb)
        chiral_par.append([hi, lo])  # DG table: {volUpper, volLower}
        chk_chiral.append(([c] + nb, [lo, hi]))
    dg["chiral"] = (np.array(chiral_idx, dtype=np.int16).reshape(-1, 4), np.array(chiral_par).reshape(-1, 2))
    tet = [([c] + list(nbrs[c]), [0.0]) for c in quat[2:3] if rng.random() < 0.3]  # RDKit: ring-fusion carbons only
    pairs_cd = []
    for (cen, par) in chk_chiral:  # RDKit checks the bounds among the chiral centre and its neighbours; the pseudo
        for nb_ in cen[1:]:        # molecules keep the bonded pairs only (their 1-3 windows are synthetic)
            a, b = sorted((cen[0], nb_))
            pairs_cd.append(([a, b], [smoothed[b, a], smoothed[a, b]]))
    checks = {
        "tetrahedral": (np.array([t[0] for t in tet], dtype=np.int16).reshape(-1, 5), np.array([t[1] for t in tet]).reshape(-1, 1)),
        "chiral": (np.array([c[0] for c in chk_chiral], dtype=np.int16).reshape(-1, 5),
                   np.array([c[1] for c in chk_chiral]).reshape(-1, 2)),
        "chiralDist": (np.array([p[0] for p in pairs_cd], dtype=np.int16).reshape(-1, 2),
                       np.array([p[1] for p in pairs_cd]).reshape(-1, 2)),
    }
    # ETK terms
    tor_i, tor_p = [], []
    db_stereo, db_geom = [], []
    for (j, k) in mol["bonds"]:
        if len(nbrs[j]) < 2 or len(nbrs[k]) < 2 or z[j] == 1 or z[k] == 1:
            continue
        i = [x for x in nbrs[j] if x != k][0]
        l = [x for x in nbrs[k] if x != j][0]
        v = np.zeros(6)
        v[2] = rng.uniform(1.0, 6.0)
        if rng.random() < 0.3:
            v[0] = rng.uniform(0.0, 3.0)
        sg = rng.choice([-1.0, 1.0], size=6)
        tor_i.append([i, j, k, l])
        tor_p.append(np.concatenate([v, sg]))
        if j in planar and k in planar and len(db_stereo) < 1:
            db_stereo.append(([i, j, k, l], [float(rng.choice([-1.0, 1.0]))]))
            db_geom.append([i, j, k])
            db_geom.append([j, k, l])
    imp_i, imp_p = [], []
    for c in sorted(planar):
        a, b, d = nbrs[c]
        for (p, q, r) in ((a, b, d), (a, d, b), (b, d, a)):
            imp_i.append([p, c, q, r])
            imp_p.append([1.0, -1.0, 0.0, 10.0 / 3.0])
    d12 = [(i, j) for i, j in mol["bonds"]]
    xyz12 = np.array([[smoothed[max(i, j), min(i, j)], smoothed[min(i, j), max(i, j)], 100.0, 0.0] for i, j in d12]).reshape(-1, 4)
    i13, j13 = np.nonzero(np.triu(topo == 2, 1))
    xyz13 = np.stack([smoothed[j13, i13], smoothed[i13, j13], np.full(len(i13), 100.0),
                      np.array([1.0 if (set(nbrs[a]) & set(nbrs[b]) & planar) else 0.0 for a, b in zip(i13, j13)])], 1)
    ang = [(a, c, b) for c in sorted(planar) for (a, b) in ((nbrs[c][0], nbrs[c][1]),)]
    ilr, jlr = np.nonzero(np.triu(topo > 3, 1))
    etk = {
        "torsion": (np.array(tor_i, dtype=np.int16).reshape(-1, 4), np.array(tor_p).reshape(-1, 12)),
        "improper": (np.array(imp_i, dtype=np.int16).reshape(-1, 4), np.array(imp_p).reshape(-1, 4)),
        "dist12": (np.array(d12, dtype=np.int16).reshape(-1, 2), xyz12),
        "dist13": (np.stack([i13, j13], 1), xyz13.reshape(-1, 4)),
        "angle13": (np.array(ang, dtype=np.int16).reshape(-1, 3), np.tile([115.0, 125.0], (len(ang), 1)).reshape(-1, 2)),
        "longrange": (np.stack([ilr, jlr], 1), np.stack([smoothed[jlr, ilr], smoothed[ilr, jlr], np.full(len(ilr), 10.0)], 1)),
    }
    if not strict_checks:  # bench workloads: the synthetic 1-2 / stereo windows are not chemically consistent, so these
        pairs_cd, db_stereo = [], []  # two checks would only burn attempts; tests keep them to exercise every kernel
        checks["chiralDist"] = (np.zeros((0, 2), np.int16), np.zeros((0, 2)))
    checks["dbStereo"] = (np.array([d[0] for d in db_stereo], dtype=np.int16).reshape(-1, 4),
                          np.array([d[1] for d in db_stereo]).reshape(-1, 1))
    checks["dbGeom"] = (np.array(db_geom, dtype=np.int16).reshape(-1, 3), np.zeros((len(db_geom), 0)))
    return dg, etk, checks, len(planar)


def random_embed_molecules(n_mols: int, min_heavy: int = 6, max_heavy: int = 25, seed: int = SEED,
                           strict_checks: bool = True):
    """Pseudo molecules with everything ETKDG needs. Returns (FlatEmbedMolecules, raw molecule dicts)."""
    from nvmolkit_b200.embedMolecules import FlatEmbedMolecules
    from nvmolkit_b200.forcefield import CheckTables, FlatSystem

    rng = np.random.default_rng(seed)
    mols, dgs, etks, chks, nimp = [], [], [], [], []
    for _ in range(n_mols):
        m = random_mmff_molecule(rng, int(rng.integers(min_heavy, max_heavy + 1)))
        m["planar"] = {a for a in range(len(m["z"])) if len(m["nbrs"][a]) == 3 and m["z"][a] in (6, 7) and rng.random() < 0.4}
        b = bounds_matrix(m)
        n = len(m["z"])
        iu = np.triu_indices(n, 1)
        ub = np.zeros((n, n))
        lb = np.zeros((n, n))
        ub[iu] = b[iu]
        ub.T[iu] = b[iu]
        lb[iu] = b.T[iu]
        lb.T[iu] = b.T[iu]
        ub, lb = smooth_bounds_numpy(ub, lb)
        sm = np.zeros((n, n))
        sm[iu] = ub[iu]
        sm.T[iu] = lb[iu]
        m["bounds_raw"], m["bounds"] = b, sm
        dg, etk, chk, np_ = etkdg_tables(rng, m, sm, strict_checks)
        mols.append(m)
        dgs.append(dg)
        etks.append(etk)
        chks.append(chk)
        nimp.append(np_)
    counts = [len(m["z"]) for m in mols]
    flat = FlatEmbedMolecules(FlatSystem.from_molecules("dg", counts, dgs), FlatSystem.from_molecules("etk", counts, etks),
                              CheckTables.from_molecules(counts, chks, nimp))
    return flat, mols


def random_uff_system(n_mols: int, min_heavy: int = 10, max_heavy: int = 50, seed: int = SEED):
    """UFF-shaped term tables for the same pseudo molecules (every angle order 0..4 and torsion order 2/3/6 occurs)."""
    from nvmolkit_b200.forcefield import FlatSystem

    rng = np.random.default_rng(seed)
    mols = [random_mmff_molecule(rng, int(rng.integers(min_heavy, max_heavy + 1))) for _ in range(n_mols)]
    tabs = []
    for m in mols:
        t = m["terms"]
        b_idx, b_par = t["bond"]
        ang = t["angle"][0]
        theta0 = np.deg2rad(t["angle"][1][:, 0])
        order = rng.choice([0, 0, 0, 1, 2, 3, 4], size=len(ang)).astype(np.float64)
        ka = rng.uniform(50, 150, len(ang))
        c2 = 1.0 / (4.0 * np.maximum(np.sin(theta0) ** 2, 1e-3))
        c1 = -4.0 * c2 * np.cos(theta0)
        c0 = c2 * (2.0 * np.cos(theta0) ** 2 + 1.0)
        tor = t["torsion"][0]
        t_order = rng.choice([2.0, 3.0, 3.0, 6.0], size=len(tor))
        inv = t["oop"][0]
        pairs = t["vdw"][0]
        tabs.append({
            "bond": (b_idx, np.stack([b_par[:, 0], rng.uniform(500, 900, len(b_idx))], 1)),
            "angle": (ang, np.stack([theta0, ka, order, c0, c1, c2], 1)),
            "torsion": (tor, np.stack([rng.uniform(0.2, 2.5, len(tor)), t_order, rng.choice([-1.0, 1.0], len(tor))], 1)),
            "inversion": (inv, np.tile([6.0 / 3.0, 1.0, -1.0, 0.0], (len(inv), 1))),
            "vdw": (pairs, np.stack([t["vdw"][1][:, 0], t["vdw"][1][:, 1], np.full(len(pairs), 10.0)], 1)),
        })
    system = FlatSystem.from_molecules("uff", [len(m["z"]) for m in mols], tabs)
    return system, [m["xyz"] for m in mols], mols
