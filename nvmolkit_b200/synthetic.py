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
