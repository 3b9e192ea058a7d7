"""Host-side logic that needs no GPU: term ordering, the Butina round rule (emulated), oracle threading."""
import numpy as np

import oracle
from nvmolkit_b200 import synthetic as S
from nvmolkit_b200.forcefield import FlatSystem, _diagonal_order


def test_diagonal_order_spreads_atoms_and_keeps_energies():
    system, xyz, _ = S.random_mmff_system(3, 20, 30, seed=5)
    starts, idx, par = system.tables["vdw"]
    for m in range(system.n_mols):
        pairs = idx[starts[m]:starts[m + 1]].astype(int)
        d = np.abs(pairs[:, 1] - pairs[:, 0])
        assert (np.diff(d) >= 0).all()  # sorted by |j - i| ...
        for a in range(0, len(pairs) - 32, 32):  # ... so a warp's worth of terms touches (almost) only distinct atoms
            blk = pairs[a:a + 32]
            if d[a] == d[a + 31] and d[a] >= 32:
                assert len(set(blk.ravel().tolist())) == 64
    # any order gives the same energies up to summation order: shuffle the pair tables and compare through the oracle
    rng = np.random.default_rng(0)
    tables = {}
    for name, (st, ix, pr) in system.tables.items():
        if ix.shape[1] == 2:
            perm = np.concatenate([st[m] + rng.permutation(st[m + 1] - st[m]) for m in range(system.n_mols)]).astype(int)
            ix, pr = ix[perm], pr[perm]
        tables[name] = (st, ix, pr)
    for m, x in enumerate(xyz):
        e0 = oracle.ff_energy_grad("mmff", system.atom_counts, system.tables, m, x)[0]
        e1 = oracle.ff_energy_grad("mmff", system.atom_counts, tables, m, x)[0]
        assert abs(e0 - e1) <= 1e-11 * max(1.0, abs(e0))
    # idempotent
    i2, p2 = _diagonal_order(starts, idx, par)
    assert (i2 == idx).all() and (p2 == par).all()
    assert isinstance(system, FlatSystem)


def _butina_rounds(adj, n):
    """Pure-Python statement of butinaRoundsKernel (csrc/butina.cu): commit every 2-hop local maximum of the key per
    round; ids = rank of the keys the centres had when chosen; leftovers = singletons in descending index order."""
    ids = -np.ones(n, dtype=np.int64)
    counts = np.array([len(a) for a in adj], dtype=np.int64)
    sel = {}
    key = lambda i: (int(counts[i]), i) if ids[i] < 0 else (-1, -1)  # noqa: E731
    rounds = 0
    while True:
        best1 = [max([key(m)] + [key(r) for r in adj[m]]) if ids[m] < 0 else (-1, -1) for m in range(n)]
        new = [p for p in range(n) if ids[p] < 0 and counts[p] > 0 and max([best1[m] for m in adj[p]] + [(-1, -1)]) <= key(p)]
        if not new:
            break
        for p in new:
            sel[p] = key(p)
        for p in new:
            ids[p] = p
            for m in adj[p]:
                if ids[m] >= 0:
                    continue
                ids[m] = p
                for i in adj[m]:
                    if i != p:
                        counts[i] -= 1
        rounds += 1
    centres = sorted(sel, key=lambda p: sel[p], reverse=True)
    id_of = {p: r for r, p in enumerate(centres)}
    out = np.array([id_of[c] if c >= 0 else -1 for c in ids])
    free = sorted([i for i in range(n) if out[i] < 0], reverse=True)
    for k, i in enumerate(free):
        out[i] = len(centres) + k
    return out, np.array(centres + free), rounds


def test_parallel_round_rule_reproduces_the_sequential_greedy_order():
    for n, degree, seed in ((40, 3, 1), (150, 6, 2), (300, 25, 3)):
        rng = np.random.default_rng(seed)
        d = rng.random((n, n))
        d = np.minimum(d, d.T)
        np.fill_diagonal(d, 0.0)
        cutoff = 1.0 - (1.0 - degree / n) ** 0.5
        adj = [np.nonzero((d[i] <= cutoff) & (np.arange(n) != i))[0].tolist() for i in range(n)]
        ids, cen, rounds = _butina_rounds(adj, n)
        ids_cpu, cen_cpu = oracle.butina_dense(d, cutoff)
        assert (ids == ids_cpu).all() and (cen == cen_cpu).all()
        assert rounds < len(set(ids_cpu.tolist()))  # many clusters per round, not one


def test_oracle_thread_control():
    n0 = oracle.set_threads(0)
    assert oracle.set_threads(2) == 2
    assert oracle.set_threads(n0) == n0
