"""Host-side logic that needs no GPU: term ordering, the Butina round rule (emulated), oracle threading."""
import numpy as np

import oracle
from nvmolkit_b200 import synthetic as S
from nvmolkit_b200.forcefield import LAYOUT, FlatSystem, schedule_waves


def _check_waves(system):
    """Every table: waves tile the molecule's term range, hold <= 32 terms, and never name an atom twice."""
    for name, k, _ in LAYOUT[system.kind]:
        starts, idx, _par = system.tables[name]
        mol_waves, waves = system.waves[name]
        assert len(mol_waves) == system.n_mols + 1 and waves[-1] == starts[-1]
        for m in range(system.n_mols):
            w0, w1 = mol_waves[m], mol_waves[m + 1]
            if starts[m] == starts[m + 1]:
                assert w0 == w1
                continue
            assert waves[w0] == starts[m] and waves[w1] == starts[m + 1]
            for w in range(w0, w1):
                t0, t1 = waves[w], waves[w + 1]
                assert 0 < t1 - t0 <= 32
                atoms = idx[t0:t1].ravel().tolist()
                assert len(set(atoms)) == len(atoms), (name, m, w)


def test_wave_schedule_is_conflict_free_and_keeps_energies():
    system, xyz, _ = S.random_mmff_system(3, 20, 30, seed=5)
    _check_waves(system)
    flat, _ = S.random_embed_molecules(3, 6, 14, seed=9)
    _check_waves(flat.dg)
    _check_waves(flat.etk)
    # an all-pairs table fills its waves: rounds of a round-robin tournament, (M - 1) / 2 pairs each
    starts, idx, _ = flat.dg.tables["dist"]
    mol_waves, waves = flat.dg.waves["dist"]
    for m in range(flat.dg.n_mols):
        a = int(flat.dg.atom_counts[m])
        n_waves = mol_waves[m + 1] - mol_waves[m]
        rounds = a if a % 2 else a + 1
        assert n_waves <= rounds * -(-(rounds // 2) // 32)
    # any order gives the same energies up to summation order: shuffle every table and compare through the oracle
    rng = np.random.default_rng(0)
    tables = {}
    for name, (st, ix, pr) in system.tables.items():
        perm = np.concatenate([st[m] + rng.permutation(st[m + 1] - st[m]) for m in range(system.n_mols)]).astype(int)
        tables[name] = (st, ix[perm], pr[perm])
    for m, x in enumerate(xyz):
        e0, g0, _ = oracle.ff_energy_grad("mmff", system.atom_counts, system.tables, m, x)
        e1, g1, _ = oracle.ff_energy_grad("mmff", system.atom_counts, tables, m, x)
        assert abs(e0 - e1) <= 1e-11 * max(1.0, abs(e0))
        assert np.abs(g0 - g1).max() <= 1e-10 * max(1.0, np.abs(g0).max())
    # deterministic
    st, ix, pr = system.tables["vdw"]
    i2, p2, mw2, w2 = schedule_waves(st, ix, pr)
    i3, p3, mw3, w3 = schedule_waves(st, ix, pr)
    assert (i2 == i3).all() and (p2 == p3).all() and (w2 == w3).all() and (mw2 == mw3).all()
    assert isinstance(system, FlatSystem)


def test_wave_schedule_edge_cases():
    # empty table, a single term, a molecule without terms between two with terms, a star graph (one hub atom)
    z = np.zeros((0, 2), np.int16)
    ix, pr, mw, wv = schedule_waves(np.array([0, 0, 0]), z, np.zeros((0, 3)))
    assert mw.tolist() == [0, 0, 0] and wv.tolist() == [0]
    star = np.array([[0, k] for k in range(1, 41)], np.int16)
    one = np.array([[3, 1]], np.int16)
    idx = np.concatenate([one, star])
    par = np.arange(len(idx), dtype=np.float64).reshape(-1, 1)
    ix, pr, mw, wv = schedule_waves(np.array([0, 1, 1, 41]), idx, par)
    assert mw.tolist() == [0, 1, 1, 41]  # the hub is in every star term: 40 waves of one term
    assert (np.diff(wv) == 1).all() and sorted(pr.ravel().tolist()) == list(range(41))
    assert ix[0].tolist() == [3, 1]
    # four-body terms
    rng = np.random.default_rng(3)
    quad = np.array([rng.permutation(12)[:4] for _ in range(200)], np.int16)
    ix, pr, mw, wv = schedule_waves(np.array([0, 200]), quad, np.zeros((200, 1)))
    for w in range(len(wv) - 1):
        atoms = ix[wv[w]:wv[w + 1]].ravel().tolist()
        assert len(set(atoms)) == len(atoms)


def test_clusters_from_ids_matches_the_reference_output_format():
    from nvmolkit_b200.clustering import clusters_from_ids

    ids = np.array([1, 0, 0, 2, 0, 1], np.int32)
    cen = np.array([4, 5, 3])
    clusters, sizes = clusters_from_ids(ids, cen)
    assert clusters == [(4, 1, 2), (5, 0), (3,)] and sizes == [0, 3, 5, 6]
    assert clusters_from_ids(np.zeros(0, np.int32), np.zeros(0, np.int32)) == ([], [0])


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
