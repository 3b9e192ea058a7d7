"""Term construction from bounds matrices: product builders (libb200mol, host C++) against the
independent C restatement in oracle/oracle_build.c and hand-computed known answers. No GPU needed."""
import numpy as np
import pytest

import oracle
from nvmolkit_b200 import synthetic as S
from nvmolkit_b200.builders import CrystalFFDetails, dg_terms_from_bounds, etk_terms_from_details, flat_embed_molecules


def _canon(idx, par):
    idx = np.asarray(idx, np.int64)
    order = np.lexsort(tuple(idx[:, k] for k in range(idx.shape[1] - 1, -1, -1)))
    return idx[order], np.asarray(par)[order]


def _details_of(mol, rng):
    """CrystalFF-shaped details of a synthetic pseudo molecule: a torsion per rotatable heavy-atom bond, impropers on the
    planar centres (with one phosphorus-like and one C=O-like centre thrown in), all bonds, all angles (a few triple)."""
    nbrs, z = mol["nbrs"], mol["z"]
    tors, v, sg = [], [], []
    for j, k in mol["bonds"]:
        if len(nbrs[j]) < 2 or len(nbrs[k]) < 2:
            continue
        i = [x for x in nbrs[j] if x != k][0]
        l = [x for x in nbrs[k] if x != j][0]
        if len({i, j, k, l}) < 4:
            continue
        tors.append([i, j, k, l])
        m = int(rng.integers(1, 7))
        v.append(np.concatenate([rng.uniform(0, 5, m), np.zeros(6 - m)]))
        sg.append(np.concatenate([rng.choice([-1, 1], m), np.zeros(6 - m)]))
    imps = []
    for c in sorted(mol.get("planar", set())):
        a, b, d = nbrs[c]
        zc = int(rng.choice([int(z[c]), 15, 7]))
        imps.append([a, c, b, d, zc, int(rng.random() < 0.3)])
    angles = []
    for c in range(len(z)):
        nb = nbrs[c]
        for x in range(len(nb)):
            for y in range(x + 1, len(nb)):
                angles.append([nb[x], c, nb[y], int(rng.random() < 0.05)])
    return CrystalFFDetails(np.array(tors).reshape(-1, 4), np.array(v).reshape(-1, 6), np.array(sg).reshape(-1, 6),
                            np.array(imps).reshape(-1, 6), np.array(mol["bonds"]).reshape(-1, 2), np.array(angles).reshape(-1, 4),
                            float(rng.uniform(0.5, 2.0)))


def test_inversion_coefficients_known_answers():
    # sp2 carbon: C0 = 1, C1 = -1, C2 = 0, k = 6 / 3 (50 / 3 when bound to O) - dist_geom_flattened_builder.cpp:186-191
    assert oracle.inversion_coefficients(6, False) == (2.0, 1.0, -1.0, 0.0)
    assert oracle.inversion_coefficients(7, True) == (50.0 / 3.0, 1.0, -1.0, 0.0)
    # phosphorus: w = 84.4339 deg, C2 = 1, C1 = -4 cos w, C0 = -(C1 cos w + cos 2w), k = 22 / (C0 + C1 + C2) / 3
    w = np.deg2rad(84.4339)
    c1 = -4 * np.cos(w)
    c0 = -(c1 * np.cos(w) + np.cos(2 * w))
    k, a, b, c = oracle.inversion_coefficients(15, False)
    assert np.allclose([k, a, b, c], [22.0 / (c0 + c1 + 1.0) / 3.0, c0, c1, 1.0], rtol=1e-14)


def test_four_atom_chain_by_hand():
    # chain 0-1-2-3, bounds: bonds 1.5 +- 0.01, 1-3 pairs 2.4..2.6, 1-4 pair 2.5..3.9
    b = np.zeros((4, 4))
    ub = {(0, 1): 1.51, (1, 2): 1.51, (2, 3): 1.51, (0, 2): 2.6, (1, 3): 2.6, (0, 3): 3.9}
    lb = {(0, 1): 1.49, (1, 2): 1.49, (2, 3): 1.49, (0, 2): 2.4, (1, 3): 2.4, (0, 3): 2.5}
    for (i, j), u in ub.items():
        b[i, j], b[j, i] = u, lb[(i, j)]
    dg = dg_terms_from_bounds(b)
    idx, par = _canon(*dg["dist"])
    assert idx.tolist() == [[1, 0], [2, 0], [2, 1], [3, 0], [3, 1], [3, 2]]
    assert np.allclose(par[3], [2.5 ** 2, 3.9 ** 2, 1.0], rtol=0, atol=0)
    assert dg["fourth"][0].ravel().tolist() == [0, 1, 2, 3] and len(dg["chiral"][0]) == 0
    det = CrystalFFDetails([[0, 1, 2, 3]], [[0, 0, 4.0]], [[1, -1, 1]], np.zeros((0, 6)), [[0, 1], [1, 2], [2, 3]],
                           [[0, 1, 2, 0], [1, 2, 3, 1]], 1.0)
    etk, n_imp = etk_terms_from_details(b, det, True)
    assert n_imp == 0 and len(etk["improper"][0]) == 0
    assert etk["torsion"][1][0].tolist() == [0, 0, 4.0, 0, 0, 0, 1, -1, 1, 0, 0, 0]
    assert np.allclose(etk["dist12"][1], [[1.49, 1.51, 100.0, 0.0]] * 3, atol=1e-15)
    assert etk["dist13"][0].tolist() == [[0, 2]] and np.allclose(etk["dist13"][1], [[2.49, 2.51, 100.0, 0.0]], atol=1e-15)
    assert etk["angle13"][0].tolist() == [[1, 2, 3]] and etk["angle13"][1].tolist() == [[179.0, 180.0]]  # the triple bond
    assert len(etk["longrange"][0]) == 0  # 0-3 is the torsion's 1-4 pair, everything else is 1-2 / 1-3
    etk2, _ = etk_terms_from_details(b, det, False)  # without basic knowledge the triple-bond angle is a plain 1-3 window
    assert len(etk2["angle13"][0]) == 0 and etk2["dist13"][0].tolist() == [[0, 2], [1, 3]]


@pytest.mark.parametrize("basic", [True, False])
def test_builders_equal_the_oracle(basic):
    rng = np.random.default_rng(11)
    flat, mols = S.random_embed_molecules(6, 5, 18, seed=12)
    for m, mol in enumerate(mols):
        b = mol["bounds"]
        dg = dg_terms_from_bounds(b)
        i_o, p_o = oracle.dg_dist_terms(b)
        i_g, p_g = _canon(*dg["dist"])
        i_o, p_o = _canon(i_o, p_o)
        assert np.array_equal(i_g, i_o) and np.array_equal(p_g, p_o)
        # the same numbers the synthetic generator wrote directly (it lists pairs as i < j)
        st, ix, pr = flat.dg.tables["dist"]
        i_s, p_s = _canon(np.sort(ix[st[m]:st[m + 1]].astype(np.int64), axis=1)[:, ::-1], pr[st[m]:st[m + 1]])
        assert np.array_equal(i_s, i_g) and np.array_equal(p_s, p_g)
        det = _details_of(mol, rng)
        etk, n_imp = etk_terms_from_details(b, det, basic)
        want, n_imp_o = oracle.etk_terms(b, det.torsion_atoms, det.improper_atoms, det.bonds, det.angles,
                                         det.bounds_mat_force_scaling, basic)
        assert n_imp == n_imp_o == (len(det.improper_atoms) if basic else 0)
        for name in ("improper", "dist12", "dist13", "angle13", "longrange"):
            gi, gp = _canon(*etk[name])
            oi, op = _canon(*want[name])
            assert np.array_equal(gi, oi), name
            assert np.array_equal(gp, op), name
        assert np.array_equal(etk["torsion"][0], det.torsion_atoms)
        assert np.array_equal(etk["torsion"][1], np.concatenate([det.torsion_v, det.torsion_signs], axis=1))
        # every pair is covered exactly once by 1-2 | 1-3 | angle | torsion 1-4 | long-range
        n = len(mol["z"])
        pairs = set()
        for name in ("dist12", "dist13", "longrange"):
            pairs |= {tuple(sorted(p)) for p in etk[name][0].tolist()}
        pairs |= {tuple(sorted((a[0], a[2]))) for a in etk["angle13"][0].tolist()}
        pairs |= {tuple(sorted((t[0], t[3]))) for t in det.torsion_atoms.tolist()}
        assert len(pairs) == n * (n - 1) // 2


def test_flat_embed_molecules_builds_schedulable_tables_and_rejects_bad_input():
    rng = np.random.default_rng(5)
    _, mols = S.random_embed_molecules(3, 5, 10, seed=13)
    dets = [_details_of(m, rng) for m in mols]
    chir = [(np.array([[0, 1, 2, 3]]), np.array([[5.0, 100.0]])) if len(m["z"]) > 4 else (None, None) for m in mols]
    flat = flat_embed_molecules([m["bounds"] for m in mols], dets, chir)
    assert len(flat) == 3 and flat.dg.tables["chiral"][2][0].tolist() == [100.0, 5.0]  # table order {upper, lower}
    assert flat.checks.num_impropers.tolist() == [len(d.improper_atoms) for d in dets]
    with pytest.raises(ValueError):
        dg_terms_from_bounds(np.zeros((3, 4)))
    bad = CrystalFFDetails(bonds=[[0, 99]])
    with pytest.raises(ValueError):
        etk_terms_from_details(mols[0]["bounds"], bad)
