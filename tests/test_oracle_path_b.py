"""CPU pins of the path-B oracle: reference known answers (BFGS analytic systems, eigenvalues), hand-computed terms,
finite-difference consistency of every force field, triangle-smoothing invariants."""

import numpy as np
import pytest

import oracle
from nvmolkit_b200 import synthetic as S
from nvmolkit_b200.forcefield import ConformerBatch, FlatSystem


def test_bfgs_quartic_reference_system():
    # tests/test_bfgs_minimizer.cu:822-930,1014-1030: E = sum (x_i - i)^4, 400 iterations, gradTol 1e-5, no scaling,
    # positions must end within 0.1 of the targets
    n = 4 * 7
    c = np.arange(n, dtype=np.float64)
    x0 = c + np.random.default_rng(1).uniform(-2, 2, n)
    x, e, status, iters = oracle.poly_minimize(4, np.ones(n), c, x0, 400, 1e-5, scale_grads=False)
    assert np.abs(x - c).max() < 0.1
    assert e < 1e-3


def test_bfgs_harmonic_converges_and_is_idempotent():
    # tests/test_bfgs_minimizer.cu:1159-1240: harmonic systems; a second call on a converged system changes nothing
    n = 30
    rng = np.random.default_rng(2)
    w, c = rng.uniform(0.5, 3.0, n), rng.normal(0, 3, n)
    x, e, status, iters = oracle.poly_minimize(2, w, c, np.zeros(n), 200, 1e-6)
    assert status == 0 and np.abs(x - c).max() < 1e-4
    x2, e2, status2, iters2 = oracle.poly_minimize(2, w, c, x, 200, 1e-6)
    assert status2 == 0 and iters2 <= 1 and np.abs(x2 - x).max() < 1e-6


def test_eigen_known_answers():
    # tests/test_coordgen.cu:98-135 (from RDKit's PowerEigenSolver tests), tolerance 1e-2
    m1 = np.array([0.0, 1.0, 1.732, 2.268, 3.268, 1.0, 0.0, 1.0, 1.732, 2.268, 1.732, 1.0, 0.0, 1.0, 1.732, 2.268, 1.732,
                   1.0, 0.0, 1.0, 3.268, 2.268, 1.732, 1.0, 0.0]).reshape(5, 5)
    m2 = np.ones((5, 5)) - np.eye(5)
    v0 = np.random.default_rng(0).random((5, 5))
    vals, vecs, k = oracle.power_eigen(m1, 5, v0)
    assert k == 5 and np.allclose(vals, [6.981, -3.982, -1.395, -1.016, -0.586], atol=1e-2)
    vals, vecs, k = oracle.power_eigen(m2, 5, v0)
    assert k == 5 and np.allclose(vals, [4.0, -1.0, -1.0, -1.0, -1.0], atol=1e-2)


def test_metric_embedding_recovers_distances():
    rng = np.random.default_rng(3)
    xyz = rng.normal(0, 2.0, (12, 3))
    d = np.linalg.norm(xyz[:, None] - xyz[None], axis=2)
    got = oracle.metric_embed(d, 3, rng.random((3, 12)))
    assert got is not None
    d2 = np.linalg.norm(got[:, None] - got[None], axis=2)
    assert np.abs(d2 - d).max() < 5e-2  # power iteration tolerance 1e-3 on eigenvalues


def test_triangle_smoothing_invariants():
    flat, mols = S.random_embed_molecules(3, 5, 10, seed=11)
    for m in mols:
        sm, ok = oracle.triangle_smooth(m["bounds_raw"])
        assert ok
        n = len(sm)
        ub = np.triu(sm, 1) + np.triu(sm, 1).T
        lb = np.tril(sm, -1) + np.tril(sm, -1).T
        for k in range(n):  # triangle inequalities hold after smoothing
            assert (ub <= ub[:, k:k + 1] + ub[k:k + 1, :] + 1e-9).all()
        iu = np.triu_indices(n, 1)
        assert (lb[iu] <= ub[iu] + 1e-12).all()
        assert np.allclose(sm, m["bounds"], atol=1e-9)  # generator's numpy smoothing agrees
        sm2, ok2 = oracle.triangle_smooth(sm)
        assert ok2 and np.array_equal(sm2, sm)  # idempotent
    bad = np.array([[0, 1.0, 1.0], [0.9, 0, 1.0], [5.0, 0.9, 0]])  # lower(0,2)=5 > ub(0,1)+ub(1,2)=2
    assert not oracle.triangle_smooth(bad)[1]


def test_mmff_single_terms_by_hand():
    # one bond, r0 = 1.5, kb = 4.0, stretched by 0.1: E = 143.9325/2 kb dr^2 (1 - 2 dr + 7/3 dr^2)
    sys1 = FlatSystem.from_molecules("mmff", [2], [{"bond": ([[0, 1]], [[1.5, 4.0]])}])
    e, g, per = oracle.ff_energy_grad("mmff", sys1.atom_counts, sys1.tables, 0, np.array([[0, 0, 0], [1.6, 0, 0.0]]))
    want = 143.9325 / 2 * 4.0 * 0.01 * (1 - 0.2 + 7.0 / 12.0 * 4 * 0.01)
    assert abs(e - want) < 1e-12 and abs(per[0] - want) < 1e-12
    de = 143.9325 * 4.0 * 0.1 * (1 - 3 * 0.1 + 2 * 7.0 / 12.0 * 4 * 0.01)
    assert np.allclose(g, [[-de, 0, 0], [de, 0, 0]], atol=1e-12)
    # buffered 14-7 at r = R*: E = eps (1.07/1.07)^7 (1.12/1.12 - 2) = -eps
    sys2 = FlatSystem.from_molecules("mmff", [2], [{"vdw": ([[0, 1]], [[3.5, 0.08]])}])
    e, g, per = oracle.ff_energy_grad("mmff", sys2.atom_counts, sys2.tables, 0, np.array([[0, 0, 0], [3.5, 0, 0.0]]))
    assert abs(e + 0.08) < 1e-14
    assert abs(g[1, 0] - 0.08 / 3.5 * ((-7.84 / 1.12 + 14.0) / 1.07 - 7.84 / 1.12 ** 2)) < 1e-12  # dE/dr at r = R*
    # electrostatics, constant dielectric, 1-4 scaled: 0.75 * 332.0716 q / (r + 0.05)
    sys3 = FlatSystem.from_molecules("mmff", [2], [{"ele": ([[0, 1]], [[0.1, 1.0, 1.0]])}])
    e, _, _ = oracle.ff_energy_grad("mmff", sys3.atom_counts, sys3.tables, 0, np.array([[0, 0, 0], [2.0, 0, 0.0]]))
    assert abs(e - 0.75 * 332.0716 * 0.1 / 2.05) < 1e-12


def _fd(fn, pos, h=1e-6):
    num = np.zeros_like(pos)
    for idx in np.ndindex(pos.shape):
        p = pos.copy()
        p[idx] += h
        ep = fn(p)
        p[idx] -= 2 * h
        num[idx] = (ep - fn(p)) / (2 * h)
    return num


def test_mmff_gradient_matches_finite_differences():
    system, xyz, _ = S.random_mmff_system(2, 6, 9, seed=5)
    for m in range(2):
        pos = xyz[m] + np.random.default_rng(m).normal(0, 0.05, xyz[m].shape)
        e, g, per = oracle.ff_energy_grad("mmff", system.atom_counts, system.tables, m, pos)
        assert abs(per.sum() - e) < 1e-9
        num = _fd(lambda p: oracle.ff_energy_grad("mmff", system.atom_counts, system.tables, m, p, False)[0], pos)
        assert np.abs(num - g).max() < 2e-4 * max(1.0, np.abs(g).max())


def test_dg_and_etk_gradients_match_their_definition():
    flat, mols = S.random_embed_molecules(2, 5, 8, seed=6)
    rng = np.random.default_rng(0)
    for m in range(2):
        n = flat.atom_counts[m]
        pos = rng.normal(0, 1.5, (n, 4))
        # DG: chiral and fourth-dimension gradients carry RDKit's missing factor 2 -> check the distance part alone
        dist_only = {k: v for k, v in flat.dg.tables.items()}
        z = np.zeros(flat.dg.n_mols + 1, dtype=np.int32)
        dist_only["chiral"] = (z, flat.dg.tables["chiral"][1][:0], flat.dg.tables["chiral"][2][:0])
        dist_only["fourth"] = (z, flat.dg.tables["fourth"][1][:0], flat.dg.tables["fourth"][2][:0])
        e, g, _ = oracle.ff_energy_grad("dg", flat.dg.atom_counts, dist_only, m, pos, dim=4)
        num = _fd(lambda p: oracle.ff_energy_grad("dg", flat.dg.atom_counts, dist_only, m, p, False, dim=4)[0], pos)
        assert np.abs(num - g).max() < 1e-4 * max(1.0, np.abs(g).max())
        e_all, g_all, _ = oracle.ff_energy_grad("dg", flat.dg.atom_counts, flat.dg.tables, m, pos, dim=4, chiral_weight=1.0,
                                                fourth_dim_weight=0.1)
        assert np.allclose(g_all[:, 3] - g[:, 3], 0.1 * pos[:, 3])  # E = w x4^2 but gradient w x4 (RDKit quirk)
        # ETK without torsions (the 6-fold torsion gradient uses V5, RDKit quirk): distance/angle/improper terms are exact
        no_tor = dict(flat.etk.tables)
        no_tor["torsion"] = (z, flat.etk.tables["torsion"][1][:0], flat.etk.tables["torsion"][2][:0])
        e, g, _ = oracle.ff_energy_grad("etk", flat.etk.atom_counts, no_tor, m, pos)
        num = _fd(lambda p: oracle.ff_energy_grad("etk", flat.etk.atom_counts, no_tor, m, p, False)[0], pos)
        assert np.abs(num[:, :3] - g[:, :3]).max() < 2e-4 * max(1.0, np.abs(g).max())
        assert np.abs(g[:, 3]).max() == 0.0
        # torsions: V6 = 0 in the synthetic tables, so the quirk is inactive and the gradient is exact too
        e, g, _ = oracle.ff_energy_grad("etk", flat.etk.atom_counts, flat.etk.tables, m, pos)
        num = _fd(lambda p: oracle.ff_energy_grad("etk", flat.etk.atom_counts, flat.etk.tables, m, p, False)[0], pos)
        assert np.abs(num[:, :3] - g[:, :3]).max() < 2e-4 * max(1.0, np.abs(g).max())


def test_etk_window_refresh():
    flat, _ = S.random_embed_molecules(1, 5, 6, seed=8)
    n = flat.atom_counts[0]
    ref = np.random.default_rng(1).normal(0, 1.5, (n, 4))
    only12 = dict(flat.etk.tables)
    z = np.zeros(2, dtype=np.int32)
    for k in ("torsion", "improper", "dist13", "angle13", "longrange"):
        only12[k] = (z, flat.etk.tables[k][1][:0], flat.etk.tables[k][2][:0])
    # at the reference geometry every re-centred 1-2 window contains its own distance: zero energy
    e, _, _ = oracle.ff_energy_grad("etk", flat.etk.atom_counts, only12, 0, ref, False, ref_pos=ref)
    assert e == 0.0
    e_fixed, _, _ = oracle.ff_energy_grad("etk", flat.etk.atom_counts, only12, 0, ref, False)
    assert e_fixed > 0.0


def test_mmff_minimize_lowers_energy_and_reports_final_energy():
    system, xyz, _ = S.random_mmff_system(3, 6, 10, seed=9)
    batch = ConformerBatch.from_coords(system, [[x] for x in xyz])
    pos, e, conv, iters = oracle.ff_minimize("mmff", system.atom_counts, system.tables, batch.conf_mol, batch.atom_starts,
                                             batch.positions, 500, 1e-4)
    for c in range(3):
        a0, a1 = batch.atom_starts[c], batch.atom_starts[c + 1]
        e0 = oracle.ff_energy_grad("mmff", system.atom_counts, system.tables, c, batch.positions[a0:a1], False)[0]
        e1 = oracle.ff_energy_grad("mmff", system.atom_counts, system.tables, c, pos[a0:a1], False)[0]
        assert e1 < e0 and abs(e1 - e[c]) < 1e-9


def test_uniform01_is_a_pure_function():
    assert oracle.uniform01(7, 1, 2, 3) == oracle.uniform01(7, 1, 2, 3)
    u = np.array([oracle.uniform01(7, 0, 0, i) for i in range(2000)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.03


def test_uff_gradient_matches_finite_differences_and_hand_values():
    system, xyz, _ = S.random_uff_system(2, 6, 9, seed=5)
    for m in range(2):
        pos = xyz[m] + np.random.default_rng(m).normal(0, 0.05, xyz[m].shape)
        e, g, _ = oracle.ff_energy_grad("uff", system.atom_counts, system.tables, m, pos)
        num = _fd(lambda p: oracle.ff_energy_grad("uff", system.atom_counts, system.tables, m, p, False)[0], pos)
        assert np.abs(num - g).max() < 1e-6 * max(1.0, np.abs(g).max())
    # harmonic bond and 12-6 at the minimum: E = -wellDepth, zero force
    one = FlatSystem.from_molecules("uff", [2], [{"bond": ([[0, 1]], [[1.5, 700.0]]), "vdw": ([[0, 1]], [[1.7, 0.1, 10.0]])}])
    e, g, _ = oracle.ff_energy_grad("uff", one.atom_counts, one.tables, 0, np.array([[0, 0, 0], [1.7, 0, 0.0]]))
    assert abs(e - (0.5 * 700.0 * 0.04 - 0.1)) < 1e-12
    assert np.allclose(g, [[-700.0 * 0.2, 0, 0], [700.0 * 0.2, 0, 0]], atol=1e-9)
