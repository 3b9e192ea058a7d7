"""Restraint ("constraint") terms and the BatchedForcefield API. CPU: the oracle's restraint terms against
hand-computed values and finite differences, and the host-side spec -> table conversion; GPU: kernels vs oracle, API."""
import numpy as np
import pytest

import oracle
from nvmolkit_b200 import synthetic as S
from nvmolkit_b200.batchedForcefield import (MMFFBatchedForcefield, UFFBatchedForcefield, _AngleConstraint, _DistanceConstraint,
                                              _PositionConstraint, _TorsionConstraint, angle_deg, dihedral_deg, restraint_tables)
from nvmolkit_b200.forcefield import LAYOUT, ConformerBatch, FlatSystem


def _restraint_only_system(kind, n_atoms, tabs):
    terms = {}
    for name, k, p in LAYOUT[kind]:
        ix, pr = tabs.get(name, ([], []))
        terms[name] = (np.array(ix, np.int16).reshape(-1, k), np.array(pr, np.float64).reshape(-1, p))
    return FlatSystem.from_molecules(kind, [n_atoms], [terms])


def test_restraint_known_answers_and_finite_differences():
    xyz = np.array([[0.0, 0.0, 0.0], [1.5, 0.0, 0.0], [1.5, 1.5, 0.0], [3.0, 1.5, 1.0], [0.3, -0.2, 0.8]])
    # distance 0-1 is 1.5: window [1.0, 1.2], k = 10 -> 1/2 * 10 * 0.3^2
    sysd = _restraint_only_system("mmff", 5, {"distc": ([(0, 1)], [(1.0, 1.2, 10.0)])})
    e, g, _ = oracle.ff_energy_grad("mmff", sysd.atom_counts, sysd.tables, 0, xyz)
    assert abs(e - 0.45) < 1e-12 and np.allclose(g[0], [-3.0, 0, 0]) and np.allclose(g[1], [3.0, 0, 0])
    # angle 0-1-2 is 90 degrees: window [100, 120], k = 2 -> 2 * 10^2 (degrees, no 1/2)
    sysa = _restraint_only_system("uff", 5, {"anglec": ([(0, 1, 2)], [(100.0, 120.0, 2.0)])})
    assert abs(oracle.ff_energy_grad("uff", sysa.atom_counts, sysa.tables, 0, xyz, False)[0] - 200.0) < 1e-9
    # position: atom 4 displaced by 0.5 from its anchor, free radius 0.2, k = 4 -> 1/2 * 4 * 0.3^2
    sysp = _restraint_only_system("mmff", 5, {"posc": ([(4,)], [(0.3, -0.2, 0.3, 0.2, 4.0)])})
    assert abs(oracle.ff_energy_grad("mmff", sysp.atom_counts, sysp.tables, 0, xyz, False)[0] - 0.18) < 1e-12
    # torsion: periodic window across +-180
    phi = dihedral_deg(xyz, 0, 1, 2, 3)
    syst = _restraint_only_system("mmff", 5, {"torsc": ([(0, 1, 2, 3)], [(170.0, -170.0, 1.5)])})
    want = 1.5 * min(abs(((phi - 170.0 + 180) % 360) - 180), abs(((phi + 170.0 + 180) % 360) - 180)) ** 2
    assert abs(oracle.ff_energy_grad("mmff", syst.atom_counts, syst.tables, 0, xyz, False)[0] - want) < 1e-9 * want
    # gradients of all four by central differences
    rng = np.random.default_rng(3)
    tabs = {"distc": ([(0, 3), (1, 4)], [(0.5, 1.0, 7.0), (3.0, 4.0, 3.0)]), "posc": ([(2,)], [(1.0, 1.0, 0.4, 0.1, 5.0)]),
            "anglec": ([(0, 1, 2), (4, 3, 2)], [(100.0, 110.0, 0.3), (10.0, 20.0, 0.2)]),
            "torsc": ([(0, 1, 2, 3), (4, 0, 1, 2)], [(-60.0, -30.0, 0.05), (100.0, 140.0, 0.02)])}
    sys4 = _restraint_only_system("mmff", 5, tabs)
    x0 = xyz + rng.normal(0, 0.05, xyz.shape)
    e0, g0, _ = oracle.ff_energy_grad("mmff", sys4.atom_counts, sys4.tables, 0, x0)
    num = np.zeros_like(x0)
    for a in range(5):
        for c in range(3):
            xp, xm = x0.copy(), x0.copy()
            xp[a, c] += 1e-6
            xm[a, c] -= 1e-6
            num[a, c] = (oracle.ff_energy_grad("mmff", sys4.atom_counts, sys4.tables, 0, xp, False)[0] -
                         oracle.ff_energy_grad("mmff", sys4.atom_counts, sys4.tables, 0, xm, False)[0]) / 2e-6
    assert e0 > 0 and np.abs(num - g0).max() < 1e-5 * max(1.0, np.abs(g0).max())


def test_restraint_specs_to_tables():
    xyz = np.array([[0.0, 0.0, 0.0], [1.5, 0.0, 0.0], [1.5, 1.5, 0.0], [3.0, 1.5, 1.0]])
    t = restraint_tables([_DistanceConstraint(0, 1, True, -0.1, 0.1, 9.0), _PositionConstraint(3, 0.2, 4.0),
                          _AngleConstraint(0, 1, 2, True, -5.0, 5.0, 1.0), _TorsionConstraint(0, 1, 2, 3, True, 170.0, 200.0, 2.0)], xyz)
    assert np.allclose(t["distc"][1], [(1.4, 1.6, 9.0)]) and t["posc"][1] == [(3.0, 1.5, 1.0, 0.2, 4.0)]
    assert np.allclose(t["anglec"][1], [(85.0, 95.0, 1.0)]) and abs(angle_deg(xyz, 0, 1, 2) - 90.0) < 1e-12
    phi = dihedral_deg(xyz, 0, 1, 2, 3)
    mn, mx, _k = t["torsc"][1][0]
    assert -180.0 <= mn <= 180.0 and -180.0 <= mx <= 180.0 and abs(((mn - (phi + 170.0) + 180) % 360) - 180) < 1e-9
    with pytest.raises(ValueError):
        restraint_tables([_DistanceConstraint(0, 1, False, 2.0, 1.0, 1.0)], xyz)
    with pytest.raises(ValueError):
        restraint_tables([_AngleConstraint(0, 1, 2, False, 10.0, 200.0, 1.0)], xyz)


def _flat(kind, n, seed):
    if kind == "mmff":
        from nvmolkit_b200.mmffOptimization import FlatMMFFMolecules as F

        system, xyz, _ = S.random_mmff_system(n, 5, 14, seed=seed)
    else:
        from nvmolkit_b200.uffOptimization import FlatUFFMolecules as F

        system, xyz, _ = S.random_uff_system(n, 5, 14, seed=seed)
    rng = np.random.default_rng(seed)
    return F(system, ConformerBatch.from_coords(system, [[x, x + rng.normal(0, 0.1, x.shape)] for x in xyz])), xyz


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["mmff", "uff"])
def test_batched_forcefield_with_restraints_equals_oracle(cuda, kind):
    flat, xyz = _flat(kind, 4, 71)
    ff = (MMFFBatchedForcefield if kind == "mmff" else UFFBatchedForcefield)(flat)
    assert len(ff) == 4 and ff[1].num_atoms == len(xyz[1])
    plain = ff.compute_energy()
    ff[0].add_distance_constraint(0, 1, False, 3.0, 3.5, 50.0)
    ff[0].add_torsion_constraint(0, 1, 2, 3, True, 30.0, 60.0, 0.1)
    ff[1].add_position_constraint(2, 0.0, 20.0)
    ff[2].add_angle_constraint(0, 1, 2, True, 10.0, 15.0, 0.5)
    ff[2].add_distance_constraint(1, 3, True, 0.2, 0.3, 30.0)
    with pytest.raises(IndexError):
        ff[3].add_position_constraint(10 ** 6, 0.1, 1.0)
    e, g = ff.compute_energy(), ff.compute_gradients()
    assert [len(x) for x in e] == [2, 2, 2, 2] and e[3] == plain[3] and e[0][0] > plain[0][0]
    sysc, b = ff._system, ff._conf_batch
    for c in range(b.n_conf):
        a0, a1 = b.atom_starts[c], b.atom_starts[c + 1]
        eo, go, _ = oracle.ff_energy_grad(kind, sysc.atom_counts, sysc.tables, c, b.positions[a0:a1])
        m, k = int(flat.batch.conf_mol[c]), c % 2
        assert abs(e[m][k] - eo) <= 1e-10 * max(1.0, abs(eo))
        assert np.abs(np.array(g[m][k]).reshape(-1, 3) - go).max() <= 1e-8 * max(1.0, np.abs(go).max())
    energies, converged = ff.minimize(maxIters=300)
    pos = ff.positions()
    # the position restraint (free radius 0) holds atom 2 of molecule 1 near its anchor; minimised energies do not rise
    for k in range(2):
        anchor = flat.batch.positions[flat.batch.atom_starts[2 + k]:flat.batch.atom_starts[3 + k]][2]
        assert np.linalg.norm(pos[1][k][2] - anchor) < 0.35
    assert all(energies[m][k] <= e[m][k] + 1e-9 for m in range(4) for k in range(2))
    pos_o, e_o, conv_o, _ = oracle.ff_minimize(kind, sysc.atom_counts, sysc.tables, b.conf_mol, b.atom_starts, b.positions, 300, 1e-4)
    flat_e = np.array([energies[int(flat.batch.conf_mol[c])][c % 2] for c in range(b.n_conf)])
    both = np.array([converged[int(flat.batch.conf_mol[c])][c % 2] for c in range(b.n_conf)]) & (conv_o == 1)
    assert both.any() and (np.abs(flat_e[both] - e_o[both]) <= 1e-4 * np.maximum(1.0, np.abs(e_o[both]))).mean() >= 0.75
