"""The RDKit seam, EXECUTED against tests/fake_rdkit: a stand-in for exactly the
RDKit calls nvmolkit_b200/rdkit_adapter.py makes, answering from the synthetic pseudo-molecules. Checks the adapters'
index conventions, table layouts, fallbacks and error contracts; RDKit's chemistry itself is out of reach here."""
import os
import sys

import numpy as np
import pytest

import oracle
from nvmolkit_b200 import synthetic as S
from nvmolkit_b200.forcefield import LAYOUT

FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_rdkit")


@pytest.fixture
def fake_rdkit(monkeypatch):
    monkeypatch.syspath_prepend(FAKE)
    for name in [m for m in sys.modules if m == "rdkit" or m.startswith("rdkit.")]:
        monkeypatch.delitem(sys.modules, name)
    import rdkit
    from rdkit import Chem  # noqa: F401

    assert rdkit.__file__.startswith(FAKE)
    yield Chem
    for name in [m for m in sys.modules if m == "rdkit" or m.startswith("rdkit.")]:
        sys.modules.pop(name, None)


def _canon(idx, par):
    idx = np.asarray(idx, np.int64)
    order = np.lexsort(tuple(idx[:, k] for k in range(idx.shape[1] - 1, -1, -1))) if len(idx) else np.zeros(0, np.int64)
    return idx[order], np.asarray(par)[order]


def _mols(n, seed, with_uff=False):
    rng = np.random.default_rng(seed)
    if with_uff:
        sysu, _xyz, mols = S.random_uff_system(n, 5, 12, seed=seed)
        flat = None
        for k, m in enumerate(mols):
            m["planar"] = {a for a in range(len(m["z"])) if len(m["nbrs"][a]) == 3 and m["z"][a] in (6, 7)}
            m["uff"] = {name: (sysu.tables[name][1][sysu.tables[name][0][k]:sysu.tables[name][0][k + 1]],
                               sysu.tables[name][2][sysu.tables[name][0][k]:sysu.tables[name][0][k + 1]])
                        for name, _k, _p in LAYOUT["uff"][:5]}
    else:
        flat, mols = S.random_embed_molecules(n, 5, 12, seed=seed)
    for m in mols:
        # (the generator draws the stretch-bend rest lengths independently of the bond table; RDKit - and the adapter -
        # take them from the bond parameters, so make the reference tables consistent with that)
        r0 = {}
        for (i, j), p in zip(m["terms"]["bond"][0].tolist(), m["terms"]["bond"][1]):
            r0[(i, j)] = r0[(j, i)] = p[0]
        sb_i, sb_p = m["terms"]["strbend"]
        sb_p = np.array(sb_p)
        for row, (i, j, k) in enumerate(sb_i.tolist()):
            sb_p[row, 1], sb_p[row, 2] = r0[(i, j)], r0[(k, j)]
        m["terms"]["strbend"] = (sb_i, sb_p)
        oi, op = m["terms"]["oop"]  # one koop per centre (the generator draws one per Wilson-angle permutation)
        op = np.array(op)
        for row in range(len(oi)):
            op[row] = op[3 * (row // 3)]
        m["terms"]["oop"] = (oi, op)
        q = rng.normal(0, 0.2, len(m["z"]))
        m["charges"] = q
        pairs = m["terms"]["ele"][0]
        m["terms"]["ele"] = (pairs, np.stack([q[pairs[:, 0]] * q[pairs[:, 1]], np.ones(len(pairs)),
                                              (m["topo"][pairs[:, 0], pairs[:, 1]] == 3).astype(float)], axis=1))
    return flat, mols


def test_morgan_invariant_adapter_feeds_the_fingerprint_path(fake_rdkit):
    """molgraph.from_rdkit walks atoms / bonds / ring info the way MorganInvariantsGenerator does
    (src/morgan_fingerprint_common.cpp:43-124) and its CSR batch is what b200mol_morgan / the oracle consume."""
    from nvmolkit_b200.molgraph import atom_invariant, from_rdkit

    Chem = fake_rdkit
    _flat, mols = S.random_embed_molecules(5, 5, 12, seed=17)
    mols[1]["formal_charge"] = [1 if a == 0 else 0 for a in range(len(mols[1]["z"]))]
    mols[2]["isotope_delta"] = [1.0 if a == 1 else 0.0 for a in range(len(mols[2]["z"]))]
    mols[3]["implicit_hs"] = [2 if a == 0 else 0 for a in range(len(mols[3]["z"]))]
    ring = [(0, 1, 2)]  # (only membership matters to the invariant)
    rd = [Chem.Mol(m, rings=ring if k == 4 else (), bond_types=[Chem.BondType.DOUBLE if b == 0 else Chem.BondType.SINGLE
                                                                for b in range(len(m["bonds"]))]) for k, m in enumerate(mols)]
    batch = from_rdkit(rd)
    assert batch.atom_starts.tolist() == np.concatenate([[0], np.cumsum([len(m["z"]) for m in mols])]).tolist()
    assert batch.bond_starts.tolist() == np.concatenate([[0], np.cumsum([len(m["bonds"]) for m in mols])]).tolist()
    for k, m in enumerate(mols):
        a0, b0 = batch.atom_starts[k], batch.bond_starts[k]
        for a in range(len(m["z"])):
            nbr_h = sum(1 for j in m["nbrs"][a] if m["z"][j] == 1)
            hs = m.get("implicit_hs", [0] * len(m["z"]))[a]
            want = atom_invariant(int(m["z"][a]), hs + len(m["nbrs"][a]), hs + nbr_h, m.get("formal_charge", [0] * len(m["z"]))[a],
                                  int(m.get("isotope_delta", [0] * len(m["z"]))[a]), k == 4 and a in ring[0])
            assert batch.atom_inv[a0 + a] == want, (k, a)
        for b, (i, j) in enumerate(m["bonds"]):
            assert (batch.bond_a[b0 + b], batch.bond_b[b0 + b]) == (i, j)
            assert batch.bond_inv[b0 + b] == (2 if b == 0 else 1)
    # the batch goes straight into the fingerprint path (CPU twin here; the GPU kernel is pinned to it bit for bit)
    bits = oracle.morgan(batch.atom_starts, batch.bond_starts, batch.atom_inv, batch.bond_inv, batch.bond_a, batch.bond_b, 2, 2048)
    assert bits.shape == (5, 64) and all(int(np.unpackbits(r.view(np.uint8)).sum()) > 0 for r in bits)
    with pytest.raises(ValueError):
        from_rdkit([None])


def test_mmff_adapter_reproduces_the_parameter_tables(fake_rdkit):
    Chem = fake_rdkit
    from nvmolkit_b200 import rdkit_adapter as A

    _flat, mols = _mols(4, 41)
    rd = [Chem.Mol(m, conformers=[m["xyz"], m["xyz"] + 0.1]) for m in mols]
    out = A.mmff_from_rdkit(rd)
    assert out.batch.n_conf == 8 and out.conf_ids == [[0, 1]] * 4
    for k, m in enumerate(mols):
        for name in ("bond", "angle", "torsion", "vdw", "ele"):
            st, ix, pr = out.system.tables[name]
            gi, gp = _canon(ix[st[k]:st[k + 1]], pr[st[k]:st[k + 1]])
            wi, wp = m["terms"][name]
            if name in ("vdw", "ele"):
                keep = np.ones(len(wi), bool) if name == "vdw" else np.abs(wp[:, 0]) > 1e-10
                wi, wp = wi[keep], wp[keep]
            wi, wp = _canon(wi, wp)
            assert len(gi) == len(wi), (name, len(gi), len(wi))
            same = (gi == wi).all(axis=1) | (gi == wi[:, ::-1]).all(axis=1)
            assert same.all() or sorted(map(tuple, np.sort(gi, 1).tolist())) == sorted(map(tuple, np.sort(wi, 1).tolist())), name
        # energies through the oracle agree with the generator's own tables (same terms, any order)
        e_a = oracle.ff_energy_grad("mmff", out.system.atom_counts, out.system.tables, k, m["xyz"], False)[0]
        ref_tabs = {n: (np.array([0, len(m["terms"][n][0])], np.int32), np.asarray(m["terms"][n][0], np.int16).reshape(-1, kk),
                        np.asarray(m["terms"][n][1], np.float64).reshape(-1, pp)) for n, kk, pp in LAYOUT["mmff"] if n in m["terms"]}
        e_r = oracle.ff_energy_grad("mmff", np.array([len(m["z"])], np.int32), ref_tabs, 0, m["xyz"], False)[0]
        assert abs(e_a - e_r) <= 1e-9 * max(1.0, abs(e_r))


def test_mmff_adapter_handles_missing_term_types_and_errors(fake_rdkit):
    Chem = fake_rdkit
    from nvmolkit_b200 import rdkit_adapter as A

    _flat, mols = _mols(2, 42)
    m = mols[0]
    for name, k, p in LAYOUT["mmff"][:7]:  # a molecule with bonds only (empty term lists crashed the reshape)
        if name != "bond":
            m["terms"][name] = (np.zeros((0, k), np.int16), np.zeros((0, p)))
    m["charges"] = np.zeros(len(m["z"]))
    out = A.mmff_from_rdkit([Chem.Mol(m, conformers=[m["xyz"]])])
    assert len(out.system.tables["bond"][1]) == len(m["bonds"])
    assert all(out.system.tables[n][1].shape == (0, k) for n, k, _p in LAYOUT["mmff"] if n != "bond")  # restraints too
    mols[1]["no_mmff"] = True
    with pytest.raises(ValueError) as e:
        A.mmff_from_rdkit([None, Chem.Mol(mols[1], conformers=[mols[1]["xyz"]])])
    assert e.value.args[1] == {"none": [0], "no_params": [1]}


def test_uff_adapter(fake_rdkit):
    Chem = fake_rdkit
    from nvmolkit_b200 import rdkit_adapter as A

    _flat, mols = _mols(3, 43, with_uff=True)
    rd = [Chem.Mol(m, conformers=[m["xyz"]]) for m in mols]
    out = A.uff_from_rdkit(rd, vdwThreshold=10.0)
    H = Chem.HybridizationType
    for k, (m, r) in enumerate(zip(mols, rd)):
        st, ix, pr = out.system.tables["bond"]
        gi, gp = _canon(np.sort(ix[st[k]:st[k + 1]], 1), pr[st[k]:st[k + 1]])
        wi, wp = _canon(np.sort(m["uff"]["bond"][0], 1), m["uff"]["bond"][1])
        assert np.array_equal(gi, wi) and np.allclose(gp, wp)
        st, ix, pr = out.system.tables["angle"]
        for row, p in zip(ix[st[k]:st[k + 1]], pr[st[k]:st[k + 1]]):
            order = int(p[2])
            assert order == (3 if r.hyb[row[1]] == H.SP2 else 0)  # sp2 centres: cos 3 theta; sp3: the general C0..C2 form
            if order == 0:
                c2 = 1.0 / (4.0 * max(np.sin(p[0]) ** 2, 1e-8))
                assert np.allclose(p[3:], [c2 * (2 * np.cos(p[0]) ** 2 + 1), -4 * c2 * np.cos(p[0]), c2])
        st, ix, pr = out.system.tables["torsion"]
        tor = pr[st[k]:st[k + 1]]
        assert set(np.unique(tor[:, 1])) <= {2.0, 3.0, 6.0} and set(np.unique(tor[:, 2])) <= {-1.0, 1.0}
        st, ix, pr = out.system.tables["vdw"]
        assert np.allclose(pr[st[k]:st[k + 1], 2], 10.0 * pr[st[k]:st[k + 1], 0])
        st, ix, pr = out.system.tables["inversion"]
        assert len(ix[st[k]:st[k + 1]]) % 3 == 0
    assert out.system.waves  # schedulable
    mols[0]["no_uff"] = True
    with pytest.raises(ValueError):
        A.uff_from_rdkit([Chem.Mol(mols[0], conformers=[mols[0]["xyz"]])])


def test_embed_adapter_builds_the_embedding_inputs(fake_rdkit, monkeypatch):
    Chem = fake_rdkit
    import nvmolkit_b200.dgprep as dgprep
    from nvmolkit_b200 import rdkit_adapter as A
    from nvmolkit_b200.builders import dg_terms_from_bounds
    from nvmolkit_b200.embedMolecules import EmbedParameters

    def cpu_smooth(mats, tol=0.0):  # the adapter smooths on the GPU; here the oracle stands in for that kernel
        res = [oracle.triangle_smooth(b, tol) for b in mats]
        return [r[0] for r in res], [bool(r[1]) for r in res]

    monkeypatch.setattr(dgprep, "triangle_smooth", cpu_smooth)
    _flat, mols = _mols(4, 44)
    CT, BT, BS = Chem.ChiralType, Chem.BondType, Chem.BondStereo
    rd = []
    for m in mols:
        quat = [a for a in range(len(m["z"])) if len(m["nbrs"][a]) == 4 and m["z"][a] == 6]
        chiral = {quat[0]: CT.CHI_TETRAHEDRAL_CCW} if quat else {}
        if len(quat) > 1:
            chiral[quat[1]] = CT.CHI_TETRAHEDRAL_CW
        rings = [tuple(quat[2:3] + m["nbrs"][quat[2]][:3]), tuple(quat[2:3] + m["nbrs"][quat[2]][1:4])] if len(quat) > 2 else []
        planar = sorted(m.get("planar", set()))
        types = [BT.SINGLE] * len(m["bonds"])
        stereo = {}
        for kb, (i, j) in enumerate(m["bonds"]):
            if i in planar and j in planar:
                types[kb] = BT.DOUBLE
                oi = [x for x in m["nbrs"][i] if x != j][0]
                oj = [x for x in m["nbrs"][j] if x != i][0]
                stereo[kb] = (BS.STEREOZ, (oi, oj))
                break
        m["exp_torsions"] = [{"atomIndices": (m["nbrs"][i][0] if m["nbrs"][i][0] != j else m["nbrs"][i][1], i, j,
                                              m["nbrs"][j][0] if m["nbrs"][j][0] != i else m["nbrs"][j][1]),
                              "V": [0.0, 0.0, 4.0], "signs": [1, 1, -1]}
                             for i, j in m["bonds"][:3] if len(m["nbrs"][i]) > 1 and len(m["nbrs"][j]) > 1]
        rd.append(Chem.Mol(m, chiral=chiral, rings=rings, bond_types=types, bond_stereo=stereo))
    flat = A.embed_molecules_from_rdkit(rd, EmbedParameters())
    assert len(flat) == 4
    for k, (m, r) in enumerate(zip(mols, rd)):
        sm, ok = oracle.triangle_smooth(m["bounds_raw"])
        assert ok
        want = dg_terms_from_bounds(sm)
        st, ix, pr = flat.dg.tables["dist"]
        gi, gp = _canon(ix[st[k]:st[k + 1]], pr[st[k]:st[k + 1]])
        wi, wp = _canon(*want["dist"])
        assert np.array_equal(gi, wi) and np.array_equal(gp, wp)
        # chiral sets: CCW -> positive volume window on the DG term and the check table, CW negative
        cst, cix, cpr = flat.checks.tables["chiral"]
        for row, p in zip(cix[cst[k]:cst[k + 1]], cpr[cst[k]:cst[k + 1]]):
            tag = r.chiral[int(row[0])]
            assert (p[0], p[1]) == ((5.0, 100.0) if tag == CT.CHI_TETRAHEDRAL_CCW else (-100.0, -5.0))
            assert sorted(row[1:].tolist()) == sorted(m["nbrs"][int(row[0])])
        tst, tix, tpr = flat.checks.tables["tetrahedral"]
        for row in tix[tst[k]:tst[k + 1]]:
            assert r.GetRingInfo().NumAtomRings(int(row[0])) >= 2 and int(row[0]) not in r.chiral
        sst, six, spr = flat.checks.tables["dbStereo"]
        assert (spr[sst[k]:sst[k + 1]] == -1.0).all()  # Z -> -1
        # ETK: three terms per improper centre, torsions copied through with six coefficients
        ist, iix, ipr = flat.etk.tables["improper"]
        assert (ist[k + 1] - ist[k]) == 3 * flat.checks.num_impropers[k]
        tst2, tix2, tpr2 = flat.etk.tables["torsion"]
        assert tpr2[tst2[k]:tst2[k + 1]].shape[1] == 12 and (tst2[k + 1] - tst2[k]) == len(m["exp_torsions"])
    # a molecule whose bounds do not smooth takes the relaxed matrix; one that still fails raises unless told to go on
    bad = dict(mols[0])
    broken = np.array(bad["bounds_raw"])
    broken[0, 1], broken[1, 0] = 0.5, 5.0  # lower bound above the upper bound
    bad["bounds_raw"], bad["bounds_relaxed"] = broken, mols[0]["bounds_raw"]
    flat2 = A.embed_molecules_from_rdkit([Chem.Mol(bad)], EmbedParameters())
    assert np.array_equal(flat2.dg.tables["dist"][2], flat.dg.tables["dist"][2][:len(flat2.dg.tables["dist"][2])])
    bad["bounds_relaxed"] = broken
    with pytest.raises(ValueError):
        A.embed_molecules_from_rdkit([Chem.Mol(bad)], EmbedParameters())
    p = EmbedParameters()
    p.ignoreSmoothingFailures = True
    assert len(A.embed_molecules_from_rdkit([Chem.Mol(bad)], p)) == 1


def test_conformer_write_back_and_pruning(fake_rdkit):
    Chem = fake_rdkit
    from nvmolkit_b200 import rdkit_adapter as A

    _flat, mols = _mols(1, 45)
    m = mols[0]
    mol = Chem.Mol(m, conformers=[m["xyz"]])
    A.write_back_conformers([mol], [[m["xyz"] + 1.0]])
    assert np.allclose(mol.GetConformers()[0].GetPositions(), m["xyz"] + 1.0)
    rot = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    mol2 = Chem.Mol(m)
    far = m["xyz"] + np.random.default_rng(1).normal(0, 1.0, m["xyz"].shape)
    A.add_conformers([mol2], [[m["xyz"], m["xyz"] @ rot.T + 3.0, far]], prune_rms_thresh=0.5)
    assert len(mol2.GetConformers()) == 2  # the rotated copy is the same conformer, the scrambled one is not
