"""Golden fixtures of tests/golden/ (written by tests/golden/make_golden.py; provenance in its docstring).

CPU: the oracle must still produce them (freezes oracle/). GPU: the library must produce them through the C-ABI.
"""
import os

import numpy as np
import pytest

import oracle
from nvmolkit_b200 import synthetic as S

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
A = np.load(os.path.join(HERE, "path_a.npz"))
B = np.load(os.path.join(HERE, "path_b.npz"))
CUTS = ((0.2, "0p2"), (0.3, "0p3"), (0.5, "0p5"))


def _inputs_a():
    a = S.random_fingerprints(200, seed=int(A["cross_seed_a"]), near_dups=40)
    b = S.random_fingerprints(300, seed=int(A["cross_seed_b"]), near_dups=60)
    fp = S.clustered_fingerprints(60, 25, seed=int(A["butina_seed"]))
    g = S.random_molgraphs(40, seed=int(A["morgan_seed"]))
    return a, b, fp, g


def test_golden_files_match_the_oracle():
    a, b, fp, g = _inputs_a()
    assert (oracle.similarity_cross(a, b) == A["cross_tanimoto"]).all()
    assert (oracle.similarity_cross(a, b, metric="cosine") == A["cross_cosine"]).all()
    for cutoff, key in CUTS:
        assert (oracle.count_ge(fp, fp, cutoff) == A[f"counts_{key}"]).all()
        ids, cen = oracle.butina_fp(fp, cutoff)
        assert (ids == A[f"butina_ids_{key}"]).all() and (cen == A[f"butina_centroids_{key}"]).all()
    for r in range(4):
        bits = oracle.morgan(g.atom_starts, g.bond_starts, g.atom_inv, g.bond_inv, g.bond_a, g.bond_b, r, 2048)
        assert (bits == A[f"morgan_r{r}_2048"]).all()
    system, xyz, _ = S.random_mmff_system(6, 10, 30, seed=int(B["mmff_seed"]))
    e = np.array([oracle.ff_energy_grad("mmff", system.atom_counts, system.tables, m, x)[0] for m, x in enumerate(xyz)])
    assert np.allclose(e, B["mmff_energy"], rtol=1e-13, atol=1e-12)  # (summation order of the diagonal term sort is fixed)
    usys, uxyz, _ = S.random_uff_system(4, 8, 20, seed=int(B["uff_seed"]))
    eu = np.array([oracle.ff_energy_grad("uff", usys.atom_counts, usys.tables, m, x)[0] for m, x in enumerate(uxyz)])
    assert np.allclose(eu, B["uff_energy"], rtol=1e-13, atol=1e-12)
    x, e4, _, _ = oracle.poly_minimize(4, np.ones(8), np.arange(8, dtype=np.float64), np.zeros(8), 400, 1e-4)
    assert np.allclose(x, B["quartic_x"], atol=1e-12) and abs(e4 - float(B["quartic_e"])) < 1e-15
    assert np.abs(B["quartic_x"] - np.arange(8)).max() < 0.05  # tests/test_bfgs_minimizer.cu:1014: minimum at x_i = i


@pytest.mark.gpu
def test_gpu_reproduces_golden_path_a():
    import torch

    from nvmolkit_b200.clustering import fused_butina_device
    from nvmolkit_b200.fingerprints import MorganFingerprintGenerator
    from nvmolkit_b200.similarity import crossCosineSimilarity, crossTanimotoSimilarity

    dev = torch.device("cuda", 0)
    a, b, fp, g = _inputs_a()
    ta, tb = torch.from_numpy(a.view(np.int32)).to(dev), torch.from_numpy(b.view(np.int32)).to(dev)
    assert (crossTanimotoSimilarity(ta, tb).torch().cpu().numpy() == A["cross_tanimoto"]).all()
    got = crossCosineSimilarity(ta, tb).torch().cpu().numpy()
    assert np.abs(got - A["cross_cosine"]).max() <= 1e-15  # sqrt/div are correctly rounded on both sides; see similarity tests
    tf = torch.from_numpy(fp.view(np.int32)).to(dev)
    for cutoff, key in CUTS:
        ids, cen = fused_butina_device(tf, cutoff)
        assert (ids.cpu().numpy() == A[f"butina_ids_{key}"]).all() and (cen.cpu().numpy() == A[f"butina_centroids_{key}"]).all()
    for r in range(4):
        bits = MorganFingerprintGenerator(r, 2048).GetFingerprints(g).torch().cpu().numpy().view(np.uint32)
        assert (bits == A[f"morgan_r{r}_2048"]).all()


@pytest.mark.gpu
def test_gpu_reproduces_golden_path_b():
    from nvmolkit_b200.forcefield import ConformerBatch
    from nvmolkit_b200.minimizer import energy_and_grad, poly_minimize

    system, xyz, _ = S.random_mmff_system(6, 10, 30, seed=int(B["mmff_seed"]))
    batch = ConformerBatch.from_coords(system, [[x] for x in xyz])
    e, g = energy_and_grad(system, batch)
    assert np.allclose(e.cpu().numpy(), B["mmff_energy"], rtol=1e-11, atol=1e-10)
    gg = B["mmff_grad"].reshape(-1, 3)
    assert np.abs(g.cpu().numpy().reshape(-1, 3) - gg).max() < 1e-9 * max(1.0, np.abs(gg).max())
    usys, uxyz, _ = S.random_uff_system(4, 8, 20, seed=int(B["uff_seed"]))
    ub = ConformerBatch.from_coords(usys, [[x] for x in uxyz])
    eu, _ = energy_and_grad(usys, ub)
    assert np.allclose(eu.cpu().numpy(), B["uff_energy"], rtol=1e-11, atol=1e-10)
    x, e4, status, _ = poly_minimize(np.array([0, 8], dtype=np.int32), 4, np.ones(8), np.arange(8, dtype=np.float64), np.zeros(8), 400, 1e-4,
                                     False)
    assert np.abs(x.cpu().numpy() - B["quartic_x"]).max() < 1e-4 and abs(float(e4[0]) - float(B["quartic_e"])) < 1e-10
