"""Writes the golden fixtures of tests/golden/ (run here, on the CPU: `python tests/golden/make_golden.py`).

Provenance. The reference (nvMolKit) cannot be built or imported in this container (RDKit / Boost / a GPU are absent,
DESIGN.md §2), so these vectors are NOT outputs of the reference. They are outputs of the CPU oracle (`oracle/*.c`),
which `tests/test_oracle_golden.py` / `tests/test_oracle_path_b.py` pin against the known answers the reference's own
tests hold. Their job: (1) freeze the oracle — any later edit of `oracle/` that changes a result fails
`test_golden_files_match_the_oracle` on the CPU; (2) give the GPU tests inputs + expected outputs that do not depend on
running the oracle on the GPU box. Inputs are regenerated from the seeds stored beside the outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from nvmolkit_b200 import synthetic as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def path_a():
    out = {}
    # config 1 shape, reduced: 200 x 300 cross Tanimoto / cosine of Bernoulli(0.025) fingerprints with planted near-duplicates
    a = S.random_fingerprints(200, seed=101, near_dups=40)
    b = S.random_fingerprints(300, seed=102, near_dups=60)
    out["cross_seed_a"], out["cross_seed_b"] = 101, 102
    out["cross_tanimoto"] = oracle.similarity_cross(a, b)
    out["cross_cosine"] = oracle.similarity_cross(a, b, metric="cosine")
    # thresholded neighbour counts and Butina ids (fused definition = RDKit ClusterData(reordering=True))
    fp = S.clustered_fingerprints(60, 25, seed=103)
    out["butina_seed"] = 103
    for cutoff in (0.2, 0.3, 0.5):
        key = str(cutoff).replace(".", "p")
        out[f"counts_{key}"] = oracle.count_ge(fp, fp, cutoff)
        ids, cen = oracle.butina_fp(fp, cutoff)
        out[f"butina_ids_{key}"], out[f"butina_centroids_{key}"] = ids, cen
    # Morgan bits of seeded random molecular graphs, radii 0..3
    g = S.random_molgraphs(40, seed=104)
    out["morgan_seed"] = 104
    for r in range(4):
        out[f"morgan_r{r}_2048"] = oracle.morgan(g.atom_starts, g.bond_starts, g.atom_inv, g.bond_inv, g.bond_a, g.bond_b, r, 2048)
    np.savez_compressed(os.path.join(HERE, "path_a.npz"), **out)
    return out


def path_b():
    out = {}
    system, xyz, _ = S.random_mmff_system(6, 10, 30, seed=201)
    out["mmff_seed"] = 201
    e, g = [], []
    for m, x in enumerate(xyz):
        em, gm, _ = oracle.ff_energy_grad("mmff", system.atom_counts, system.tables, m, x)
        e.append(em)
        g.append(gm)
    out["mmff_energy"] = np.array(e)
    out["mmff_grad"] = np.concatenate(g)
    usys, uxyz, _ = S.random_uff_system(4, 8, 20, seed=202)
    out["uff_seed"] = 202
    out["uff_energy"] = np.array([oracle.ff_energy_grad("uff", usys.atom_counts, usys.tables, m, x)[0] for m, x in enumerate(uxyz)])
    # the reference's analytic BFGS system: E = sum_i (x_i - i)^4, tests/test_bfgs_minimizer.cu:822-930
    x, e4, _, _ = oracle.poly_minimize(4, np.ones(8), np.arange(8, dtype=np.float64), np.zeros(8), 400, 1e-4)
    out["quartic_x"], out["quartic_e"] = x, e4
    np.savez_compressed(os.path.join(HERE, "path_b.npz"), **out)
    return out


if __name__ == "__main__":
    a, b = path_a(), path_b()
    print("wrote", sorted(a), sorted(b))
