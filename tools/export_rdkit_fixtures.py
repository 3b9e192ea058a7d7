#!/usr/bin/env python
"""Export RDKit-pinned fixtures for the "parity unpinned" rows of DESIGN.md 2 / SURVEY.md 8c.

Run this WHERE RDKIT EXISTS (any RDKit 2025.03 .. 2026.03; no GPU needed):

    python tools/export_rdkit_fixtures.py [--smiles FILE] [--n 64] [--out tests/golden/rdkit_fixtures.npz]

and commit the resulting .npz. `tests/test_rdkit_fixtures.py` then replays it (skipped while the file is absent):
  * Morgan: the flattened molecular graphs + RDKit's own ECFP bits (radius 0..3, 2048 bits)   -> oracle / GPU bit-exact
  * Tanimoto: RDKit BulkTanimotoSimilarity matrix of those bit vectors                         -> bit-exact (fp64)
  * Butina: RDKit ML.Cluster.Butina.ClusterData(reordering=True) on the distance matrix        -> identical clusters
  * MMFF / UFF: the flattened term tables (nvmolkit_b200.rdkit_adapter) + RDKit's energies, gradients and 200-iteration
    minimised energies on the same conformers (tests/test_mmff.cu:1521-1608 style)              -> <= 1e-4 relative
  * ETKDG: raw + smoothed bounds matrices, flattened DG / ETK / check tables, RDKit's EmbedMultipleConfs coordinates
    (per-molecule RMSD / bounds-violation statistics are compared, not coordinates)

The default molecule set is a small drug-like SMILES list embedded below (aspirin, ibuprofen, caffeine, ...); pass
--smiles benchmarks/data/chembl_10k.smi of the reference checkout for the BASELINE configs' real input.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT_SMILES = [
    "CC(=O)OC1=CC=CC=C1C(=O)O", "CC(C)CC1=CC=C(C=C1)C(C)C(=O)O", "CN1C=NC2=C1C(=O)N(C(=O)N2C)C", "CCCCC",
    "O=C(O)CC1CC1", "OCCCCO", "C[C@H](N)C(=O)O", "C1CCC2CCCCC2C1", "CC(=O)NC1=CC=C(O)C=C1", "C/C=C/C(=O)O",
    "CC#CCO", "c1ccc2[nH]ccc2c1", "CN(C)C(=O)c1ccccc1", "OC(=O)[C@@H]1CCCN1", "CCOC(=O)C1=CC=CN=C1", "CS(=O)(=O)Nc1ccccc1",
    "FC(F)(F)c1ccccc1", "CC(C)(C)OC(=O)N1CCNCC1", "O=C1CCCCC1", "NC(=O)c1cccnc1", "CCN(CC)CCOC(=O)c1ccc(N)cc1",
    "Clc1ccc(cc1)C(c1ccccc1)N1CCNCC1", "CC1(C)SC2C(NC(=O)Cc3ccccc3)C(=O)N2C1C(=O)O", "OP(=O)(O)OCC",
]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--smiles", default=None, help="SMILES file (one per line, first column)")
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--confs", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "rdkit_fixtures.npz"))
    args = ap.parse_args()

    import rdkit
    from rdkit import Chem, DataStructs
    from rdkit.Chem import AllChem, rdDistGeom, rdFingerprintGenerator
    from rdkit.Chem import rdForceFieldHelpers as FFH
    from rdkit.ML.Cluster import Butina

    from nvmolkit_b200 import rdkit_adapter as A
    from nvmolkit_b200.embedMolecules import EmbedParameters
    from nvmolkit_b200.molgraph import from_rdkit

    smiles = DEFAULT_SMILES
    if args.smiles:
        smiles = [line.split()[0] for line in open(args.smiles) if line.strip()][: args.n]
    mols = [m for m in (Chem.MolFromSmiles(s) for s in smiles[: args.n]) if m is not None]
    out = {"rdkit_version": np.array(rdkit.__version__), "smiles": np.array([Chem.MolToSmiles(m) for m in mols])}

    # ---- Morgan / Tanimoto / Butina
    g = from_rdkit(mols)
    for k in ("atom_starts", "bond_starts", "atom_inv", "bond_inv", "bond_a", "bond_b"):
        out[f"graph_{k}"] = getattr(g, k)
    fps = None
    for radius in range(4):
        gen = rdFingerprintGenerator.GetMorganGenerator(radius=radius, fpSize=2048)
        bvs = [gen.GetFingerprint(m) for m in mols]
        bits = np.zeros((len(mols), 2048), dtype=bool)
        for i, bv in enumerate(bvs):
            bits[i, list(bv.GetOnBits())] = True
        out[f"morgan_bits_r{radius}"] = np.packbits(bits, axis=1, bitorder="little").view(np.uint32)
        if radius == 2:
            fps = bvs
    n = len(mols)
    sim = np.array([DataStructs.BulkTanimotoSimilarity(fps[i], fps) for i in range(n)])
    out["tanimoto_r2"] = sim
    dists = [1.0 - sim[i, j] for i in range(1, n) for j in range(i)]  # RDKit's lower-triangle order
    for cutoff in (0.3, 0.6):
        clusters = Butina.ClusterData(dists, n, cutoff, isDistData=True, reordering=True)
        ids = np.full(n, -1, dtype=np.int32)
        for c, members in enumerate(clusters):
            ids[list(members)] = c
        key = str(cutoff).replace(".", "p")
        out[f"butina_ids_{key}"] = ids
        out[f"butina_centroids_{key}"] = np.array([c[0] for c in clusters], dtype=np.int32)

    # ---- conformers, MMFF / UFF
    hmols = []
    for m in mols:
        mh = Chem.AddHs(m)
        p = rdDistGeom.ETKDGv3()
        p.randomSeed = 42
        p.useRandomCoords = True
        if (len(rdDistGeom.EmbedMultipleConfs(mh, args.confs, p)) == args.confs and FFH.MMFFHasAllMoleculeParams(mh)
                and FFH.UFFHasAllMoleculeParams(mh)):
            hmols.append(mh)
    out["ff_smiles"] = np.array([Chem.MolToSmiles(m) for m in hmols])
    for kind, flatten in (("mmff", A.mmff_from_rdkit), ("uff", A.uff_from_rdkit)):
        flat = flatten(hmols)
        out[f"{kind}_atom_counts"] = flat.system.atom_counts
        for name, (st, ix, pr) in flat.system.tables.items():
            out[f"{kind}_{name}_starts"], out[f"{kind}_{name}_idx"], out[f"{kind}_{name}_par"] = st, ix, pr
        out[f"{kind}_conf_mol"], out[f"{kind}_atom_starts"], out[f"{kind}_positions"] = flat.batch.conf_mol, flat.batch.atom_starts, flat.batch.positions
        e0, grads, e_min = [], [], []
        for m in hmols:
            props = FFH.MMFFGetMoleculeProperties(m) if kind == "mmff" else None
            for conf in m.GetConformers():
                ff = (FFH.MMFFGetMoleculeForceField(m, props, confId=conf.GetId()) if kind == "mmff"
                      else FFH.UFFGetMoleculeForceField(m, confId=conf.GetId()))
                e0.append(ff.CalcEnergy())
                grads.append(np.array(ff.CalcGrad()).reshape(-1, 3))
            work = Chem.Mol(m)
            res = (FFH.MMFFOptimizeMoleculeConfs(work, maxIters=200) if kind == "mmff"
                   else FFH.UFFOptimizeMoleculeConfs(work, maxIters=1000))
            e_min += [[conv, e] for conv, e in res]
        out[f"{kind}_rdkit_energy"] = np.array(e0)
        out[f"{kind}_rdkit_grad"] = np.concatenate(grads)
        out[f"{kind}_rdkit_minimised"] = np.array(e_min)  # [not-converged flag, energy] per conformer

    # ---- ETKDG inputs and RDKit's own embedding of the same molecules
    params = EmbedParameters()
    raw, smooth = [], []
    for m in hmols:
        raw.append(np.asarray(rdDistGeom.GetMoleculeBoundsMatrix(m, doTriangleSmoothing=False), dtype=np.float64).ravel())
        smooth.append(np.asarray(rdDistGeom.GetMoleculeBoundsMatrix(m, doTriangleSmoothing=True), dtype=np.float64).ravel())
    out["bounds_raw"], out["bounds_smoothed"] = np.concatenate(raw), np.concatenate(smooth)
    out["bounds_starts"] = np.concatenate([[0], np.cumsum([len(b) for b in raw])])
    try:  # the adapter smooths on the GPU; substitute RDKit's own smoothing so the export needs no GPU
        import nvmolkit_b200.dgprep as dgprep

        def rdkit_smooth(mats, tol=0.0):
            done = [np.asarray(rdDistGeom.GetMoleculeBoundsMatrix(m, doTriangleSmoothing=True)) for m in hmols]
            return done, [True] * len(done)

        dgprep.triangle_smooth = rdkit_smooth
        flat = A.embed_molecules_from_rdkit(hmols, params)
        for sysname, system in (("dg", flat.dg), ("etk", flat.etk)):
            for name, (st, ix, pr) in system.tables.items():
                out[f"{sysname}_{name}_starts"], out[f"{sysname}_{name}_idx"], out[f"{sysname}_{name}_par"] = st, ix, pr
        for name, (st, ix, pr) in flat.checks.tables.items():
            out[f"check_{name}_starts"], out[f"check_{name}_idx"], out[f"check_{name}_par"] = st, ix, pr
        out["check_num_impropers"] = flat.checks.num_impropers
    except Exception as e:  # keep the rest of the export usable
        out["etkdg_export_error"] = np.array(repr(e))
    out["rdkit_embedded_positions"] = np.concatenate([c.GetPositions() for m in hmols for c in m.GetConformers()])
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {len(mols)} molecules, {len(hmols)} with conformers, RDKit {rdkit.__version__}")


if __name__ == "__main__":
    main()
