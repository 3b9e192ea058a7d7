import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
import oracle
from nvmolkit_b200 import synthetic as S
from nvmolkit_b200.forcefield import ConformerBatch
from nvmolkit_b200.minimizer import minimize
system, xyz, _ = S.random_uff_system(8, 4, 30, seed=31)
rng = np.random.default_rng(2)
batch = ConformerBatch.from_coords(system, [[x + rng.normal(0, 0.05, x.shape), x + rng.normal(0, 0.15, x.shape)] for x in xyz])
pos0, _, _, _ = oracle.ff_minimize("uff", system.atom_counts, system.tables, np.arange(8, dtype=np.int32),
                                   np.concatenate([[0], np.cumsum(system.atom_counts)]).astype(np.int32), np.concatenate(xyz), 2000, 1e-4)
st0 = np.concatenate([[0], np.cumsum(system.atom_counts)])
relaxed = [pos0[st0[m]:st0[m + 1]] for m in range(8)]
b2 = ConformerBatch.from_coords(system, [[r + rng.normal(0, 0.05, r.shape) for _ in range(2)] for r in relaxed])
pos_o, e_o, conv_o, it_o = oracle.ff_minimize("uff", system.atom_counts, system.tables, b2.conf_mol, b2.atom_starts, b2.positions, 1000, 1e-4)
for rep in range(8):
    res = minimize(system, b2, 1000, 1e-4)
    eg, st = res.energies.cpu().numpy(), res.status.cpu().numpy()
    np.set_printoptions(precision=2, linewidth=200); print("rel", np.abs(eg - e_o) / np.maximum(1, np.abs(e_o)))
    print("st", st, "conv_o", conv_o)
print(e_o)
