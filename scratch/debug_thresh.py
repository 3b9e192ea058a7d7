import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
import oracle
from nvmolkit_b200 import _lib, synthetic as S
from nvmolkit_b200.similarity import crossTanimotoSimilarity
rng = np.random.default_rng(0)
rows = []
for u, c in [(10, 7), (20, 14), (100, 70), (1000, 700), (10, 6), (3, 2), (7, 5)]:
    a = np.zeros(2048, dtype=bool); b = np.zeros(2048, dtype=bool)
    perm = rng.permutation(2048); a[perm[:u]] = True; b[perm[:c]] = True
    rows.append((S.pack_bits(a[None])[0], S.pack_bits(b[None])[0]))
x = np.stack([r[0] for r in rows]); y = np.stack([r[1] for r in rows])
dx = torch.from_numpy(x.view(np.int32)).cuda(); dy = torch.from_numpy(y.view(np.int32)).cuda()
sim = crossTanimotoSimilarity(dx, dy).numpy()
print("diag sims", np.diag(sim).tolist())
print("oracle diag", np.diag(oracle.similarity_cross(x, y)).tolist())
print("1-sim<=0.3 numpy:", ((1.0 - sim) <= 0.3).sum(1))
for cutoff in (0.29, 0.3, 0.30000000000000004, 0.31):
    counts = torch.zeros(7, dtype=torch.int32, device="cuda")
    _lib.call("b200mol_tanimoto_count_ge", dx.data_ptr(), 7, dy.data_ptr(), 7, 64, 0, cutoff, 1, counts.data_ptr(), torch.cuda.current_stream().cuda_stream)
    print(cutoff, counts.cpu().numpy(), oracle.count_ge(x, y, cutoff))
