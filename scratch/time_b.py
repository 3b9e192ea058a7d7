import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from nvmolkit_b200 import _lib
_lib.profile_enable(True)
pool = int(sys.argv[1]); nm = int(sys.argv[2]); confs = int(sys.argv[3])
if len(sys.argv) > 4: bench.MAX_ATTEMPTS = int(sys.argv[4])
t = time.time(); flat, mmff = bench.path_b_pool(pool, 20260924); print("pool gen", time.time() - t, "atoms", flat.atom_counts.mean())
dev = torch.device("cuda", 0)
import os
for c in [int(x) for x in os.environ.get("CTAS", "2,3,4").split(",")]:
    _lib.set_option('bfgs_ctas_per_sm', c)
    r = bench.run_path_b_gpu(flat, mmff, nm, confs, dev, 1, 1)
    print(c, {k: r[k] for k in ('mols_per_s','ms_per_step','phases_ms','conformers_embedded_frac','mean_attempts')})
