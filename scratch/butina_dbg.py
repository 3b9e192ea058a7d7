import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import oracle
from nvmolkit_b200 import _lib
import os
if os.environ.get("B200LIB"): _lib.LIB_PATH = os.environ["B200LIB"]
from nvmolkit_b200.clustering import butina
n, degree, mc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(n * 7 + degree)
d = rng.random((n, n)); d = np.minimum(d, d.T); np.fill_diagonal(d, 0.0)
cutoff = 1.0 - (1.0 - degree / n) ** 0.5
_lib.set_option("butina_min_round_commits", mc)
print("launch", flush=True)
ids, cen = butina(torch.from_numpy(d).to("cuda"), cutoff, return_centroids=True)
torch.cuda.synchronize()
print("done", flush=True)
ids_cpu, cen_cpu = oracle.butina_dense(d, cutoff)
print("match", (ids.numpy() == ids_cpu).all(), (cen.numpy() == cen_cpu).all(), len(cen_cpu))
