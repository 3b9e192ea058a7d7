import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from nvmolkit_b200 import _lib
_lib.LIB_PATH = "/root/repo/scratch/variants/libb200mol_tct.so"
from nvmolkit_b200 import synthetic as S
from nvmolkit_b200.clustering import fused_butina_device
mode = int(sys.argv[1])
_lib.set_option("similarity_tensor_cluster", mode)
fp = S.clustered_fingerprints(6000, 50, seed=S.SEED)  # 300k x 300k
d = torch.from_numpy(fp.view(np.int32)).cuda()
for _ in range(2):
    ids, cen = fused_butina_device(d, 0.3)
    torch.cuda.synchronize()
print("mode", mode, "clusters", int(cen.numel()))
