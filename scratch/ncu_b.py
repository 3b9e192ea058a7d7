"""One ETKDG + MMFF step on a small batch, for `ncu -k regex:etkdgKernel|bfgsKernel`."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 60
flat, mmff = bench.path_b_pool(64, 20260924)
r = bench.run_path_b_gpu(flat, mmff, nm, 10, torch.device("cuda", 0), 1, 0)
print({k: r[k] for k in ('mols_per_s', 'ms_per_step', 'phases_ms')})
