import sys, os, subprocess
for name, ctas in [(x.split(":")[0], int(x.split(":")[1])) for x in sys.argv[1:]]:
    code = f"""
import sys; sys.path.insert(0, '/root/repo')
from nvmolkit_b200 import _lib
_lib.LIB_PATH = '/root/repo/scratch/variants/libb200mol_{name}.so'
import bench, torch
_lib.profile_enable(True)
_lib.set_option('bfgs_ctas_per_sm', {ctas})
flat, mmff = bench.path_b_pool(64, 20260924)
r = bench.run_path_b_gpu(flat, mmff, 1024, 10, torch.device('cuda', 0), 1, 1)
print('variant {name} ctas', {ctas}, round(r['mols_per_s'], 1), {{k: round(v) for k, v in r['phases_ms'].items()}})
"""
    subprocess.run([sys.executable, "-c", code])
