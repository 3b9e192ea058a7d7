"""Build a -DB200_BFGS_TIMING copy of the library, run the path-B workload, print the clock breakdown."""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, "/root/repo")
src = "/root/repo/nvmolkit_b200/csrc"
out = "/tmp/libb200mol_timing.so"
srcs = [f for f in os.listdir(src) if f.endswith(".cu")]
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DB200_BFGS_TIMING", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-shared", "-o", out] + [os.path.join(src, f) for f in srcs] + ["-lcudart_static"], check=True)
from nvmolkit_b200 import _lib
_lib.LIB_PATH = out
import bench, torch
_lib.profile_enable(True)
flat, mmff = bench.path_b_pool(64, 20260924)
dev = torch.device("cuda", 0)
import numpy as np
L = _lib.load()
buf = (C.c_ulonglong * 8)()
for name in ("etkdg", "bfgs"):
    getattr(L, f"b200mol_debug_clocks_{name}")(buf)  # reset
import sys
if len(sys.argv) > 1: _lib.set_option('bfgs_ctas_per_sm', int(sys.argv[1]))
r = bench.run_path_b_gpu(flat, mmff, 256, 10, dev, 1, 0)
print({k: r[k] for k in ("mols_per_s", "phases_ms", "mean_attempts")})
labels = ["energy evals", "gradient evals", "H*dGrad pass", "H update+dir pass", "iterations", "total in bfgsMinimize"]
for name in ("etkdg", "bfgs"):
    getattr(L, f"b200mol_debug_clocks_{name}")(buf)
    v = np.array(list(buf), dtype=np.float64)
    tot = v[5]
    print(name, "iterations", int(v[4]), "clk/iter", tot / max(v[4], 1))
    for k in (0, 1, 2, 3):
        print(f"   {labels[k]:22s} {100 * v[k] / tot:5.1f} %   {v[k] / max(v[4],1):9.0f} clk/iter")
    print(f"   {'everything else':22s} {100 * (tot - v[:4].sum()) / tot:5.1f} %   {(tot - v[:4].sum()) / max(v[4],1):9.0f} clk/iter")
