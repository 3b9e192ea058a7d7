"""Per-phase clock breakdown of the conformer kernels: builds a -DB200_BFGS_TIMING copy of the library (clock64() around
the energy evaluations, the gradient evaluation and the inverse-Hessian sweep of every BFGS iteration), runs the
config-3 style workload on it and prints clocks per iteration.   python tools/path_b_timing.py [mols] [confs] [ctas/SM]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

os.environ["B200_NO_CORE"] = "1"  # bind the instrumented library through ctypes, not the pybind module of the main one

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.environ.get("B200_TIMING_LIB") or os.path.join(ROOT, "nvmolkit_b200", "lib", "libb200mol_timing.so")
if not os.environ.get("B200_TIMING_LIB"):
    subprocess.run(["make", "-C", os.path.join(ROOT, "nvmolkit_b200", "csrc"), "-j", "8", "-s", "timing"], check=True)
from nvmolkit_b200 import _lib  # noqa: E402

_lib.LIB_PATH = out
import torch  # noqa: E402

import bench  # noqa: E402

n_mols = int(sys.argv[1]) if len(sys.argv) > 1 else 512
confs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pool = bench.conformer_pool(n_mols, 20260924, 16)
torch.cuda.set_device(0)
_lib.profile_enable(True)
if len(sys.argv) > 3:
    _lib.set_option("bfgs_ctas_per_sm", int(sys.argv[3]))
if os.environ.get("B200_L2_PERSIST"):
    _lib.set_option("bfgs_l2_persist", 1)
L = _lib.load()
buf = (C.c_ulonglong * 8)()
leg = bench.ConformerLeg(pool[0], pool[1], torch.device("cuda", 0), 1, 0)
ids = np.arange(n_mols, dtype=np.int32)
leg.step(ids[:64], confs)
for name in ("etkdg", "bfgs"):
    getattr(L, f"b200mol_debug_clocks_{name}")(buf)  # reset
_lib.stats_read(reset=True)
ms, (raw, res) = bench._event_timed(lambda: leg.step(ids, confs), torch.device("cuda", 0), 1)
print({"mols_per_s": n_mols / (ms * 1e-3), "etkdg_ms": _lib.profile_read("etkdg"), "bfgs_ms": _lib.profile_read("bfgs"),
       "mean_attempts": float(raw.attempts.float().mean())})
print(_lib.stats_read())
labels = ["energy evals", "gradient evals", "Hessian sweep", "(unused)"]
for name in ("etkdg", "bfgs"):
    getattr(L, f"b200mol_debug_clocks_{name}")(buf)
    v = np.array(list(buf), dtype=np.float64)
    tot = v[5]
    print(name, "iterations", int(v[4]), "clk/iter", tot / max(v[4], 1))
    for k in (0, 1, 2):
        print(f"   {labels[k]:22s} {100 * v[k] / tot:5.1f} %   {v[k] / max(v[4], 1):9.0f} clk/iter")
    print(f"   {'everything else':22s} {100 * (tot - v[:3].sum()) / tot:5.1f} %   {(tot - v[:3].sum()) / max(v[4], 1):9.0f} clk/iter")
