"""clock64() attribution of the pair-pass tile (MMA thread: waiting for operands / for a free accumulator; epilogue:
waiting for the accumulator / at its staging barrier). Uses the -DB200_TC_TIMING build (make -C nvmolkit_b200/csrc tctiming).
    python tools/pair_pass_timing.py [n_centres] [cluster variant ...]"""
import ctypes as C
import os
import sys

import numpy as np

os.environ["B200_NO_CORE"] = "1"  # bind the instrumented library through ctypes, not the pybind module of the main one

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvmolkit_b200 import _lib  # noqa: E402

# B200_TC_LIB=mmaonly: the library built by `make mmaonly` (MMA instruction stream alone; results are garbage)
_lib.LIB_PATH = os.path.join(ROOT, "nvmolkit_b200", "lib", f"libb200mol_{os.environ.get('B200_TC_LIB', 'tctiming')}.so")
import torch  # noqa: E402

from nvmolkit_b200 import synthetic  # noqa: E402
from nvmolkit_b200.clustering import fused_butina_device  # noqa: E402

n_centres = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
variants = [int(v) for v in sys.argv[2:]] or [1, 3]
fp = torch.from_numpy(synthetic.clustered_fingerprints(n_centres, 50).view(np.int32)).cuda()
L = _lib.load()
buf = (C.c_ulonglong * 8)()
_lib.profile_enable(True)
_lib.set_option("similarity_superpose_auto", 0)  # no pilot passes in the attribution
_lib.set_option("similarity_superpose_cols", int(os.environ.get("SUPER_COLS", "2")))
for v in variants:
    _lib.set_option("similarity_tensor_cluster", v)
    for _ in range(2):
        fused_butina_device(fp, 0.3)
    L.b200mol_debug_clocks_tc(buf)
    fused_butina_device(fp, 0.3)
    ms = _lib.profile_read("neighbor_pass_tc")
    L.b200mol_debug_clocks_tc(buf)
    c = np.array(list(buf), dtype=np.float64)
    tiles = max(c[6], 1)
    print(f"variant {v}: pass {ms:.2f} ms, tiles/CTA-thread {tiles:.0f}; per tile clocks: MMA thread {c[0] / tiles:.0f} "
          f"(operand wait {c[1] / tiles:.0f}, accumulator wait {c[2] / tiles:.0f}, issue {(c[0] - c[1] - c[2]) / tiles:.0f}); "
          f"epilogue warp {c[3] / tiles:.0f} (accumulator wait {c[4] / tiles:.0f}, staging barrier {c[5] / tiles:.0f}); "
          f"producer stage wait {c[7] / tiles:.0f}")
