"""Run bench.py against an experimental build of the library (A/B of compile-time tile parameters):
    python tools/bench_with_lib.py nvmolkit_b200/lib/libb200mol_n192.so --workload butina --steps 5 ...
The library is bound through ctypes (the pybind module belongs to the main build)."""
import os
import runpy
import sys

os.environ["B200_NO_CORE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvmolkit_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
