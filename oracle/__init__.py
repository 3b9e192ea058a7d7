"""CPU oracle — TEST INFRASTRUCTURE ONLY.

ctypes front-end to ``oracle/liboracle.so`` (plain-C restatements of the reference algorithms, see the header of each
``oracle_*.c``). Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this
package; nothing under ``nvmolkit_b200/`` does.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no GPU involved). Returns the library path."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.startswith("oracle_") and f.endswith(".c")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _declare(L: C.CDLL) -> None:
    u32p, i32p, f64p, u16p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint16)
    L.oracle_similarity_cross.argtypes = [u32p, C.c_long, u32p, C.c_long, C.c_int, C.c_int, f64p]
    L.oracle_similarity_cross.restype = None
    L.oracle_count_ge.argtypes = [u32p, C.c_long, u32p, C.c_long, C.c_int, C.c_int, C.c_double, C.c_int, i32p]
    L.oracle_count_ge.restype = None
    L.oracle_butina_dense.argtypes = [f64p, C.c_long, C.c_double, i32p, i32p]
    L.oracle_butina_dense.restype = C.c_int
    L.oracle_butina_fp.argtypes = [u32p, C.c_long, C.c_int, C.c_int, C.c_double, i32p, i32p]
    L.oracle_butina_fp.restype = C.c_int
    L.oracle_morgan_atom_invariant.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_int]
    L.oracle_morgan_atom_invariant.restype = C.c_uint32
    L.oracle_morgan_one.argtypes = [C.c_int, C.c_int, u32p, u32p, u16p, u16p, C.c_int, C.c_int, u32p, u32p]
    L.oracle_morgan_one.restype = C.c_int
    L.oracle_morgan.argtypes = [i32p, i32p, u32p, u32p, u16p, u16p, C.c_long, C.c_int, C.c_int, u32p]
    L.oracle_morgan.restype = None
    for name, fn in _LATE_DECL.items():
        if hasattr(L, name):
            fn(getattr(L, name))


_LATE_DECL: dict = {}

METRIC = {"tanimoto": 0, "cosine": 1}


def _fp(a) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a).view(np.uint32) if np.asarray(a).dtype == np.int32 else a, dtype=np.uint32)
    assert a.ndim == 2
    return a


# ------------------------------------------------------------------ path A
def similarity_cross(a, b=None, metric: str = "tanimoto") -> np.ndarray:
    a = _fp(a)
    b = a if b is None else _fp(b)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    lib().oracle_similarity_cross(_p(a, C.c_uint32), a.shape[0], _p(b, C.c_uint32), b.shape[0], a.shape[1],
                                  METRIC[metric], _p(out, C.c_double))
    return out


def count_ge(x, y, cutoff: float, metric: str = "tanimoto", sign: int = 1, counts=None) -> np.ndarray:
    x, y = _fp(x), _fp(y)
    if counts is None:
        counts = np.zeros(x.shape[0], dtype=np.int32)
    lib().oracle_count_ge(_p(x, C.c_uint32), x.shape[0], _p(y, C.c_uint32), y.shape[0], x.shape[1], METRIC[metric],
                          float(cutoff), int(sign), _p(counts, C.c_int32))
    return counts


def butina_dense(dist, cutoff: float):
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    n = dist.shape[0]
    ids = np.empty(n, dtype=np.int32)
    cen = np.empty(max(n, 1), dtype=np.int32)
    k = lib().oracle_butina_dense(_p(dist, C.c_double), n, float(cutoff), _p(ids, C.c_int32), _p(cen, C.c_int32))
    return ids, cen[:k].copy()


def butina_fp(fp, cutoff: float, metric: str = "tanimoto"):
    fp = _fp(fp)
    n = fp.shape[0]
    ids = np.empty(n, dtype=np.int32)
    cen = np.empty(max(n, 1), dtype=np.int32)
    k = lib().oracle_butina_fp(_p(fp, C.c_uint32), n, fp.shape[1], METRIC[metric], float(cutoff), _p(ids, C.c_int32),
                               _p(cen, C.c_int32))
    return ids, cen[:k].copy()


def morgan_atom_invariant(z, total_degree, total_hs, charge=0, delta_mass=0, in_ring=False) -> int:
    return int(lib().oracle_morgan_atom_invariant(z, total_degree, total_hs, charge, delta_mass, int(bool(in_ring))))


def morgan_codes(atom_inv, bond_inv, bond_a, bond_b, radius: int) -> np.ndarray:
    """Unfolded environment codes of one molecule (what RDKit's sparse fingerprint counts)."""
    atom_inv = np.ascontiguousarray(atom_inv, dtype=np.uint32)
    bond_inv = np.ascontiguousarray(bond_inv, dtype=np.uint32)
    bond_a = np.ascontiguousarray(bond_a, dtype=np.uint16)
    bond_b = np.ascontiguousarray(bond_b, dtype=np.uint16)
    codes = np.zeros((radius + 1) * max(len(atom_inv), 1), dtype=np.uint32)
    n = lib().oracle_morgan_one(len(atom_inv), len(bond_inv), _p(atom_inv, C.c_uint32), _p(bond_inv, C.c_uint32),
                                _p(bond_a, C.c_uint16), _p(bond_b, C.c_uint16), radius, 2048, None,
                                _p(codes, C.c_uint32))
    return codes[:n].copy()


def morgan(atom_starts, bond_starts, atom_inv, bond_inv, bond_a, bond_b, radius: int, fp_bits: int) -> np.ndarray:
    atom_starts = np.ascontiguousarray(atom_starts, dtype=np.int32)
    bond_starts = np.ascontiguousarray(bond_starts, dtype=np.int32)
    atom_inv = np.ascontiguousarray(atom_inv, dtype=np.uint32)
    bond_inv = np.ascontiguousarray(bond_inv, dtype=np.uint32)
    bond_a = np.ascontiguousarray(bond_a, dtype=np.uint16)
    bond_b = np.ascontiguousarray(bond_b, dtype=np.uint16)
    n = len(atom_starts) - 1
    out = np.zeros((n, fp_bits // 32), dtype=np.uint32)
    lib().oracle_morgan(_p(atom_starts, C.c_int32), _p(bond_starts, C.c_int32), _p(atom_inv, C.c_uint32),
                        _p(bond_inv, C.c_uint32), _p(bond_a, C.c_uint16), _p(bond_b, C.c_uint16), n, radius, fp_bits,
                        _p(out, C.c_uint32))
    return out
