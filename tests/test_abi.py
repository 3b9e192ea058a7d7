"""The C-ABI library loads here (no GPU) and exports every symbol include/b200mol.h declares."""

import os
import re

from nvmolkit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            names |= set(re.findall(r"\b(b200mol_[a-z0-9_]+)\s*\(", text))
    return names


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    declared = _declared()
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(built_lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SIGNATURES), sorted(declared ^ set(_lib.SIGNATURES))


def test_abi_version_and_error_string(built_lib):
    assert built_lib.b200mol_abi_version() >= 1
    assert isinstance(built_lib.b200mol_last_error(), bytes)


def test_no_cpu_fallback_without_device(built_lib):
    import torch

    if torch.cuda.is_available():
        return
    assert built_lib.b200mol_check_device(0) == _lib.ERR_NODEVICE
    assert b"no CPU fallback" in built_lib.b200mol_last_error() or b"sm_100a" in built_lib.b200mol_last_error()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "nvmolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert "liboracle" not in text, f
