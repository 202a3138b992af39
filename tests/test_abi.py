"""The C-ABI library loads and exports every symbol include/b200gym.h declares; the ctypes mirror matches the C
struct layouts; without a GPU the library fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from isaacgymenv_b200 import _abi, _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200gym.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2g_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(_lib.LIB_PATH), "build the library first: python __graft_entry__.py"
    lib = C.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in b200gym.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) <= set(declared)


def test_struct_layouts_match():
    lib = _lib.load()      # load() itself verifies ABI version and sizeof() of every public POD
    assert lib.b2g_abi_version() == _abi.B2G_ABI_VERSION
    assert lib.b2g_sizeof(0) == C.sizeof(_abi.Model)
    assert lib.b2g_sizeof(5) == C.sizeof(_abi.AnymalCfg)


def test_error_paths_without_compute():
    lib = _lib.load()
    assert lib.b2g_sim_destroy(None) == -1
    assert b"null" in lib.b2g_last_error()
    sim = C.c_void_p()
    assert lib.b2g_sim_create(0, None, C.byref(sim)) == -1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is for CPU-only hosts")
    import isaacgymenv_b200

    with pytest.raises(_lib.B2GError, match="no usable CUDA device|no CPU"):
        isaacgymenv_b200.make(42, "Anymal", 16, "cuda:0", "cuda:0")
    with pytest.raises(_lib.B2GError, match="no CPU simulation path"):
        isaacgymenv_b200.make(42, "Anymal", 16, "cpu", "cpu")


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (or the host emulator)."""
    pkg = os.path.join(ROOT, "isaacgymenv_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, flags=re.M), f"{f} imports test infrastructure"
                assert "oracle_dyn" not in src and "liboracle" not in src and "libb2g_emu" not in src
