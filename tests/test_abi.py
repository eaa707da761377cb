"""The C-ABI library loads and exports every symbol include/tacult_b200.h declares (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tacult_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_library):
    names = declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(built_library)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_covers_header():
    from tacult_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()
    lib = _lib.load()
    assert lib.tc_abi_version() == _lib.ABI_VERSION == 2


def test_library_is_sm100a(built_library):
    out = subprocess.run(["cuobjdump", "-lelf", built_library], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_header():
    from tacult_b200 import _lib
    assert _lib.EXAMPLE_DTYPE.itemsize == 216
    assert ctypes.sizeof(_lib.Config) == 40
    assert ctypes.sizeof(_lib.PlayoutResult) == 48
    assert ctypes.sizeof(_lib.SelfPlayConfig) == 32
    assert ctypes.sizeof(_lib.GameStateStruct) == _lib.GAMESTATE_DTYPE.itemsize == 44


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tacult_b200 import Engine, _lib
    with pytest.raises(_lib.TacultError):
        Engine(4, 64)
