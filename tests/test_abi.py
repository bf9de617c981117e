"""CPU: the C-ABI library loads without a GPU and exports every symbol include/straps_b200.h declares."""
import ctypes
import os
import re

from conftest import REPO
from straps_b200 import _lib


def _declared():
    src = open(os.path.join(REPO, 'include', 'straps_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(straps_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 18
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_version_and_error_string():
    L = _lib.lib()
    assert L.straps_abi_version() == 3
    assert isinstance(L.straps_last_error(), bytes)
    assert L.straps_launch_count() >= 0
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert L.straps_rot6d_to_rotmat(None, 4, None, None) != 0
    assert b'null' in L.straps_last_error()
    h = ctypes.c_void_p()
    assert L.straps_regressor_create(ctypes.byref(h), 99, 4) != 0
    assert b'c_in' in L.straps_last_error()
