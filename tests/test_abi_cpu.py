"""The C-ABI library loads without a GPU and exports every symbol include/lsps_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "lsps_b200.h")) as fh:
        txt = fh.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lsps_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from lsps_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert lib.lsps_abi_version() == 1


def test_python_binding_covers_the_header():
    from lsps_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_ctx_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from lsps_b200 import _lib
    import pytest
    with pytest.raises(_lib.LspsError):
        _lib.Context(0)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    ctx = ctypes.c_void_p()
    assert lib.lsps_ctx_create(ctypes.byref(ctx), 0) != 0 and not ctx.value
