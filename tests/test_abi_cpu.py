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


def test_struct_layouts_match_the_header(tmp_path):
    """lsps_conv_shape / lsps_conv_ext cross the boundary by pointer: the ctypes mirrors in lsps_b200/_lib.py must have
    the header's field order, offsets and size (checked against gcc's view of include/lsps_b200.h)."""
    import subprocess
    from lsps_b200 import _lib
    src = tmp_path / "layout.c"
    fields = [f for f, _ in _lib.ConvExt._fields_]
    sfields = [f for f, _ in _lib.ConvShape._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lsps_b200.h"', 'int main(void) {',
             '  printf("%zu %zu\\n", sizeof(lsps_conv_shape), sizeof(lsps_conv_ext));']
    lines += ['  printf("%%zu\\n", offsetof(lsps_conv_shape, %s));' % f for f in sfields]
    lines += ['  printf("%%zu\\n", offsetof(lsps_conv_ext, %s));' % f for f in fields]
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.ConvShape) and int(out[1]) == ctypes.sizeof(_lib.ConvExt)
    got = [int(v) for v in out[2:]]
    want = [getattr(_lib.ConvShape, f).offset for f in sfields] + [getattr(_lib.ConvExt, f).offset for f in fields]
    assert got == want
