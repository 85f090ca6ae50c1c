"""CPU-side checks of the boundary: the shared library loads and exports every symbol the
header declares; struct layouts agree between ctypes and the C compiler."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kx_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kx_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from kernex_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    names = _declared_functions()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/kx_b200.h but not exported"


def test_ctypes_signatures_cover_the_header(lib):
    from kernex_b200 import _lib
    assert sorted(_lib.exported_symbols()) == _declared_functions()
    assert lib.kx_abi_version() == _lib.KX_ABI_VERSION


def test_struct_layouts_match_the_c_compiler(tmp_path):
    from kernex_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "kx_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(KxGeom),sizeof(KxKeyItem),sizeof(KxKey),sizeof(KxFunc),sizeof(KxProgram),'
                    'offsetof(KxProgram,tap_off),offsetof(KxProgram,coef),sizeof(KxHaloSide),sizeof(KxHaloPull),'
                    'offsetof(KxHaloPull,counter),offsetof(KxHaloPull,side));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    want = [C.sizeof(_lib.KxGeom), C.sizeof(_lib.KxKeyItem), C.sizeof(_lib.KxKey), C.sizeof(_lib.KxFunc),
            C.sizeof(_lib.KxProgram), _lib.KxProgram.tap_off.offset, _lib.KxProgram.coef.offset,
            C.sizeof(_lib.KxHaloSide), C.sizeof(_lib.KxHaloPull), _lib.KxHaloPull.counter.offset, _lib.KxHaloPull.side.offset]
    assert [int(v) for v in out] == want


def test_validation_errors_do_not_need_a_gpu(lib):
    from kernex_b200 import _lib
    g, p = _lib.KxGeom(), _lib.KxProgram()
    g.ndim = 9
    rc = lib.kx_map(C.byref(g), C.byref(p), None, None, 0, None, 1, None)
    assert rc == 1 and b"ndim" in lib.kx_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from kernex_b200 import _lib
    from kernex_b200.errors import EngineUnavailableError
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(EngineUnavailableError):
        _lib.load()
