"""CPU: the C-ABI library loads without a GPU and exports every symbol include/imrcd.h declares; the Python binding's
struct layouts match the header; creating a context without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "imrcd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(imrcd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from inmyroom_vulkan_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/imrcd.h but not exported by libimrcd.so"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "the ctypes binding and the header disagree"
    assert b"sm_100a" in lib.imrcd_version()


def test_struct_layouts_match_header():
    from inmyroom_vulkan_b200 import _lib
    from inmyroom_vulkan_b200.collision import HIT_DTYPE, PAIR_DTYPE
    assert C.sizeof(_lib.EntityPair) == 80 == PAIR_DTYPE.itemsize
    assert C.sizeof(_lib.TriHit) == 40 == HIT_DTYPE.itemsize
    for (name, _), f in zip(_lib.EntityPair._fields_, PAIR_DTYPE.names):
        assert name == f and getattr(_lib.EntityPair, name).offset == PAIR_DTYPE.fields[f][1]


def test_frame_stats_layout_matches_the_library():
    """The ctypes mirror of imrcd_frame_stats against sizeof / offsetof as libimrcd.so was compiled (imrcd_abi_layout)."""
    from inmyroom_vulkan_b200 import _lib
    lib = _lib.load()
    n = lib.imrcd_abi_layout(None, 0)
    out = (C.c_uint64 * n)()
    assert lib.imrcd_abi_layout(out, n) == n
    v = list(out)
    assert v[0] == C.sizeof(_lib.EntityPair) and v[1] == C.sizeof(_lib.TriHit) and v[2] == C.sizeof(_lib.FrameStats)
    fields = [name for name, _ in _lib.FrameStats._fields_]
    assert len(fields) == n - 3, "imrcd_frame_stats has a field the binding does not know (or the other way round)"
    for name, off in zip(fields, v[3:]):
        assert getattr(_lib.FrameStats, name).offset == off, name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from inmyroom_vulkan_b200.collision import Context, ImrcdError
    with pytest.raises(ImrcdError, match="no sm_100 device"):
        Context(0)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "inmyroom_vulkan_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt and "libimr_ref" not in txt, fn


def test_header_is_plain_c(tmp_path):
    """include/imrcd.h is the drop-in boundary: it has to compile as C99 (plain pointers and sizes, no C++), warnings on."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "imrcd.h"\nint main(void) { imrcd_ctx* c = 0; (void)c; return (int)sizeof(imrcd_entity_pair) - 80; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
