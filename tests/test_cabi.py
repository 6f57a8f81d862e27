"""The C-ABI shared library loads without a GPU and exports exactly what include/l2d_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions(name="l2d_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(l2d_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from live2diff_b200 import _lib

    assert header_functions() == sorted(_lib.SIGNATURES)
    assert header_functions("l2d_b200_debug.h") == sorted(_lib.DEBUG_SIGNATURES)      # developer hooks: separate header
    assert not set(_lib.DEBUG_SIGNATURES) & set(header_functions())


def test_library_exports_every_symbol():
    from live2diff_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions() + header_functions("l2d_b200_debug.h"):
        assert hasattr(handle, name), f"libl2d_b200.so does not export {name}"
    lib = _lib.lib()
    assert lib.l2d_abi_version() == _lib.ABI_VERSION == 3


def test_library_carries_the_hash_of_the_sources_next_to_it():
    """A stale .so (sources changed, library not rebuilt) must be refused, not silently used."""
    from live2diff_b200 import _lib
    from live2diff_b200.csrc import build

    lib = _lib.lib()
    assert lib.l2d_build_hash().decode() == build.source_hash() == build.built_hash()
    assert lib.l2d_launch_count() >= 0


def test_argument_errors_are_reported_without_a_gpu():
    from live2diff_b200 import _lib

    lib = _lib.lib()
    rc = lib.l2d_kv_attn(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 16, 16, 64, 8, 0)
    assert rc == -1 and b"null pointer" in lib.l2d_last_error()
    rc = lib.l2d_gemm(16, 8, 16, 16, 8, 4, 12, 8, 0, 0, 0, 0, 0, 0, 0)       # N % 8 != 0
    assert rc == -1 and b"multiples of 8" in lib.l2d_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_product_has_no_oracle_dependency():
    """The shipped package must never import the oracle (it is the checker, not the product)."""
    pkg = os.path.join(ROOT, "live2diff_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
