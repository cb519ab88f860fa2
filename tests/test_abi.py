"""The C-ABI library loads and exports every symbol include/jutul_b200.h declares; the ctypes
prototype table matches the header; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "jutul_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(J):
    lib = J._lib.load()
    names = _header_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"libjutul_b200.so does not export {n}"


def test_prototype_table_matches_header(J):
    assert sorted(J._lib.PROTOTYPES) == _header_functions()
    assert J._lib.load().jb_version() == 100


def test_sm100a_cubin_present(J):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", J._lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback(J):
    with pytest.raises(J.JutulB200Error) as e:
        J.B200Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jutul.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "oracle.cpp" not in txt and "liborc" not in txt, f
