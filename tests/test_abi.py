"""CPU: the C-ABI library loads and exports every symbol include/trgt_denovo_b200.h declares."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry


@pytest.fixture(scope="module")
def built():
    entry.build()
    return entry.LIB


def _declared_symbols():
    hdr = open(os.path.join(entry.ROOT, "include", "trgt_denovo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tdb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_binding_matches_header(built):
    import trgt_denovo_b200 as tdb
    from importlib import import_module
    B = import_module("trgt_denovo_b200.binding")
    assert sorted(B.EXPORTS) == _declared_symbols()
    assert tdb.lib().tdb_abi_version() == 5
    assert ctypes.sizeof(B.AlleleOut) == 96 and ctypes.sizeof(B.TrioLocus) == 84


def test_no_cpu_fallback(built):
    """Without a CUDA device the product refuses to run instead of falling back to a CPU path."""
    import trgt_denovo_b200 as tdb
    if tdb.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(tdb.TdbError) as e:
        tdb.Context()
    assert e.value.code == -2


def test_product_does_not_import_oracle():
    pkg = os.path.join(entry.ROOT, "trgt-denovo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references oracle/"


def test_headers_are_plain_c():
    """The drop-in boundary is a C ABI: both public headers must compile as C99 on their own (no C++-isms, no
    torch or CUDA types in any signature)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for h in ("trgt_denovo_b200.h", "tdb_synth.h"):
        subprocess.check_call(["gcc", "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", "-I",
                               os.path.join(root, "include"), os.path.join(root, "include", h)])
    text = open(os.path.join(root, "include", "trgt_denovo_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # declarations only: the comments may name what is absent
    assert "torch" not in code and "cudaStream" not in code and "at::" not in code and "#include <cuda" not in code
