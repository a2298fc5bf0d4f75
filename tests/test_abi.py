"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/b200_clover.h declares; the ctypes binding table matches the header; calls without a device fail loudly
(there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_clover.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from chroma_b200 import build
    return build.build()


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("b200_create", "b200_destroy", "b200_load_gauge", "b200_load_clover", "b200_make_clover", "b200_dslash",
                 "b200_clover_apply", "b200_clover_matpc", "b200_invert", "b200_qprop", "b200_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in declared_symbols():
        assert hasattr(lib, name), "libb200clover.so does not export %s" % name


def test_ctypes_table_matches_header(built_lib):
    from chroma_b200 import lib as L
    assert sorted(L.SYMBOLS) == declared_symbols()
    L.load()


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the engine refuses to construct (and says why) instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from chroma_b200 import lib as L
    from chroma_b200.solver import Context
    with pytest.raises(L.B200Error) as e:
        Context((4, 4, 4, 4))
    assert e.value.code == L.B200_ERR_CUDA
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_create_validates_the_lattice_before_touching_a_device(built_lib):
    """Host-side argument checks of b200_create (engine_impl.cuh::init) come before the device probe: odd extents, a process grid
    that splits x or y, and a local lattice too big for the kernels' division-free site decode (2^28 sites per checkerboard)."""
    from chroma_b200 import lib as L
    from chroma_b200.solver import Context
    for latt, what in (((4, 4, 4, 6 + 1), "even"), ((128, 128, 128, 256), "2^28")):
        with pytest.raises(L.B200Error) as e:
            Context(latt)
        assert e.value.code == L.B200_ERR_ARG and what in str(e.value), str(e.value)


def test_product_never_imports_the_oracle():
    """Nothing under chroma_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "chroma_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/oracle.py", "").lower() or f == "fields.py", (dirpath, f)
