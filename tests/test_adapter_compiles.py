"""The Chroma adapter (chroma_adapter/) cannot be linked here -- QDP++/QMP/libxml2 are absent -- but it can be
type-checked: compile every adapter source against tests/mock_chroma/, a minimal stand-in for the slice of the
QDP++/Chroma API it touches, and against the real include/b200_clover.h.  This catches
misspelt ABI calls, wrong argument lists and C++ errors; it says nothing about QDP++ semantics."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "chroma_adapter")
SOURCES = ["syssolver_b200_clover_params.cc", "syssolver_linop_clover_b200_w.cc", "syssolver_mdagm_clover_b200_w.cc",
           "multi_syssolver_mdagm_clover_b200_w.cc"]


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    """Lay the adapter out where Chroma's include paths expect it: lib/actions/ferm/invert/b200_solvers/."""
    top = tmp_path_factory.mktemp("chroma_lib")
    dst = top / "actions" / "ferm" / "invert" / "b200_solvers"
    dst.mkdir(parents=True)
    for f in os.listdir(ADAPTER):
        if f.endswith((".h", ".cc")):
            shutil.copy(os.path.join(ADAPTER, f), dst / f)
    return top


@pytest.mark.parametrize("src", SOURCES)
def test_adapter_source_type_checks(tree, src):
    cmd = ["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Wno-unused",
           "-I", str(tree), "-I", os.path.join(ROOT, "tests", "mock_chroma"), "-I", os.path.join(ROOT, "include"),
           str(tree / "actions" / "ferm" / "invert" / "b200_solvers" / src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_adapter_calls_only_declared_abi():
    """Every b200_* identifier the adapter uses is declared in include/b200_clover.h."""
    import re
    hdr = open(os.path.join(ROOT, "include", "b200_clover.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)) | set(re.findall(r"\b(b200_[a-z_]+)\b(?=\s*[;{*])", hdr))
    declared |= {"b200_ctx", "b200_comm", "b200_solve_info", "b200_field", "b200_clover", "b200_solvers"}
    used = set()
    for f in os.listdir(ADAPTER):
        if f.endswith((".h", ".cc")):
            used |= set(re.findall(r"\b(b200_[a-z0-9_]+)\b", open(os.path.join(ADAPTER, f)).read()))
    unknown = {u for u in used if u not in declared and not u.endswith("_w") and not u.endswith("_params") and not u.endswith("_engine")}
    assert not unknown, unknown
