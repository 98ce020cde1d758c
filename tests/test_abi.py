"""The C-ABI library loads and exports every symbol include/jqmc_b200.h declares (no compute calls)."""

import ctypes
import os
import re

import pytest

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "jqmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qe_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as g

    g.build()
    from jqmc_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jqmc_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert _lib.load().qe_version() >= 100


def test_missing_library_fails_loudly(monkeypatch):
    from jqmc_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libjqmc_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_engine_requires_cuda(water):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jqmc_b200.engine import WalkerEngine

    with pytest.raises(RuntimeError, match="CUDA"):
        WalkerEngine(water)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jqmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_build_is_deterministic_by_construction():
    """The build must not use `nvcc --split-compile`: it produced different machine code (and different speed) from identical
    sources on every run (profiles/r02_walker_history.md).  The large kernel families are split over translation units instead."""
    import __graft_entry__ as g

    assert not any("split-compile" in f for f in g.NVCC_FLAGS)
    src = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert '"--split-compile"' not in src
    csrc = os.path.join(ROOT, "jqmc_b200", "csrc")
    units = [f for f in os.listdir(csrc) if f.endswith(".cu")]
    assert sum(f.startswith("qe_walker_i_") for f in units) == 12 and sum(f.startswith("qe_mcmc_i_") for f in units) == 3
