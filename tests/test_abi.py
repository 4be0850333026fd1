"""CPU tests of the boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/autoinst_ncuts.h declares.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from autoinst_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "autoinst_ncuts.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ancuts_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.ancuts_version() >= 100


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.Params) == 6 * 8 + 4 * 4 + 8 + 8      # 6 doubles, 4 ints, double, 2 ints
    assert ctypes.sizeof(_lib.NodeStat) == 8 * 4 + 2 * 8


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.ancuts_create(0, ctypes.byref(h))
    assert rc != 0
    assert b"no CPU fallback" in lib.ancuts_last_error()
    from autoinst_b200 import api
    with pytest.raises(RuntimeError):
        api.Handle.get(0)


def test_workspace_planning(lib):
    import numpy as np
    from autoinst_b200 import api
    one = api.workspace_bytes([8192])
    two = api.workspace_bytes([8192, 8192])
    assert 2 * 8192 * 8192 * 4 < one < two < 2.2 * one
    batches = api.plan_batches([8192] * 5, budget_bytes=int(2.5 * one))
    assert sorted(sum(batches, [])) == list(range(5)) and len(batches) >= 2


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under autoinst_b200/ or ncuts/ may reference it."""
    for pkg in ("autoinst_b200", "ncuts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
