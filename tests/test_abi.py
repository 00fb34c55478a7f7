"""The C-ABI library loads and exports every symbol include/bg_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from building_gan_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "bg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    handle = lib.load()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in bg_b200.h but not exported by libbgb200.so"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature in lib.SIGNATURES"
    assert sorted(lib.SIGNATURES) == names
    assert handle.bg_version() >= 100


def test_struct_layouts_match_header():
    # sizes follow from the header's field lists (x86-64 SysV): guards against the ctypes mirror drifting
    assert ctypes.sizeof(lib.BgGraph) == 6 * 8 + 2 * 8 + 2 * 4
    assert ctypes.sizeof(lib.BgSeg) == 24
    assert lib.MAX_SEG == int(re.search(r"#define BG_MAX_SEG (\d+)", open(os.path.join(ROOT, "include", "bg_b200.h")).read()).group(1))


def test_host_errors_are_reported():
    import torch
    with pytest.raises(RuntimeError, match="out of range"):
        lib.csr_build_host(torch.tensor([[5], [0]]), 2)
    assert "out of range" in lib.last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "building-gan-graph-conditioned-architectural-volume-generation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"
