"""CPU test: libb200lk.so loads and exports every symbol include/b200lk.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from iqtree2_b200 import build as blk
from iqtree2_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200lk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200lk_[a-z0-9_]+)\s*\(", src)))


def declared_tree_symbols():
    src = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200(?:pt|st|ut)_[a-z0-9_]+)\s*\(", src)))


def test_tree_level_symbols_exported():
    """include/b200pt.h: the PhyloTree mirror (b200pt_*) and the partition model (b200st_*)"""
    lib = ctypes.CDLL(blk.build())
    syms = declared_tree_symbols()
    assert len([s for s in syms if s.startswith("b200st_")]) >= 20
    assert len([s for s in syms if s.startswith("b200ut_")]) >= 11
    for sym in syms:
        assert hasattr(lib, sym), sym


def test_header_symbols_match_binding_list():
    assert declared_symbols() == sorted(engine.ABI_SYMBOLS)


def test_library_builds_and_exports_all_symbols():
    lib_path = blk.build()
    lib = ctypes.CDLL(lib_path)
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym
    assert lib.b200lk_abi_version() == 1


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from iqtree2_b200 import models
    with pytest.raises(engine.EngineError, match="no CPU fallback"):
        engine.LikelihoodEngine(models.make_gtr_gamma(), 4, 16)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "iqtree2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # no import, link, dlopen or path reference to the checker from product code
                for pat in (r"^\s*(import|from)\s+oracle", r"oracle[/.]", r"lk_oracle", r"liboracle"):
                    assert not re.search(pat, text, flags=re.M), (dirpath, f, pat)
