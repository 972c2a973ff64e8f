"""CPU-only: the C-ABI library builds, loads and exports every symbol include/nrx.h declares, with the
ctypes signatures in news_recsys_b200/_lib.py covering exactly that set.  No compute call is made."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "nrx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nrx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = _header_symbols()
    assert len(syms) >= 30
    for must in ("nrx_embed_pool_fwd", "nrx_embed_bwd_plan", "nrx_embed_bwd_apply", "nrx_fm_fused_fwd", "nrx_tower_fwd",
                 "nrx_tower_bwd", "nrx_dcn_cross_fwd", "nrx_topk_search", "nrx_topk_merge", "nrx_last_error"):
        assert must in syms


def test_library_builds_and_exports_every_header_symbol():
    from news_recsys_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in _header_symbols() if not hasattr(lib, s)]
    assert not missing, f"libnrx.so does not export {missing}"
    lib.nrx_version.restype = ctypes.c_int
    assert lib.nrx_version() == 100


def test_ctypes_table_matches_header():
    from news_recsys_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    _lib.load()
    assert _lib.MISSING == []


def test_struct_sizes_match_c_layout():
    """NrxFeat / NrxRowOpt / NrxTower mirror the C structs (checked against sizes computed from the header)."""
    from news_recsys_b200 import _lib
    assert ctypes.sizeof(_lib.NrxFeat) == 8 + 8 + 4 * 4 + 8 + 4 + 4 + 8 + 8 + 4 + 4  # 72
    assert ctypes.sizeof(_lib.NrxRowOpt) == 6 * 4 + 16 * 8 * 2 + 8
    assert ctypes.sizeof(_lib.NrxTower) == 4 + 9 * 4 + 8 * 8 * 2 + 4 + 4


def test_product_path_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may touch oracle/."""
    pkg = os.path.join(ROOT, "news_recsys_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), f"{f} references the oracle"


def test_tools_never_import_the_oracle():
    """Dev tools under tools/ are not allowed to use the oracle either (measurement helpers that time it live in
    tests/tools)."""
    tools = os.path.join(ROOT, "tools")
    for f in os.listdir(tools):
        if f.endswith(".py"):
            src = open(os.path.join(tools, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"tools/{f} imports the oracle"


def test_ops_fail_loudly_without_cuda():
    import torch
    from news_recsys_b200 import ops
    from news_recsys_b200._lib import NrxError
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    w = torch.randn(10, 16)
    with pytest.raises(NrxError, match="CUDA"):
        ops.FeatBinding([ops.FeatSpec("a", "t", 0, 16, 1, False, 0)], {"t": w}, {"a": torch.tensor([1, 2])})
