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


def test_every_struct_layout_matches_the_header_as_compiled_by_gcc(tmp_path):
    """include/nrx.h is plain C99: gcc compiles it, prints sizeof and every field offset of every struct, and the ctypes
    mirrors in news_recsys_b200/_lib.py must agree field by field (what a cgo / JNI binding of the same header would see)."""
    import subprocess
    from news_recsys_b200 import _lib
    structs = ["NrxFeat", "NrxRowOpt", "NrxIngestCol", "NrxPeerStep", "NrxShardFeat", "NrxTower", "NrxTowerHead", "NrxTopkPeer"]
    src = open(os.path.join(ROOT, "include", "nrx.h")).read()
    declared = re.findall(r"^\} (Nrx[A-Za-z]+);", src, flags=re.M)
    assert sorted(declared) == sorted(structs), "a struct of nrx.h has no layout check here"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nrx.h"', "int main(void) {"]
    for s in structs:
        cls = getattr(_lib, s)
        lines.append(f'  printf("{s} %zu\\n", sizeof({s}));')
        for name, _ in cls._fields_:
            lines.append(f'  printf("{s}.{name} %zu\\n", offsetof({s}, {name}));')
    lines += ["  return 0;", "}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for s in structs:
        cls = getattr(_lib, s)
        assert ctypes.sizeof(cls) == int(out[s]), f"sizeof({s}): ctypes {ctypes.sizeof(cls)} != C {out[s]}"
        for name, _ in cls._fields_:
            assert getattr(cls, name).offset == int(out[f"{s}.{name}"]), f"{s}.{name}: ctypes offset {getattr(cls, name).offset} != C {out[s + '.' + name]}"


def test_ctypes_argument_counts_match_the_header():
    """Every entry point: the number of parameters in include/nrx.h == len(argtypes) in _lib.SIGNATURES."""
    from news_recsys_b200 import _lib
    src = open(os.path.join(ROOT, "include", "nrx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    bad = []
    for m in re.finditer(r"\b(nrx_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        if name in _lib.SIGNATURES and len(_lib.SIGNATURES[name][1]) != n:
            bad.append((name, n, len(_lib.SIGNATURES[name][1])))
    assert not bad, f"(entry point, header params, ctypes params): {bad}"


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md (the binding a maintainer of the reference would write) mentions every entry point of include/nrx.h,
    by name, by a `nrx_a/b/c` shorthand or by the `nrx_*_workspace_bytes` family."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    named = set(re.findall(r"nrx_[a-z0-9_]+", doc))
    for m in re.finditer(r"`(nrx_[a-z0-9_]+(?:/[a-z0-9_]+)+)`", doc):
        parts = m.group(1).split("/")
        stem = parts[0][: parts[0].rfind("_") + 1]
        named |= {stem + p for p in parts[1:]}
    missing = [s for s in _header_symbols() if s not in named and not s.endswith("_workspace_bytes")]
    assert not missing, missing
