"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports every symbol declared in
include/mvgpu.h, and refuses to compute (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mvgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvgpu_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported():
    import __graft_entry__ as ge
    ge.build()
    from minivite_b200 import gpu
    L = gpu.lib()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(gpu.EXPORTS) == syms


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from minivite_b200 import gpu
    with pytest.raises(gpu.MvgpuError):
        gpu.LouvainGPU(0, 0, 1)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under minivite_b200/ may import, link or call it."""
    pats = [r"^\s*(from|import)\s+oracle\b", r"liblouvain_oracle", r"\bmvo_[a-z_]+\s*\(", r"#include\s*[<\"].*oracle",
            r"miniVite_ref", r"_ref/"]
    for dp, _, files in os.walk(os.path.join(ROOT, "minivite_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                for line in open(os.path.join(dp, f), errors="replace"):
                    code = line.split("//")[0]
                    for pat in pats:
                        assert not re.search(pat, code), (os.path.join(dp, f), line.strip())
