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


def test_compact_upload_host_pass():
    """The host half of compact_upload (narrow.cpp, AVX2 or scalar body): 16-byte {tail, weight} records -> int32 tails,
    with the same validation as the device pass and the count of non-owned tails.  Pure host code inside libmvgpu.so."""
    import ctypes
    import numpy as np
    from minivite_b200 import gpu as G
    from minivite_b200 import hostgraph as hg
    L = G.lib()
    fn = L.mv_narrow_edges
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                   ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int)]

    def run(edges, nv, base, bound):
        out = np.full(len(edges) + 4, -7, np.int32)
        nrem, bad = ctypes.c_longlong(0), ctypes.c_int(0)
        fn(edges.ctypes.data, len(edges), nv, base, bound, out.ctypes.data, ctypes.byref(nrem), ctypes.byref(bad))
        assert np.all(out[len(edges):] == -7)                      # nothing written past the end
        return out[:len(edges)], nrem.value, bad.value

    rng = np.random.RandomState(3)
    for n in (0, 1, 3, 4, 5, 63, 64, 1001):
        e = np.zeros(n, hg.EDGE_DTYPE)
        e["tail"] = rng.randint(0, 5000, size=n)
        e["weight"] = 1.0
        out, nrem, bad = run(e, 5000, 1000, 3000)
        assert bad == 0 and np.array_equal(out, e["tail"].astype(np.int32))
        assert nrem == int(np.sum((e["tail"] < 1000) | (e["tail"] >= 3000)))
    e = np.zeros(37, hg.EDGE_DTYPE)
    e["tail"] = np.arange(37)
    e["weight"] = 1.0
    for pos in (0, 5, 35, 36):                                      # vector body and scalar tail
        for field, val in (("weight", 0.5), ("weight", float("nan")), ("tail", -1), ("tail", 37)):
            f = e.copy()
            f[field][pos] = val
            assert run(f, 37, 0, 37)[2] == 1, (pos, field, val)
    assert run(e, 37, 0, 37)[1:] == (0, 0)


def test_narrow_edges_host_pass_matches_numpy():
    """mv_narrow_edges (host half of the compact upload; AVX-512 / AVX2 / scalar bodies picked at run time): tails,
    remote count and the validation flag against a numpy restatement, on ragged lengths and misaligned outputs."""
    import ctypes
    import numpy as np
    from minivite_b200 import gpu as G
    L = G.lib()
    L.mv_narrow_edges.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                                  ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int)]
    L.mv_narrow_edges.restype = None
    rng = np.random.default_rng(5)
    nv, base, bound = 1 << 20, 300000, 700000
    for n in (0, 1, 7, 8, 9, 63, 1000, 4099):
        for off in (0, 1, 3):
            rec = np.zeros(n, np.dtype([("tail", "<i8"), ("weight", "<f8")]))
            rec["tail"] = rng.integers(0, nv, n)
            rec["weight"] = 1.0
            buf = np.zeros(n + 16, np.int32)
            dst = buf[off:off + n]
            nrem, bad = ctypes.c_longlong(0), ctypes.c_int(0)
            L.mv_narrow_edges(rec.ctypes.data, n, nv, base, bound, dst.ctypes.data, ctypes.byref(nrem), ctypes.byref(bad))
            assert bad.value == 0 and np.array_equal(dst, rec["tail"].astype(np.int32))
            assert nrem.value == int(((rec["tail"] < base) | (rec["tail"] >= bound)).sum())
            assert buf[:off].sum() == 0 and buf[off + n:].sum() == 0
            if n:
                for field, val in (("weight", 2.0), ("tail", -1), ("tail", nv)):
                    r2 = rec.copy()
                    r2[field][n // 2] = val
                    bad = ctypes.c_int(0)
                    L.mv_narrow_edges(r2.ctypes.data, n, nv, base, bound, dst.ctypes.data, ctypes.byref(nrem), ctypes.byref(bad))
                    assert bad.value == 1, (n, field, val)
