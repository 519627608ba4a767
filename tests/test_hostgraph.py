"""CPU tests of the host-side graph tools (reference Graph / BinaryEdgeList / GenerateRGG surface)."""
import numpy as np
import pytest

from minivite_b200 import hostgraph as hg


def test_binary_file_roundtrip_and_vertex_split(tmp_path):
    """write (graph.hpp:342-403 format) -> read on 1 and 3 ranks (graph.hpp:344: M*(r+1)/p - M*r/p vertices each)."""
    ss = hg.generate_rgg(8192, 2)
    path = str(tmp_path / "g.bin")
    ss.write(path)
    whole = hg.read_graph(path, 0, 1).shards[0]
    assert whole.nv == 8192 and whole.lne == sum(s.lne for s in ss.shards) and whole.ne == whole.lne
    assert np.array_equal(whole.edges["tail"], np.concatenate([s.edges["tail"] for s in ss.shards]))
    raw = np.fromfile(path, dtype=np.int64, count=2)
    assert list(raw) == [8192, whole.lne]
    parts = [(8192 * r) // 3 for r in range(4)]
    lne = 0
    for r in range(3):
        sh = hg.read_graph(path, r, 3).shards[0]
        assert sh.base == parts[r] and sh.lnv == parts[r + 1] - parts[r] and sh.rowptr[0] == 0
        assert np.array_equal(sh.edges["tail"], whole.edges["tail"][whole.rowptr[parts[r]]:whole.rowptr[parts[r + 1]]])
        lne += sh.lne
    assert lne == whole.lne


def test_generator_argument_checks():
    with pytest.raises(RuntimeError, match="divisible"):
        hg.generate_rgg(1000, 3)
    with pytest.raises(RuntimeError, match="power of 2"):
        hg.generate_rgg(1200, 6)
    with pytest.raises(RuntimeError):
        hg.generate_rgg(64, 32)            # 1/p > rn violated (graph.hpp:633)


def test_strip_built_alone_equals_strip_built_with_all():
    all_ = hg.generate_rgg(16384, 4)
    for r in (0, 2, 3):
        one = hg.generate_rgg(16384, 4, r, r + 1).shards[0]
        assert one.base == all_.shards[r].base
        assert np.array_equal(one.rowptr, all_.shards[r].rowptr) and np.array_equal(one.edges, all_.shards[r].edges)


def test_graph_is_symmetric_and_sorted():
    ss = hg.generate_rgg(8192, 1, random_edge_percent=10.0)
    sh = ss.shards[0]
    src = np.repeat(np.arange(sh.lnv), np.diff(sh.rowptr))
    dst = sh.edges["tail"]
    fwd = np.stack([src, dst], 1)
    rev = np.stack([dst, src], 1)
    a = fwd[np.lexsort((fwd[:, 1], fwd[:, 0]))]
    b = rev[np.lexsort((rev[:, 1], rev[:, 0]))]
    assert np.array_equal(a, b)                      # every edge has its reverse (multi-edges included)
    for v in range(0, sh.lnv, 97):
        seg = dst[sh.rowptr[v]:sh.rowptr[v + 1]]
        assert np.all(np.diff(seg) >= 0)


def test_generator_rng_known_answers():
    """SURVEY.md 8(c) RNG known answers of the reference generator (utils.hpp:91-98, graph.hpp:682-700): the seed,
    the minstd_rand0 stream behind std::default_random_engine and the first coordinates at 1 rank."""
    seed = hg.reseeder(1)
    assert seed == 1967017404
    raw, x = [], seed
    for _ in range(4):
        x = x * 16807 % 2147483647
        raw.append(x)
    assert raw == [1298247110, 1205324250, 671427599, 1804575055]
    X, Y = hg.rgg_points(16384, 1, 0, 2)
    assert repr(float(X[0])) == "0.5612728422168126" and float(X[0]) == 0.56127284221681262
    assert float(Y[0]) == 0.84032074361727549 and float(X[1]) == 0.27303256924115216
    # every strip restarts from the same seed: same X, Y shifted into the strip (graph.hpp:697-700)
    X2, Y2 = hg.rgg_points(16384, 4, 3, 2)
    X0, Y0 = hg.rgg_points(16384, 4, 0, 2)
    assert np.array_equal(X2, X0) and np.all((Y2 >= 0.75) & (Y2 < 1.0)) and np.all(Y0 < 0.25)
