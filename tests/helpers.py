"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from minivite_b200 import hostgraph as hg


def case_graph(case):
    """Shards (parts, rowptrs, edge arrays) for a golden case, rebuilt with OUR generator / from the stored graph."""
    kind = case["kind"]
    p = case["nranks"]
    if kind == "rgg":
        args = case["args"]
        ss = hg.generate_rgg(case["n"], p, lcg="-l" in args, unit_weight="-w" not in args)
        return ss.shards[0].parts, [s.rowptr for s in ss.shards], [s.edges for s in ss.shards], ss
    if kind == "file_rgg":
        ss = hg.generate_rgg(case["n"], case["strips"], unit_weight=case["unit_weight"])
        nv = case["n"]
        rowptr = np.concatenate([[0]] + [s.rowptr[1:] + off for s, off in
                                         zip(ss.shards, np.cumsum([0] + [s.lne for s in ss.shards[:-1]]))])
        edges = np.concatenate([s.edges for s in ss.shards])
        return split_global(nv, rowptr.astype(np.int64), edges, p) + (ss,)
    if kind == "file_balanced":
        ss = hg.generate_rgg(case["n"], 1, random_edge_percent=case["pct"])
        sh = ss.shards[0]
        parts = np.array(case["parts"], dtype=np.int64)
        rps, eds = [], []
        for r in range(p):
            a, b = parts[r], parts[r + 1]
            rps.append(np.ascontiguousarray(sh.rowptr[a:b + 1] - sh.rowptr[a]))
            eds.append(np.ascontiguousarray(sh.edges[sh.rowptr[a]:sh.rowptr[b]]))
        return parts, rps, eds, ss
    if kind == "rmat":
        n, rowptr, edges = rmat_graph(case["scale"], case["edge_factor"], case["seed"])
        if case.get("balanced"):
            parts = np.array(case["parts"], dtype=np.int64)     # the reference's own -b bins (graph.hpp:466-572)
            rps = [np.ascontiguousarray(rowptr[a:b + 1] - rowptr[a]) for a, b in zip(parts[:-1], parts[1:])]
            eds = [np.ascontiguousarray(edges[rowptr[a]:rowptr[b]]) for a, b in zip(parts[:-1], parts[1:])]
            return parts, rps, eds, None
        return split_global(n, rowptr, edges, p) + (None,)
    if kind == "hand":
        g = case["graph"]
        edges = np.zeros(len(g["tails"]), hg.EDGE_DTYPE)
        edges["tail"] = g["tails"]
        edges["weight"] = g["weights"]
        return split_global(g["nv"], np.array(g["rowptr"], np.int64), edges, p) + (None,)
    raise ValueError(kind)


def split_global(nv, rowptr, edges, p):
    """Vertex-range split parts[r] = nv*r/p of a global CSR (reference graph.hpp:112-113 / 344-355)."""
    parts = np.array([(nv * r) // p for r in range(p + 1)], dtype=np.int64)
    rps, eds = [], []
    for r in range(p):
        a, b = parts[r], parts[r + 1]
        rp = rowptr[a:b + 1] - rowptr[a]
        rps.append(np.ascontiguousarray(rp))
        eds.append(np.ascontiguousarray(edges[rowptr[a]:rowptr[b]]))
    return parts, rps, eds


def assert_trace_matches(case, iters, modularity, trace, final_chash=None, comm=None, exact=True, tol=1e-6):
    """Compare a run (C oracle or CUDA path) with a golden reference record."""
    assert iters == case["iters"], (iters, case["iters"])
    gm = float(case["modularity"])
    if exact:
        assert modularity == gm, (repr(modularity), case["modularity"])
    else:
        assert abs(modularity - gm) <= tol
    assert len(trace) == len(case["trace"])
    for k, (t, g) in enumerate(zip(trace, case["trace"])):
        tm = float(t["modularity"])
        if exact:
            assert tm == float(g["modularity"]), (k, repr(tm), g["modularity"])
            assert int(t["moved"]) == g["moved"], (k, int(t["moved"]), g["moved"])
            assert int(t["chash"]) == int(g["chash"], 16), (k, hex(int(t["chash"])), g["chash"])
        else:
            assert abs(tm - float(g["modularity"])) <= tol
    if final_chash is not None and exact:
        assert final_chash == int(case["final_chash"], 16)
    if comm is not None and "comm" in case and exact:
        assert [int(x) for x in comm] == case["comm"]


def rmat_graph(scale, edge_factor, seed, a=0.57, b=0.19, c=0.19):
    """Power-law (R-MAT) graph in the reference's CSR format: symmetric, unit weights, no self loops, no parallel
    edges, adjacency sorted by tail.  numpy's legacy RandomState keeps the stream stable across versions, so the graph
    is a function of (scale, edge_factor, seed) and only its golden TRACE needs committing."""
    rng = np.random.RandomState(seed)
    n, m = 1 << scale, edge_factor << scale
    src = np.zeros(m, np.int64)
    dst = np.zeros(m, np.int64)
    for level in range(scale):
        r = rng.random_sample(m)
        right = (r >= a) & (r < a + b) | (r >= a + b + c)          # quadrants b and d set the column bit
        down = r >= a + b                                          # quadrants c and d set the row bit
        src |= down.astype(np.int64) << level
        dst |= right.astype(np.int64) << level
    perm = rng.permutation(n)                                      # scatter the hubs over the id range
    src, dst = perm[src], perm[dst]
    keep = src != dst
    lo, hi = np.minimum(src[keep], dst[keep]), np.maximum(src[keep], dst[keep])
    key = np.unique(lo * n + hi)
    lo, hi = key // n, key % n
    s2, d2 = np.concatenate([lo, hi]), np.concatenate([hi, lo])
    order = np.lexsort((d2, s2))
    s2, d2 = s2[order], d2[order]
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, s2 + 1, 1)
    rowptr = np.cumsum(rowptr)
    edges = np.zeros(len(d2), hg.EDGE_DTYPE)
    edges["tail"] = d2
    edges["weight"] = 1.0
    return n, rowptr, edges
