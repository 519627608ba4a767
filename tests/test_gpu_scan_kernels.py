"""Both shipped neighbour-scan kernels -- k_scan_pw (scan_variant 4: persistent warps, TMA-fed double buffer,
scan_pipe.cuh) and k_scan_ws (3: one CTA per 128-vertex tile), k_scan_pq (5: k_scan_pw with a per-warp ring of hard vertices) -- against the reference goldens and the C oracle,
including the paths only unusual graphs reach: groups whose edges overflow one staging buffer (sub-ranges with
synchronous bulk copies), vertices handed to the high-degree kernel, ragged last groups, weights."""
import numpy as np
import pytest

from helpers import assert_trace_matches
from test_gpu_parity import as_single, gpu, is_weighted, run_single  # noqa: F401  (gpu is a fixture)

pytestmark = pytest.mark.gpu
EDGE = np.dtype([("tail", "<i8"), ("weight", "<f8")])
VARIANTS = [3, 4, 5, 6]


@pytest.mark.parametrize("variant", VARIANTS)
def test_golden_cases(gpu, golden, variant):
    for name, case in golden.items():
        nv, parts, rowptr, edges = as_single(case)
        res = run_single(gpu, parts, rowptr, edges, nv, scan_variant=variant)
        if is_weighted(name, case):
            if case["nranks"] == 1:
                assert abs(res["modularity"] - float(case["modularity"])) <= 1e-6, name
        else:
            assert_trace_matches(case, res["iters"], res["modularity"], res["trace"], None, res["comm"])


@pytest.mark.parametrize("variant", VARIANTS)
def test_options_keep_results(gpu, golden, variant):
    for name in ("rgg_n65536_p1", "hand_loops_multi_p1", "hand_star41_p1", "hand_k66_p1"):
        case = golden[name]
        nv, parts, rowptr, edges = as_single(case)
        for opts in ({"cache_policy": 0}, {"reorder": 1, "region_size": 64}, {"force_weighted": 1}, {"first_iter": 0},
                     {"first_iter": 0, "reorder": 1, "region_size": 128},
                     {"force_heavy_deg": 8, "reorder": 1, "region_size": 32}, {"force_weighted": 1, "reorder": 1, "force_heavy_deg": 5}):
            res = run_single(gpu, parts, rowptr, edges, nv, scan_variant=variant, **opts)
            assert_trace_matches(case, res["iters"], res["modularity"], res["trace"], None, res["comm"] if "comm" in case else None)


def random_graph(n, avg_deg, seed, hubs=0, hub_deg=0, self_loops=0, multi=0, blocks=0):
    """Symmetric random multigraph in the reference's CSR format (unit weights), adjacency sorted by tail.
    blocks > 0: planted partition (90 % of the edges inside `blocks` equal groups of scattered vertex ids)."""
    rng = np.random.default_rng(seed)
    m = n * avg_deg // 2
    a, b = rng.integers(0, n, m), rng.integers(0, n, m)
    if blocks:
        inside = rng.random(m) < 0.9
        b = np.where(inside, (b // blocks) * blocks + a % blocks, b) % n      # same residue class = same block
    keep = a != b
    a, b = a[keep], b[keep]
    for h in range(hubs):
        t = rng.choice(n, hub_deg, replace=False)
        t = t[t != h]
        a, b = np.concatenate([a, np.full(len(t), h)]), np.concatenate([b, t])
    key = np.unique(np.minimum(a, b) * n + np.maximum(a, b))          # simple graph first
    a, b = key // n, key % n
    if multi:
        pick = rng.integers(0, len(a), multi)
        a, b = np.concatenate([a, a[pick]]), np.concatenate([b, b[pick]])
    src, dst = np.concatenate([a, b]), np.concatenate([b, a])
    if self_loops:
        s = rng.integers(0, n, self_loops)
        src, dst = np.concatenate([src, s]), np.concatenate([dst, s])
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, src + 1, 1)
    rowptr = np.cumsum(rowptr)
    edges = np.zeros(len(dst), EDGE)
    edges["tail"] = dst
    edges["weight"] = 1.0
    return rowptr, edges


@pytest.mark.parametrize("variant", VARIANTS)
def test_dense_and_skewed_graphs_against_oracle(gpu, variant):
    """Graphs the RGG goldens never produce: 32-vertex groups with thousands of edges (sub-ranges), genuine hubs
    above every tile capacity (high-degree kernel, unforced), self loops, multi-edges, a ragged last group."""
    from oracle import oracle as O
    for (n, deg, kw) in [(1000, 60, {}), (777, 150, {"self_loops": 40, "multi": 300}), (6000, 8, {"hubs": 3, "hub_deg": 3000}),
                         (4099, 30, {"hubs": 2, "hub_deg": 900, "multi": 100}), (33, 20, {}), (2500, 700, {}),
                         (3000, 80, {"blocks": 25}), (20000, 40, {"blocks": 400, "hubs": 1, "hub_deg": 2000})]:
        rowptr, edges = random_graph(n, deg, seed=n + deg, **kw)
        parts = np.array([0, n], np.int64)
        ref = O.louvain(parts, [rowptr], [edges])
        for opts in ({}, {"reorder": 1, "region_size": 64}, {"first_iter": 0}):
            res = run_single(gpu, parts, rowptr, edges, n, scan_variant=variant, **opts)
            assert res["iters"] == ref["iters"] and res["modularity"] == ref["modularity"], (n, deg, kw, opts)
            assert np.array_equal(res["comm"], ref["comm"][0]), (n, deg, kw, opts)
            assert [int(x) for x in res["trace"]["chash"]] == [int(x) for x in ref["trace"]["chash"]]
        if res["info"]["maxdeg"] > 2048:                  # above every tile capacity: the high-degree kernel ran, unforced
            assert res["info"]["nheavy"] > 0


@pytest.mark.parametrize("variant", VARIANTS)
def test_power_law_graphs_reach_the_high_degree_kernel_unforced(gpu, golden_rmat, variant):
    """R-MAT graphs read the way `miniVite -f` reads them: hubs of degree 3 684 / 15 706 are far above every tile
    capacity, so k_scan_heavy runs without the force_heavy_deg test hook; traces equal the unmodified reference's."""
    for name in ("rmat_s14_p1", "rmat_s17_p1"):
        case = golden_rmat[name]
        nv, parts, rowptr, edges = as_single(case)
        for opts in ({}, {"reorder": 1, "region_size": 128}):
            res = run_single(gpu, parts, rowptr, edges, nv, scan_variant=variant, **opts)
            assert res["info"]["maxdeg"] == case["maxdeg"] and res["info"]["nheavy"] > 0
            assert_trace_matches(case, res["iters"], res["modularity"], res["trace"], None, None)
            assert repr(res["constant"]) == case["constant"]
