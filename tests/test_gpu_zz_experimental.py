"""Experimental, off-by-default library options (prepared without GPU time at the end of round 1): whatever they do
to speed, results must not move.  Kept in the last test module so that the established parity tests run first."""
import pytest

from helpers import assert_trace_matches
from test_gpu_parity import as_single, gpu, run_single  # noqa: F401  (gpu is a fixture)

pytestmark = pytest.mark.gpu


def test_experimental_layout_and_fold_options_keep_results(gpu, golden):
    """fold_variant=1 (16-byte accesses, tail of 1-3 slots handled separately) and degree_sort (layout) are
    experiments that stay off by default; whatever they do to speed, results must not move."""
    for name in ("rgg_n16384_p1", "rgg_n65536_p1", "hand_loops_multi_p1", "hand_star41_p1", "hand_k66_p1", "hand_path16_p1"):
        case = golden[name]
        nv, parts, rowptr, edges = as_single(case)
        for opts in ({"fold_variant": 1}, {"reorder": 1, "region_size": 64, "degree_sort": 256},
                     {"reorder": 1, "region_size": 64, "degree_sort": 512},
                     {"fold_variant": 1, "reorder": 1, "region_size": 64, "degree_sort": 1024},
                     {"reorder": 1, "region_size": 512, "degree_sort": 2048}):
            res = run_single(gpu, parts, rowptr, edges, nv, **opts)
            assert_trace_matches(case, res["iters"], res["modularity"], res["trace"], None, res["comm"])
