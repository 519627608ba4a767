"""Device-side RGG generator: byte-identical to the host generator, which is itself
byte-identical to the reference's GenerateRGG (tests/test_oracle.py pins that through the golden traces)."""
import numpy as np
import pytest

from helpers import assert_trace_matches

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,ranks,unit", [(16384, 1, [0], True), (16384, 2, [0, 1], True), (16384, 4, [0, 1, 2, 3], True),
                                            (16384, 8, [0, 3, 7], True), (32768, 4, [1, 2], False),
                                            (1048576, 1, [0], True), (1048576, 8, [0, 5], True)])
def test_device_generator_matches_host_generator(n, p, ranks, unit):
    from minivite_b200 import gpu as G
    from minivite_b200 import hostgraph as hg
    for r in ranks:
        ref = hg.generate_rgg(n, p, r, r + 1, unit_weight=unit).shards[0]
        g = G.LouvainGPU(0, r, p)          # no communicator needed: generation is local to the rank
        try:
            lne = g.generate_rgg(n, unit_weight=unit)
            assert lne == ref.lne, (n, p, r, lne, ref.lne)
            rowptr, edges = g.download_shard()
        finally:
            g.close()
        assert np.array_equal(rowptr, ref.rowptr)
        assert np.array_equal(edges["tail"], ref.edges["tail"])
        assert np.array_equal(edges["weight"], ref.edges["weight"])   # bit-exact, also for Euclidean weights


def test_generate_then_louvain_matches_reference(golden):
    """`miniVite -n 16384` end to end on the device: generator + Louvain phase == reference trace."""
    from minivite_b200 import gpu as G
    case = golden["rgg_n16384_p1"]
    g = G.LouvainGPU(0, 0, 1)
    try:
        g.set_option("trace", 1)
        assert g.generate_rgg(16384) == case["ne"]
        mod, iters = g.louvain()
        assert_trace_matches(case, iters, mod, g.trace(), None, g.communities())
    finally:
        g.close()
