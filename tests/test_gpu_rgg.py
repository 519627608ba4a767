"""Device-side RGG generator: byte-identical to the host generator, which is itself
byte-identical to the reference's GenerateRGG (tests/test_oracle.py pins that through the golden traces)."""
import numpy as np
import pytest

from helpers import assert_trace_matches

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,ranks,unit,lcg", [(16384, 1, [0], True, False), (16384, 2, [0, 1], True, False),
                                                (16384, 4, [0, 1, 2, 3], True, False), (16384, 8, [0, 3, 7], True, False),
                                                (32768, 4, [1, 2], False, False), (1048576, 1, [0], True, False),
                                                (1048576, 8, [0, 5], True, False),
                                                # -l: the reference's LCG stream, a different point pattern per strip
                                                (16384, 1, [0], True, True), (16384, 2, [0, 1], True, True),
                                                (32768, 4, [0, 1, 2, 3], False, True), (1048576, 8, [0, 4, 7], True, True)])
def test_device_generator_matches_host_generator(n, p, ranks, unit, lcg):
    from minivite_b200 import gpu as G
    from minivite_b200 import hostgraph as hg
    for r in ranks:
        ref = hg.generate_rgg(n, p, r, r + 1, unit_weight=unit, lcg=lcg).shards[0]
        g = G.LouvainGPU(0, r, p)          # no communicator needed: generation is local to the rank
        try:
            lne = g.generate_rgg(n, unit_weight=unit, lcg=lcg)
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


def test_generate_lcg_then_louvain_matches_reference(golden):
    """`miniVite -n 16384 -l` on 2 ranks (golden from the reference's own generator), both strips generated on the device
    and run as one shard: generator (-l) + Louvain phase == reference trace."""
    from minivite_b200 import gpu as G
    case = golden["rgg_n16384_p2_l"]
    rps, eds = [], []
    for r in range(2):
        g = G.LouvainGPU(0, r, 2)
        try:
            g.generate_rgg(16384, lcg=True)
            rp, ed = g.download_shard()
        finally:
            g.close()
        rps.append(rp)
        eds.append(ed)
    assert sum(len(e) for e in eds) == case["ne"]
    rowptr = np.concatenate([rps[0], rps[1][1:] + rps[0][-1]])
    g = G.LouvainGPU(0, 0, 1)
    try:
        g.set_option("trace", 1)
        g.upload(16384, np.array([0, 16384], np.int64), rowptr, np.concatenate(eds))
        mod, iters = g.louvain()
        assert_trace_matches(case, iters, mod, g.trace(), None, None)
    finally:
        g.close()
