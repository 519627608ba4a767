"""world_size-2 CPU test of the N>1 host path (gloo): every rank builds only its own strip, the driver's
collectives (edge-count sum, max-over-ranks, communicator-id broadcast) work, and the union of the
independently built strips is exactly the 2-rank reference graph (checked through the pinned C oracle
against the golden trace of the unmodified reference run on 2 ranks)."""
import json
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from minivite_b200 import dist as D
    R = D.Ranks("gloo")
    ss = D.my_strip(16384, world, rank)
    sh = ss.shards[0]
    ne = int(R.allreduce(float(sh.lne)))
    mx = R.allreduce(float(rank), "max")
    ident = R.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
    rps = R.gather_arrays(np.array(sh.rowptr))
    eds = R.gather_arrays(np.array(sh.edges))
    R.barrier()
    if rank == 0:
        from oracle import oracle as O
        res = O.louvain(D.strip_parts(16384, world), rps, eds)
        json.dump({"ne": ne, "max": mx, "id_ok": ident == bytes(range(128)), "iters": res["iters"],
                   "mod": repr(res["modularity"]), "chash": "%016x" % res["chash_final"],
                   "base": [int(sh.base)], "lnv": int(sh.lnv)}, open(os.path.join(out_dir, "r0.json"), "w"))
    else:
        json.dump({"id_ok": ident == bytes(range(128)), "base": int(sh.base), "lnv": int(sh.lnv)},
                  open(os.path.join(out_dir, f"r{rank}.json"), "w"))
    R.shutdown()


def test_two_rank_host_path(tmp_path, golden):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = json.load(open(tmp_path / "r0.json"))
    r1 = json.load(open(tmp_path / "r1.json"))
    case = golden["rgg_n16384_p2"]
    assert r0["ne"] == case["ne"] and r0["max"] == 1.0
    assert r0["id_ok"] and r1["id_ok"]
    assert r1["base"] == 8192 and r1["lnv"] == 8192
    assert r0["iters"] == case["iters"] and r0["mod"] == repr(float(case["modularity"]))
    assert r0["chash"] == case["final_chash"]
