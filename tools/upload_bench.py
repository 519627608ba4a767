#!/usr/bin/env python3
"""Time mvgpu_upload_shard variants: python tools/upload_bench.py NV"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from minivite_b200 import gpu as G  # noqa: E402
from minivite_b200 import hostgraph as hg  # noqa: E402

nv = int(sys.argv[1])
ss = hg.generate_rgg(nv, 1)
sh = ss.shards[0]
h_rowptr = torch.from_numpy(np.ascontiguousarray(sh.rowptr)).pin_memory()
h_edges = torch.from_numpy(np.ascontiguousarray(sh.edges).view(np.uint8)).pin_memory()
rp, ed = h_rowptr.numpy(), h_edges.numpy().view(hg.EDGE_DTYPE)
print("omp env", os.environ.get("OMP_NUM_THREADS"), "cores", len(os.sched_getaffinity(0)), flush=True)
for spec in ["compact_upload=0", "compact_upload=1,host_threads=4", "compact_upload=1,host_threads=8", "compact_upload=1,host_threads=16",
             "compact_upload=1,host_threads=32", "compact_upload=1,host_threads=64"]:
    ctx = G.LouvainGPU(0, 0, 1)
    for kv in spec.split(","):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ts = []
    for r in range(4):
        t = time.perf_counter()
        ctx.upload(nv, sh.parts, rp, ed)
        ts.append(time.perf_counter() - t)
    mod, it = ctx.louvain()
    tm = ctx.timings()
    t = time.perf_counter(); comm = ctx.communities(); tc = time.perf_counter() - t
    print(f"{spec:40s} upload wall ms {[round(x*1e3,1) for x in ts]} h2d_event_ms {tm['h2d_s']*1e3:.1f} bytes {tm['h2d_bytes']} louvain {tm['total_s']*1e3:.1f} ms setup {tm['setup_s']*1e3:.2f} comm_d2h {tc*1e3:.1f} ms", flush=True)
    ctx.close()
