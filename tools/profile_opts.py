#!/usr/bin/env python3
"""Like profile_run.py but with options: python tools/profile_opts.py NV 'k=v,k=v' [runs]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from minivite_b200 import gpu as G  # noqa: E402
from minivite_b200 import hostgraph as hg  # noqa: E402

nv = int(sys.argv[1])
spec = sys.argv[2] if len(sys.argv) > 2 else ""
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ss = hg.generate_rgg(nv, 1)
sh = ss.shards[0]
d_rowptr = torch.from_numpy(np.ascontiguousarray(sh.rowptr)).cuda()
d_edges = torch.from_numpy(np.ascontiguousarray(sh.edges).view(np.uint8)).cuda()
torch.cuda.synchronize()
ctx = G.LouvainGPU(0, 0, 1)
for kv in spec.split(","):
    if kv:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
ctx.attach_device(nv, sh.parts, sh.lnv, sh.lne, d_rowptr.data_ptr(), d_edges.data_ptr())
for r in range(runs):
    mod, iters = ctx.louvain()
    tm = ctx.timings()
    print(f"run {r}: mod={mod:.17g} iters={iters} total={tm['total_s']*1e3:.3f}ms scan={tm['scan_s']*1e3:.3f}ms", flush=True)
ctx.close()
