#!/usr/bin/env python3
"""Minimal driver for ncu captures: one RGG, device-resident inputs, N Louvain phases, nothing else.
usage: python tools/profile_run.py [nv] [runs] [pct_random_edges]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from minivite_b200 import gpu as G  # noqa: E402
from minivite_b200 import hostgraph as hg  # noqa: E402

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 16777216
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
t = time.time()
ss = hg.generate_rgg(nv, 1, random_edge_percent=pct)
sh = ss.shards[0]
print(f"generated nv={nv} ne={sh.lne} in {time.time() - t:.1f}s", flush=True)
d_rowptr = torch.from_numpy(np.ascontiguousarray(sh.rowptr)).cuda()
d_edges = torch.from_numpy(np.ascontiguousarray(sh.edges).view(np.uint8)).cuda()
torch.cuda.synchronize()
ctx = G.LouvainGPU(0, 0, 1)
ctx.attach_device(nv, sh.parts, sh.lnv, sh.lne, d_rowptr.data_ptr(), d_edges.data_ptr())
for r in range(runs):
    mod, iters = ctx.louvain()
    tm = ctx.timings()
    print(f"run {r}: mod={mod:.17g} iters={iters} total={tm['total_s']*1e3:.3f}ms setup={tm['setup_s']*1e3:.3f}ms "
          f"scan={tm['scan_s']*1e3:.3f}ms ({tm['scan_s']/iters*1e3:.3f} ms/iter) fold={tm['fold_s']*1e3:.3f}ms "
          f"edges/s={sh.lne*iters/tm['total_s']:.4g}", flush=True)
    if r == runs - 1:
        print("scan ms per iteration:", " ".join(f"{x*1e3:.2f}" for x in ctx.scan_times()), flush=True)
ctx.close()
