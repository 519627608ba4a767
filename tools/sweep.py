#!/usr/bin/env python3
"""Option sweep on one graph: python tools/sweep.py NV 'scan_variant=0' 'scan_variant=1,cache_policy=0' ..."""
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
d_rowptr = torch.from_numpy(np.ascontiguousarray(sh.rowptr)).cuda()
d_edges = torch.from_numpy(np.ascontiguousarray(sh.edges).view(np.uint8)).cuda()
torch.cuda.synchronize()
for spec in sys.argv[2:]:
    ctx = G.LouvainGPU(0, 0, 1)
    for kv in spec.split(","):
        if kv:
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
    ctx.attach_device(nv, sh.parts, sh.lnv, sh.lne, d_rowptr.data_ptr(), d_edges.data_ptr())
    for r in range(3):
        mod, iters = ctx.louvain()
    tm = ctx.timings()
    st = ctx.scan_times() * 1e3
    print(f"{spec:40s} mod={mod:.17g} iters={iters} total={tm['total_s']*1e3:8.3f}ms scan={tm['scan_s']*1e3:8.3f}ms "
          f"avg={tm['scan_s']/iters*1e3:.3f} first3={st[0]:.2f},{st[1]:.2f},{st[2]:.2f} last={st[-1]:.2f} fold={tm['fold_s']*1e3:.2f} setup={tm['setup_s']*1e3:.2f} reorder={tm['reorder_s']*1e3:.2f}",
          flush=True)
    ctx.close()
