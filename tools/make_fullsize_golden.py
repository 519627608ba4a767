#!/usr/bin/env python3
"""Full-size golden trace from the UNMODIFIED reference (oracle/_ref) for a benchmark-size RGG.
Run on a box with enough cores/RAM (the GPU box): writes gpurun_out/golden_full_<nv>_p<strips>.json, which is
then committed under tests/golden/.  The graph comes from this repo's generator (validated byte-identical to the
reference generator at small sizes, tests/test_oracle.py); the reference reads it with -f.
usage: python tools/make_fullsize_golden.py NV STRIPS [RANKS] [THREADS] [PCT_RANDOM_EDGES]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minivite_b200 import hostgraph as hg  # noqa: E402
from oracle import oracle as O  # noqa: E402

nv, strips = int(sys.argv[1]), int(sys.argv[2])
ranks = int(sys.argv[3]) if len(sys.argv) > 3 else strips
threads = int(sys.argv[4]) if len(sys.argv) > 4 else max(1, len(os.sched_getaffinity(0)) // ranks)
pct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
t = time.time()
ss = hg.generate_rgg(nv, strips, random_edge_percent=pct)
path = os.path.join(tempfile.mkdtemp(prefix="mvfull_"), "g.bin")
ss.write(path)
ne = sum(s.lne for s in ss.shards)
ss.close()
print(f"graph nv={nv} strips={strips} ne={ne} written in {time.time()-t:.1f}s", flush=True)
t = time.time()
ref = O.run_reference(["-f", path], nranks=ranks, threads=threads, trace=True, arena_gb=200)
print(f"reference done in {time.time()-t:.1f}s: {ref['result']}", flush=True)
os.unlink(path)
out = {"nv": nv, "strips": strips, "random_edge_percent": pct, "ne": ref["result"]["ne"], "iters": ref["result"]["iters"],
       "modularity": ref["final"]["mod_repr"], "constant": repr(ref["final"]["constant"]),
       "final_chash": "%016x" % ref["final"]["chash"], "ref_ranks": ranks, "ref_threads": threads,
       "ref_time_s_with_trace_hooks": ref["result"]["time"],
       "trace": [{"modularity": x["mod_repr"], "moved": x["moved"], "chash": "%016x" % x["chash"]} for x in ref["trace"]]}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
fn = os.path.join(ROOT, "gpurun_out", f"golden_full_{nv}_p{strips}" + (f"_r{int(pct)}" if pct else "") + ".json")
json.dump(out, open(fn, "w"), indent=0)
print("wrote", fn)
