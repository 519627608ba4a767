#!/usr/bin/env python3
"""bench.py -- Louvain-phase throughput on synthetic RGGs (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref) on host cores

A "step" is one complete Louvain phase (the scope of the reference's timer, main.cpp:162-173: init +
ghost setup + all iterations until the modularity gain drops below 1e-6) over one synthetic RGG.
Workload at N GPUs: the graph `miniVite -n (16777216*N)` builds on N ranks (BASELINE.json configs[1] at
N=1; 16M vertices per GPU for N>1 => weak scaling), produced by this repo's exact fast generator.
metric = edges/s = (directed edge count) * iterations / t_louvain, whole job.
  value : graph already resident in HBM in the reference's own array format when the clock starts
  e2e   : host (pinned) arrays -> mvgpu_upload_shard (H2D) -> mvgpu_louvain -> assignment back to host
Timing: CUDA events on the library's stream (max over ranks) for `value`; inputs (3 GB/GPU) exceed L2 so
no explicit flush is needed between steps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NV_PER_GPU = 16777216


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.f.name):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if c[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_sample(nv_sample, steps, warmup, verbose=False):
    """Time oracle/_ref/miniVite_ref (the unmodified reference) on an RGG sample of the workload, using all
    host cores: both of its modes are tried (1 rank x all threads; P ranks x 1 thread on the P-strip graph)
    and each step uses the faster one.  Returns dict(value, ms_per_step, cores, kind, sample, runs)."""
    from minivite_b200 import hostgraph as hg
    from oracle import oracle as O
    cores = host_cores()
    if not O.have_reference():
        # port fallback: the C restatement, single thread
        ss = hg.generate_rgg(nv_sample, 1)
        sh = ss.shards[0]
        ts = []
        for k in range(warmup + steps):
            t = time.time()
            r = O.louvain(sh.parts, [sh.rowptr], [sh.edges])
            if k >= warmup:
                ts.append(time.time() - t)
        t = statistics.mean(ts)
        return {"value": sh.lne * r["iters"] / t, "ms_per_step": t * 1e3, "cores": 1, "kind": "port",
                "sample": f"RGG n={nv_sample} p=1 full Louvain phase, C restatement (oracle/_ref absent)",
                "unit": "edges/s"}
    tmp = tempfile.mkdtemp(prefix="mvbench_")
    modes = []
    pranks = 1
    while pranks * 2 <= min(cores, 64) and nv_sample % (pranks * 2) == 0:
        pranks *= 2
    for (p, thr) in ([(1, cores)] + ([(pranks, max(1, cores // pranks))] if pranks > 1 else [])):
        ss = hg.generate_rgg(nv_sample, p)
        path = os.path.join(tmp, f"s{p}.bin")
        ss.write(path)
        ne = sum(s.lne for s in ss.shards)
        ss.close()
        modes.append({"p": p, "thr": thr, "path": path, "ne": ne, "times": [], "iters": None})
    # one probing run per mode (doubles as warm-up), then only the faster mode is timed for the requested steps
    for m in modes:
        r = O.run_reference(["-f", m["path"]], nranks=m["p"], threads=m["thr"], trace=False)
        m["iters"] = r["result"]["iters"]
        m["probe"] = r["result"]["time"]
        m["eps"] = m["ne"] * m["iters"] / m["probe"]
        if verbose:
            print(f"# reference mode {m['p']} ranks x {m['thr']} threads (probe): {m['eps']:.4g} edges/s "
                  f"({m['probe']:.3f} s, {m['iters']} iters)", file=sys.stderr)
    fast = max(modes, key=lambda m: m["eps"])
    for k in range(max(warmup - 1, 0) + steps):
        r = O.run_reference(["-f", fast["path"]], nranks=fast["p"], threads=fast["thr"], trace=False)
        if k >= max(warmup - 1, 0):
            fast["times"].append(r["result"]["time"])
    for m in modes:
        if not m["times"]:
            m["times"] = [m["probe"]]
        m["eps"] = m["ne"] * m["iters"] / statistics.mean(m["times"])
    for m in modes:
        os.unlink(m["path"])
    os.rmdir(tmp)
    best = max(modes, key=lambda m: m["eps"])
    return {"value": best["eps"], "ms_per_step": statistics.mean(best["times"]) * 1e3, "cores": cores,
            "kind": "reference", "unit": "edges/s",
            "sample": (f"RGG n={nv_sample}, full Louvain phase (reference timer main.cpp:162-173), unmodified reference "
                       f"via oracle/_ref; fastest of " +
                       ", ".join(f"{m['p']} rank(s) x {m['thr']} thr = {m['eps']:.3g} e/s" for m in modes)),
            "runs": [{"ranks": m["p"], "threads": m["thr"], "edges_per_s": m["eps"], "iters": m["iters"]} for m in modes]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nv-per-gpu", type=int, default=NV_PER_GPU, help="dev knob; the benchmark config is the default")
    ap.add_argument("--cpu-sample-nv", type=int, default=2097152)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compact-upload", type=int, default=0, metavar="THREADS",
                    help="dev knob for the e2e leg: narrow unit-weight shards to 4-byte tails with THREADS host threads "
                         "before the H2D copy (library option compact_upload; default off)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.gpus
    # exactly ONE JSON line may reach stdout (libraries such as NCCL print banners there): park the real stdout
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:   # the host-side graph generator is OpenMP code: do not oversubscribe the cores across ranks
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_cores() // world))

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world != N and world != 1:
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {N}")
    nv_total = args.nv_per_gpu * N
    workload = f"RGG -n {nv_total} on {N} rank(s) (BASELINE.json configs[1] per GPU), unit weights, full Louvain phase"

    import __graft_entry__ as ge

    if args.impl == "reference":
        if rank != 0:
            return 0
        ge.build_host_only()
        r = reference_sample(args.cpu_sample_nv, args.steps, args.warmup, args.verbose)
        line = {"impl": "reference", "metric": "louvain_phase_edges_per_sec", "value": r["value"], "unit": "edges/s",
                "n_gpus": N, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload, "sample": r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    ge.build()
    from minivite_b200 import gpu as G
    from minivite_b200 import hostgraph as hg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- synthetic input: this rank's strip of the N-strip RGG (exact reference graph)
    hg.set_num_threads(max(1, host_cores() // max(world, 1)))
    t0 = time.time()
    ss = hg.generate_rgg(nv_total, N, rank, rank + 1)
    sh = ss.shards[0]
    gen_s = time.time() - t0
    ne_total = int(allsum(float(sh.lne)))
    parts = np.array([(nv_total * r) // N for r in range(N + 1)], dtype=np.int64)
    if args.verbose and rank == 0:
        print(f"# generated strip: lnv={sh.lnv} lne={sh.lne} in {gen_s:.1f}s", file=sys.stderr)

    ctx = G.LouvainGPU(local_rank, rank, N)
    ctx.set_option("host_threads", max(1, host_cores() // max(world, 1)))
    if N > 1:
        idt = torch.zeros(G.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(G.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()))

    # ---- value: inputs resident in HBM (reference array format) when the clock starts
    h_rowptr = torch.from_numpy(np.ascontiguousarray(sh.rowptr)).pin_memory()
    h_edges = torch.from_numpy(np.ascontiguousarray(sh.edges).view(np.uint8)).pin_memory()
    d_rowptr = h_rowptr.cuda(non_blocking=True)
    d_edges = h_edges.cuda(non_blocking=True)
    torch.cuda.synchronize()
    ctx.attach_device(nv_total, parts, sh.lnv, sh.lne, d_rowptr.data_ptr(), d_edges.data_ptr(), keepalive=(d_rowptr, d_edges))
    for _ in range(args.warmup):
        barrier()
        mod, iters = ctx.louvain()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    step_dev, step_wall, scan_s, scan_n, launches = [], [], 0.0, 0, 0
    for _ in range(args.steps):
        barrier()
        w0 = time.perf_counter()
        mod, iters = ctx.louvain()
        torch.cuda.synchronize()
        step_wall.append(time.perf_counter() - w0)
        tm = ctx.timings()
        step_dev.append(tm["total_s"])
        scan_s += tm["scan_s"]
        scan_n += tm["iters"]
        launches += tm["kernel_launches"]
    barrier()
    t_dev = allmax(sum(step_dev)) / args.steps          # device-timed (CUDA events), max over ranks
    t_wall = allmax(sum(step_wall)) / args.steps
    value = ne_total * iters / t_dev
    t_scan_iter = allmax(scan_s / max(scan_n, 1))       # avg duration of one scan launch, slowest rank
    tm_last = ctx.timings()
    info = ctx.shard_info()

    # ---- e2e: host arrays -> H2D -> Louvain -> assignment D2H, through the public API
    if args.compact_upload > 0:
        ctx.set_option("compact_upload", 1)
        ctx.set_option("host_threads", args.compact_upload)
    e2e_t = []
    h_comm = torch.empty(sh.lnv, dtype=torch.int64).pin_memory().numpy()      # the user's (pinned) result buffer
    for k in range(args.warmup + args.steps):
        barrier()
        w0 = time.perf_counter()
        ctx.upload(nv_total, parts, h_rowptr.numpy(), h_edges.numpy().view(hg.EDGE_DTYPE))
        m2, it2 = ctx.louvain()
        comm = ctx.communities(out=h_comm)
        torch.cuda.synchronize()
        if k >= args.warmup:
            e2e_t.append(time.perf_counter() - w0)
        tm_e2e = ctx.timings()
        h2d_bytes = tm_e2e["h2d_bytes"]
        assert it2 == iters and m2 == mod
    clocks = sampler.stop()                             # sampled across both timed regions (value + e2e steps)
    t_e2e = allmax(sum(e2e_t)) / args.steps
    e2e_value = ne_total * iters / t_e2e
    launches_total = int(allsum(float(launches)))
    h2d_total = int(allsum(float(h2d_bytes)))

    if rank != 0:
        return 0

    peak, peak_src = load_peaks()
    lnv, lne = sh.lnv, sh.lne
    b_alg = 24.0 * lne + 56.0 * lnv            # reference element sizes, per scan launch per GPU
    b_own = 8.0 * lne + 28.0 * lnv             # this build: 4 B tail + 4 B gathered id per edge; per vertex 4 rowptr
    #                                            + 4 cur + 8 cinfo + 4 tgt write + 8 (packed delta atomics, 2 x 57% ~ 1)
    roof = {"bound": "hbm", "achieved": b_alg / t_scan_iter / 1e9, "peak": peak, "unit": "GB/s",
            "frac": b_alg / t_scan_iter / 1e9 / peak, "traffic": None,
            "kernel": "k_scan_ws (neighbour scan)", "algorithmic_bytes_per_launch": b_alg,
            "own_layout_bytes_per_launch": b_own, "achieved_own_layout": b_own / t_scan_iter / 1e9,
            "avg_launch_ms": t_scan_iter * 1e3, "peak_source": peak_src,
            "whole_phase_gbs": b_alg * iters / t_dev / 1e9}
    prof = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(prof):
        try:
            roof["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    cpu = None
    if N == 1 and not args.no_cpu_baseline:
        r = reference_sample(args.cpu_sample_nv, 1, 1, args.verbose)
        cpu = {"value": r["value"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    line = {"metric": "louvain_phase_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": N,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "nv": nv_total, "ne": ne_total, "iterations": iters,
                       "modularity": mod, "s_per_iter": t_dev / iters, "l2": "inputs (3 GB/GPU) larger than L2; no flush",
                       "compact_upload_threads": args.compact_upload,
                       "unit_weight_path": bool(tm_last["unit_weight"]), "nghost": info["nghost"],
                       "arithmetic": "modularity gains in fp64 with the reference's rounding sequence; ids int32 on the device "
                                     "(int64 at the boundary); unit-weight degrees as exact integers",
                       "wall_ms_per_step": t_wall * 1e3, "graph_gen_s": gen_s},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "edges/s", "ms_per_step": t_e2e * 1e3,
                    "h2d_bytes_per_step": h2d_total,
                    "path": "pinned host arrays -> mvgpu_upload_shard -> mvgpu_louvain -> mvgpu_get_communities (pinned int64 result)",
                    "d2h_bytes_per_step": int(nv_total * 8 + 16)},
            "gpu_launches": launches_total, "clocks": clocks,
            "phase_ms": {"setup": tm_last["setup_s"] * 1e3, "scan": tm_last["scan_s"] * 1e3,
                         "fold": tm_last["fold_s"] * 1e3, "exchange": tm_last["exchange_s"] * 1e3,
                         "h2d_of_e2e_step": tm_e2e["h2d_s"] * 1e3}}
    emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
