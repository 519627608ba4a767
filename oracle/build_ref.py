#!/usr/bin/env python3
"""oracle/build_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Builds the *unmodified algorithm* of the miniVite reference (sources read where
they lie under /root/reference) into ``oracle/_ref/miniVite_ref`` using the
``oracle/mpi_shim/mpi.h`` stand-in for MPI.  Nothing from the reference is
copied into the repository: the sources are patched in a throw-away temp
directory (only *observation* hooks are injected, all gated by environment
variables, so an un-instrumented run executes exactly the reference's code)
and only the resulting binary lands in ``oracle/_ref/`` (git-ignored, but it
travels to the GPU box with the gpurun snapshot).

Hooks (all on stderr / side files, all off unless the env var is set):
  MV_TRACE=1          per iteration: ``ITER k mod=%.17g moved=N chash=HEX`` and at
                      the end ``FINAL prevMod=.. chashCurr=.. constant=.. ncomm=.. nprocs=..``
                      (dspl.hpp:1400 / dspl.hpp:1430; shard-combinable hash of SURVEY 8(c))
  MV_DUMP_COMM=pfx    rank r writes ``pfx.r``: int64 base, int64 nv, int64 currComm[nv]
  MV_DUMP_GRAPH=pfx   rank r writes ``pfx.r``: int64 base, lnv, lne, rowptr[lnv+1], Edge[lne]
                      right after graph creation (main.cpp:123)
  always              rank 0 prints ``RESULT mod=%.17g iters=%d time=%.9g nv=%ld ne=%ld nprocs=%d``
                      to stderr next to the reference's own report block (main.cpp:178)

``oracle/_ref/miniVite_ref32`` is the same source tree compiled with the reference's own ``-DUSE_32_BIT_GRAPH``
(utils.hpp:72-82); it pins the ``mvgpu_*32`` entry points (tests/golden/make_golden_32.py).

``--gpu`` additionally builds ``oracle/_ref/miniVite_ref_gpu``: the reference's own ``main.cpp`` with the patch of
INTEGRATION.md section 2 applied (the ``distLouvainMethod`` call replaced by ``mvgpu_upload_shard`` +
``mvgpu_louvain`` through the C ABI, everything else -- command line, graph construction, timer brackets, report --
the reference's code), linked against the shim and ``-lmvgpu``.  ``tests/test_gpu_reference_main.py`` runs it.

Usage: python oracle/build_ref.py [--ref /root/reference] [--force] [--gpu]
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
SHIM_DIR = os.path.join(HERE, "mpi_shim")

HASH_FN = r"""
static inline unsigned long long mv_mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31);
}
static inline unsigned long long mv_vhash(long long gid, long long val) {
  return mv_mix64(((unsigned long long)gid) * 0x9E3779B97F4A7C15ULL ^ (unsigned long long)val);
}
"""

HOOK_ITER = r"""    if (getenv("MV_TRACE")) { unsigned long long hl = 0, hg = 0; long long mv = 0, mvg = 0;
      for (GraphElem i = 0; i < nv; i++) { hl += mv_vhash(i + base, targetComm[i]); mv += (targetComm[i] != currComm[i]); }
      MPI_Allreduce(&hl, &hg, 1, MPI_INT64_T, MPI_SUM, gcomm); MPI_Allreduce(&mv, &mvg, 1, MPI_INT64_T, MPI_SUM, gcomm);
      if (me == 0) fprintf(stderr, "ITER %d mod=%.17g moved=%ld chash=%016llx\n", numIters, currMod, (long)mvg, hg); }
"""

HOOK_FINAL = r"""  iters = numIters;
  if (getenv("MV_TRACE")) { unsigned long long hl = 0, hg = 0; long long nc = 0, ncg = 0;
    for (GraphElem i = 0; i < nv; i++) { hl += mv_vhash(i + base, currComm[i]); }
    for (GraphElem i = 0; i < nv; i++) { nc += (localCinfo[i].size > 0); }
    MPI_Allreduce(&hl, &hg, 1, MPI_INT64_T, MPI_SUM, gcomm); MPI_Allreduce(&nc, &ncg, 1, MPI_INT64_T, MPI_SUM, gcomm);
    if (me == 0) fprintf(stderr, "FINAL prevMod=%.17g chashCurr=%016llx constant=%.17g ncomm_after_last=%ld nprocs=%d\n",
                         prevMod, hg, constantForSecondTerm, (long)ncg, nprocs); }
  if (getenv("MV_DUMP_COMM")) { char fn[4096]; snprintf(fn, sizeof fn, "%s.%d", getenv("MV_DUMP_COMM"), me);
    FILE *f = fopen(fn, "wb"); long long hdr[2] = {(long long)base, (long long)nv};
    fwrite(hdr, 8, 2, f); fwrite(currComm.data(), sizeof(GraphElem), nv, f); fclose(f); }
"""

HOOK_GRAPH = r"""  assert(g != nullptr);
  if (getenv("MV_DUMP_GRAPH")) { char fn[4096]; snprintf(fn, sizeof fn, "%s.%d", getenv("MV_DUMP_GRAPH"), me);
    FILE *f = fopen(fn, "wb"); long long hdr[3] = {(long long)g->get_base(me), (long long)g->get_lnv(), (long long)g->get_lne()};
    fwrite(hdr, 8, 3, f); fwrite(g->edge_indices_.data(), sizeof(GraphElem), g->get_lnv() + 1, f);
    fwrite(g->edge_list_.data(), sizeof(Edge), g->get_lne(), f); fclose(f); }
"""

HOOK_RESULT = r"""      double avgt = (tot_time / nprocs);
      fprintf(stderr, "RESULT mod=%.17g iters=%d time=%.9g nv=%ld ne=%ld nprocs=%d threads=%d\n", currMod, iters, avgt,
              (long)g->get_nv(), (long)g->get_ne(), nprocs, omp_get_max_threads());
"""


# ---- INTEGRATION.md section 2, as applied to a throw-away copy of the reference's main.cpp ------------------------------
GPU_INCLUDE = '#include "mvgpu.h"                       // this repo: include/mvgpu.h ; link with -lmvgpu\n'

GPU_CREATE = r"""  createCommunityMPIType();
  // one GPU per rank, communicator id via MPI (INTEGRATION.md section 2)
  unsigned char nccl_id[MVGPU_UNIQUE_ID_BYTES];
  if (me == 0 && nprocs > 1 && mvgpu_get_unique_id(nccl_id)) { std::cerr << mvgpu_last_error() << std::endl; MPI_Abort(MPI_COMM_WORLD, -99); }
  MPI_Bcast(nccl_id, sizeof nccl_id, MPI_BYTE, 0, MPI_COMM_WORLD);
  int ngpu = mvgpu_device_count();
  mvgpu_ctx *ctx = nullptr;
  if (ngpu < 1 || mvgpu_create(&ctx, me % ngpu, me, nprocs) || mvgpu_comm_init(ctx, nccl_id)) {
    std::cerr << mvgpu_last_error() << std::endl; MPI_Abort(MPI_COMM_WORLD, -99);
  }
  if (getenv("MV_TRACE")) mvgpu_set_option(ctx, "trace", 1);
"""

GPU_UPLOAD = r"""  size_t ssz = 0, rsz = 0;
  std::vector<GraphElem> parts(nprocs + 1);
  for (int r = 0; r <= nprocs; r++) parts[r] = (r < nprocs) ? g->get_base(r) : g->get_bound(nprocs - 1);
  if (mvgpu_upload_shard(ctx, g->get_nv(), parts.data(), g->get_lnv(), g->get_lne(),
                         g->edge_indices_.data(), g->edge_list_.data())) {
    std::cerr << mvgpu_last_error() << std::endl; MPI_Abort(MPI_COMM_WORLD, -99);
  }
"""

GPU_CALL = r"""  if (mvgpu_louvain(ctx, /*lower=*/currMod, threshold, &iters, &currMod)) {
    std::cerr << mvgpu_last_error() << std::endl; MPI_Abort(MPI_COMM_WORLD, -99);
  }
  if (getenv("MV_TRACE") && me == 0) {      // observation hook (same line format as the CPU build's)
    int n = 0; mvgpu_get_trace(ctx, 0, nullptr, &n);
    std::vector<mvgpu_iter_trace> tr(n);
    if (n) mvgpu_get_trace(ctx, n, tr.data(), &n);
    for (int k = 0; k < n; k++)
      fprintf(stderr, "ITER %d mod=%.17g moved=%ld chash=%016llx\n", k + 1, tr[k].modularity, (long)tr[k].moved,
              (unsigned long long)tr[k].chash);
    double cst = 0; mvgpu_get_constant(ctx, &cst);
    fprintf(stderr, "FINAL prevMod=%.17g chashCurr=%016llx constant=%.17g\n", currMod,
            (unsigned long long)(n >= 2 ? tr[n - 2].chash : 0), cst);
  }
"""

GPU_DESTROY = "  mvgpu_destroy(ctx);\n  delete g;\n"

REF_CALL = """#if defined(USE_MPI_RMA)
  currMod = distLouvainMethod(me, nprocs, *g, ssz, rsz, ssizes, rsizes, 
                svdata, rvdata, currMod, threshold, iters, commwin);
#else
  currMod = distLouvainMethod(me, nprocs, *g, ssz, rsz, ssizes, rsizes, 
                svdata, rvdata, currMod, threshold, iters);
#endif
"""


def patch_main_for_gpu(m):
    m = replace_once(m, '#include "dspl.hpp"\n', '#include "dspl.hpp"\n' + GPU_INCLUDE, "mvgpu.h include")
    m = replace_once(m, "  createCommunityMPIType();\n", GPU_CREATE, "context creation")
    m = replace_once(m, "  size_t ssz = 0, rsz = 0;\n", GPU_UPLOAD, "shard upload")
    m = replace_once(m, REF_CALL, GPU_CALL, "distLouvainMethod call")
    m = replace_once(m, "  delete g;\n", GPU_DESTROY, "context destruction")
    return m


def replace_once(text, old, new, what):
    if text.count(old) != 1:
        raise SystemExit(f"build_ref: anchor for {what!r} found {text.count(old)} times (expected 1)")
    return text.replace(old, new, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--gpu", action="store_true", help="also build miniVite_ref_gpu (reference main.cpp + INTEGRATION.md patch)")
    ap.add_argument("--no32", action="store_true", help="skip miniVite_ref32 (the reference compiled with -DUSE_32_BIT_GRAPH)")
    args = ap.parse_args()
    out_bin = os.path.join(OUT_DIR, "miniVite_ref")
    gpu_bin = os.path.join(OUT_DIR, "miniVite_ref_gpu")
    bin32 = os.path.join(OUT_DIR, "miniVite_ref32")
    repo = os.path.dirname(HERE)
    libdir = os.path.join(repo, "minivite_b200", "lib")
    if not os.path.isdir(args.ref):
        if os.path.exists(out_bin):
            print(f"build_ref: {args.ref} absent; keeping prebuilt {out_bin}")
            return 0
        print(f"build_ref: {args.ref} absent and no prebuilt binary", file=sys.stderr)
        return 1
    srcs = [os.path.join(args.ref, f) for f in ("main.cpp", "dspl.hpp", "graph.hpp", "utils.hpp")]
    stamp = max(os.path.getmtime(p) for p in srcs + [__file__, os.path.join(SHIM_DIR, "mpi.h")])
    want_gpu = args.gpu and os.path.exists(os.path.join(libdir, "libmvgpu.so"))
    gpu_stamp = max(stamp, os.path.getmtime(os.path.join(repo, "include", "mvgpu.h")))
    cpu_fresh = os.path.exists(out_bin) and os.path.getmtime(out_bin) >= stamp and (
        args.no32 or (os.path.exists(bin32) and os.path.getmtime(bin32) >= stamp))
    gpu_fresh = not want_gpu or (os.path.exists(gpu_bin) and os.path.getmtime(gpu_bin) >= gpu_stamp)
    if not args.force and cpu_fresh and gpu_fresh:
        print(f"build_ref: {out_bin} up to date")
        return 0
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="mvref_")
    try:
        for p in srcs:
            shutil.copy(p, tmp)
        d = open(os.path.join(tmp, "dspl.hpp")).read()
        d = replace_once(d, "static MPI_Datatype commType;\n", "static MPI_Datatype commType;\n" + HASH_FN, "hash fn")
        d = replace_once(d, "    // exit criteria\n    if (currMod - prevMod < thresh)",
                         HOOK_ITER + "    // exit criteria\n    if (currMod - prevMod < thresh)", "iteration hook")
        d = replace_once(d, "  iters = numIters;\n", HOOK_FINAL, "final hook")
        open(os.path.join(tmp, "dspl.hpp"), "w").write(d)
        m = open(os.path.join(tmp, "main.cpp")).read()
        m = replace_once(m, "  assert(g != nullptr);\n", HOOK_GRAPH, "graph dump hook")
        m = replace_once(m, "      double avgt = (tot_time / nprocs);\n", HOOK_RESULT, "result hook")
        open(os.path.join(tmp, "main.cpp"), "w").write(m)
        # The reference Makefile builds with -O3 -fopenmp -DPRINT_DIST_STATS (Makefile:10-16);
        # -std=c++11 implies -ffp-contract=off on x86-64 and no -march flag means no FMA.
        cmd = ["g++", "-std=c++11", "-O3", "-fopenmp", "-ffp-contract=off", "-DPRINT_DIST_STATS",
               "-I", SHIM_DIR, "-I", tmp, os.path.join(tmp, "main.cpp"), "-o", out_bin]
        print("build_ref:", " ".join(cmd))
        subprocess.check_call(cmd)
        if not args.no32:
            # the same sources with the reference's own 32-bit switch (utils.hpp:72-82: int32 ids, float weights)
            cmd32 = cmd[:5] + ["-DUSE_32_BIT_GRAPH"] + cmd[5:-1] + [bin32]
            print("build_ref:", " ".join(cmd32))
            subprocess.check_call(cmd32)
        if want_gpu:
            open(os.path.join(tmp, "main_gpu.cpp"), "w").write(patch_main_for_gpu(m))
            cmd = ["g++", "-std=c++11", "-O3", "-fopenmp", "-ffp-contract=off", "-DPRINT_DIST_STATS",
                   "-I", SHIM_DIR, "-I", tmp, "-I", os.path.join(repo, "include"), os.path.join(tmp, "main_gpu.cpp"),
                   "-o", gpu_bin, "-L", libdir, "-lmvgpu", "-Wl,-rpath,$ORIGIN/../../minivite_b200/lib"]
            print("build_ref:", " ".join(cmd))
            subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(f"build_ref: wrote {out_bin}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
