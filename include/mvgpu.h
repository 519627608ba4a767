/* mvgpu.h -- C ABI of the B200-native Louvain phase (drop-in for miniVite's distLouvainMethod).
 *
 * The reference has no FFI; its seam is one C++ call, main.cpp:168-169:
 *
 *     currMod = distLouvainMethod(me, nprocs, *g, ssz, rsz, ssizes, rsizes, svdata, rvdata,
 *                                 currMod, threshold, iters);            // dspl.hpp:1280-1283
 *
 * whose inputs are the arrays of `class Graph` (graph.hpp:85-296): `edge_indices_` (int64 local
 * offsets, lnv+1 entries), `edge_list_` ({int64 tail_; double weight_} = 16 B records with GLOBAL
 * tails), the partition `parts_` (nprocs+1 entries) and the scalars nv / lnv / lne.  The entry points
 * below take exactly those plain arrays -- no C++ or torch types cross the boundary -- and return
 * what the reference returns (modularity of the last accepted iteration, iteration count) plus the
 * final assignment, which the reference computes but never exports (dspl.hpp:1432-1438).
 *
 * One context == one rank == one GPU.  All functions return 0 on success; on failure they return a
 * non-zero code and mvgpu_last_error() describes it (no exceptions cross the boundary; the host
 * wrapper turns a failure into the reference's MPI_Abort(-99) behaviour).  A context must be driven
 * by one host thread at a time.  The library never falls back to a CPU implementation: without a
 * usable CUDA device every compute entry point fails.
 */
#ifndef MVGPU_H
#define MVGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvgpu_ctx mvgpu_ctx;

#define MVGPU_UNIQUE_ID_BYTES 128

/* Per-iteration record; mirrors the trace hook injected into the reference at dspl.hpp:1400. */
typedef struct {
  double modularity;   /* currMod of the iteration (dspl.hpp:447-448) */
  int64_t moved;       /* #{i : targetComm[i] != currComm[i]}, all ranks */
  uint64_t chash;      /* sum over vertices of mix64(gid*0x9E3779B97F4A7C15 ^ targetComm), mod 2^64, all ranks */
} mvgpu_iter_trace;

/* Device-side timing of the last mvgpu_louvain call (CUDA events on the context's stream). */
typedef struct {
  double total_s;        /* the reference's main.cpp:162-173 scope: setup + all iterations */
  double setup_s;        /* format conversion + ghost discovery (exchangeVertexReqs, dspl.hpp:1106-1272) + init (151-172) */
  double scan_s;         /* sum over iterations of the neighbour-scan kernel(s) (dspl.hpp:276-405) */
  double fold_s;         /* sum of fold + modularity partial kernels (dspl.hpp:458-471, 407-456) */
  double exchange_s;     /* sum of ghost exchange + collectives (dspl.hpp:488-952, 978-1103, 441) */
  double h2d_s;          /* host->device copy of the shard (mvgpu_upload_shard), not part of total_s */
  int64_t scan_launches; /* neighbour-scan kernel launches in the call */
  int64_t kernel_launches; /* all kernels of this library launched inside total_s */
  int32_t iters;
  int32_t unit_weight;   /* 1 if the integer fast path ran (all weights 1.0, 2m < 2^31) */
  double reorder_s;      /* part of setup_s spent on the locality renumbering (0 when it did not run) */
  int32_t reordered;     /* 1 if this rank's vertices were renumbered for locality */
  int32_t scan_kernel_chosen; /* scan_variant 6: the persistent kernel the run settled on (4 = k_scan_pw, 5 = k_scan_pq; 0: run too short) */
  int64_t h2d_bytes;     /* bytes mvgpu_upload_shard copied host->device for the current shard (0 for attached arrays) */
} mvgpu_timings;

const char *mvgpu_last_error(void);
/* Number of CUDA devices visible, or -1 (and an error string) when CUDA is unusable. */
int mvgpu_device_count(void);

/* ---- lifetime ------------------------------------------------------------------------------- */
/* Replaces the rank identity the reference takes from MPI_Comm_rank/size (main.cpp:97-98). */
int mvgpu_create(mvgpu_ctx **ctx, int device, int rank, int nranks);
int mvgpu_destroy(mvgpu_ctx *ctx);

/* ---- multi-GPU plumbing (nranks > 1 only) ---------------------------------------------------- */
/* Rank 0 creates the id and ships the 128 bytes to every rank by any means (torch.distributed
 * broadcast, a pipe, a file); every rank then calls mvgpu_comm_init.  Stands in for MPI_Init +
 * createCommunityMPIType (main.cpp:78-102). */
int mvgpu_get_unique_id(void *id128);
int mvgpu_comm_init(mvgpu_ctx *ctx, const void *id128);

/* ---- graph hand-off -------------------------------------------------------------------------- */
/* Host arrays of a reference Graph shard (graph.hpp:289-293): copies them to HBM.
 * parts: nranks+1 global vertex offsets (graph.hpp:109-113 or the -b bins, graph.hpp:511);
 * edge_indices: lnv+1 LOCAL offsets; edge_list: lne records {int64 tail; double weight}. */
int mvgpu_upload_shard(mvgpu_ctx *ctx, int64_t nv_global, const int64_t *parts, int64_t lnv, int64_t lne,
                       const int64_t *edge_indices, const void *edge_list);
/* Same, for arrays that already live in this context's device memory (not copied, not freed). */
int mvgpu_attach_shard_device(mvgpu_ctx *ctx, int64_t nv_global, const int64_t *parts, int64_t lnv, int64_t lne,
                              const int64_t *d_edge_indices, const void *d_edge_list);

/* Device-side GenerateRGG (graph.hpp:584-1213): builds this rank's strip of the graph
 * `miniVite -n nv_global` creates on nranks ranks directly in HBM, bit-identical to the reference generator, and
 * attaches it as the context's shard.  unit_weight = 0 gives Euclidean edge weights (-w).  *lne_out: local edges. */
int mvgpu_generate_rgg_shard(mvgpu_ctx *ctx, int64_t nv_global, int unit_weight, int64_t *lne_out);
/* Same with the remaining generator switches: lcg = 1 draws the coordinates from the reference's LCG class like
 * `miniVite -n nv_global -l` (utils.hpp:118-303, graph.hpp:703-729), lcg = 0 is the default engine;
 * random_edge_percent > 0 adds that percentage of random long edges like `-p` (graph.hpp:939-1122; the reference
 * seeds them from time and pid, graph.hpp:990, this library from a fixed seed -- the same one its host generator
 * uses, so both build the same graph).  With nranks > 1 and random_edge_percent > 0 the call is collective and needs
 * mvgpu_comm_init first (one count is summed over the ranks, graph.hpp:941-943). */
int mvgpu_generate_rgg_shard_ex(mvgpu_ctx *ctx, int64_t nv_global, int unit_weight, int lcg, double random_edge_percent,
                                int64_t *lne_out);
/* Copy the shard's reference-format arrays (lnv+1 offsets, lne 16-byte records) back to the host. */
int mvgpu_download_shard(mvgpu_ctx *ctx, int64_t *edge_indices, void *edge_list);

/* ---- the Louvain phase (dspl.hpp:1280-1441) --------------------------------------------------- */
/* lower/thresh as in the reference (main.cpp:149,70); *iters counts the rejected last iteration
 * (dspl.hpp:1430); *modularity is prevMod (dspl.hpp:1440).  Collective over all ranks. */
int mvgpu_louvain(mvgpu_ctx *ctx, double lower, double thresh, int *iters, double *modularity);

/* currComm of this rank's vertices at exit (global community ids), lnv entries. */
int mvgpu_get_communities(mvgpu_ctx *ctx, int64_t *out);
/* Same values left in device memory (int32 global ids); valid until the next mvgpu_louvain. */
int mvgpu_get_communities_device(mvgpu_ctx *ctx, const int32_t **d_out);

/* Options: "trace" (0/1: record moved/chash per iteration, default 0), "max_iters" (safety cap,
 * default 10000), "force_weighted" (0/1: use the fp64 path even for unit weights, default 0),
 * "force_heavy_deg" (test hook: treat vertices with degree > value as high-degree, default 0 = off),
 * "scan_variant" (6 (default): iterations 2 and 4 run k_scan_pw, iteration 3 k_scan_pq, and from iteration 5 on
 * k_scan_pq runs if it beat the geometric mean of its neighbours, else k_scan_pw; 5 = k_scan_pq: persistent warps fed by TMA bulk copies, boundary vertices reduced from a
 * per-warp ring -- unit weights; iteration 1 and weighted graphs run k_scan_pw; 4 = k_scan_pw throughout; 3 =
 * k_scan_ws: one CTA per 128-vertex tile, the default of round 1; identical results), "first_iter" (1 (default):
 * iteration 1 of a simple unit-weight graph uses the singleton-community reduction of k_scan_pw; 0: the general
 * reduction; identical results), "upload_chunk" (edges per chunk of the compact upload, default 4 Mi; test hook),
 * "cache_policy" (bit2 = L2 evict_first hint on the streamed arrays of k_scan_ws; other bits are accepted and
 * ignored; default 5), "reorder" (0 never, 1 always, 2 auto (default): renumber vertices for memory
 * locality when the given numbering has none -- layout only, results are identical), "region_size" (target
 * vertices per BFS region of the renumbering, default 512), "comm_mode" (multi-GPU per-iteration exchanges: 1 = stores /
 * flags in peer memory over NVLink (default), 0 = NCCL all-to-all-v + all-reduce), "host_transport" (set before
 * mvgpu_comm_init; 1: the setup-time exchanges go through a shared-memory segment of the host instead of NCCL --
 * required when several ranks share one device, which NCCL refuses; needs comm_mode 1; default 0), "compact_upload" (1:
 * mvgpu_upload_shard sends unit-weight shards as 4-byte tails narrowed on the host; 2: the copy engine additionally
 * takes raw chunks from the far end of the array whenever no narrowed chunk is ready (narrowed on the device);
 * 0 (default) = the 16-byte records as they are), "host_threads" (threads of that host pass, default 8).
 * The environment variable MVGPU_OPTIONS="name=value,name=value" presets options for every context of the process. */
/* In a multi-rank run every rank must set the same options (they change which collectives a run issues). */
int mvgpu_set_option(mvgpu_ctx *ctx, const char *name, int64_t value);
int mvgpu_get_trace(mvgpu_ctx *ctx, int max_entries, mvgpu_iter_trace *out, int *n);
int mvgpu_get_timings(mvgpu_ctx *ctx, mvgpu_timings *out);
/* Device time (seconds) of the neighbour-scan launch(es) of each iteration of the last run; *n = #iterations. */
int mvgpu_get_scan_times(mvgpu_ctx *ctx, int max_entries, double *out, int *n);
/* 1/(2m), the reference's constantForSecondTerm (dspl.hpp:129), of the last run. */
int mvgpu_get_constant(mvgpu_ctx *ctx, double *out);
/* Shard statistics after upload/louvain: info[0]=lnv, [1]=lne, [2]=nghost, [3]=send list length,
 * [4]=#high-degree vertices, [5]=max degree. */
int mvgpu_get_shard_info(mvgpu_ctx *ctx, int64_t *info6);

/* ---- one-call form of the reference seam ------------------------------------------------------ */
/* ---- the reference's USE_32_BIT_GRAPH build (utils.hpp:72-82: GraphElem = int32_t, GraphWeight = float) ------------
 * A miniVite compiled with -DUSE_32_BIT_GRAPH holds int32 offsets and 8-byte {int32 tail_; float weight_} records
 * (graph.hpp:60-66 with those types) and computes gains, 1/(2m) and the modularity in float.  These entry points take
 * that build's arrays and reproduce that build's arithmetic: the gain is rounded where the float build rounds it
 * (see gain_of in kernels.cuh), 1/(2m), the modularity and the exit test are evaluated in float (dspl.hpp:129,
 * 447-448, 1401).  For unit-weight graphs whose integer sums stay below 2^24 every float sum of the reference is
 * exact, and the results -- assignment, iteration count, modularity -- are bit-identical to the float build's;
 * beyond that the float build is not reproducible against itself (its OpenMP reductions round in arbitrary order)
 * and the comparison is by tolerance.  parts / edge_indices are int32 here, like every GraphElem of that build. */
int mvgpu_upload_shard32(mvgpu_ctx *ctx, int32_t nv_global, const int32_t *parts, int32_t lnv, int32_t lne,
                         const int32_t *edge_indices, const void *edge_list8);
int mvgpu_louvain32(mvgpu_ctx *ctx, float lower, float thresh, int *iters, float *modularity);
int mvgpu_get_communities32(mvgpu_ctx *ctx, int32_t *out);

/* distLouvainMethod(me, nprocs, g, ..., lower, thresh, iters) for a single-GPU run with HOST arrays:
 * create + upload + louvain (+ optional assignment download into comm_out, may be NULL) + destroy. */
int mvgpu_dist_louvain_method(int device, int64_t nv, int64_t ne_local, const int64_t *edge_indices,
                              const void *edge_list, double lower, double thresh, int *iters,
                              double *modularity, int64_t *comm_out);

#ifdef __cplusplus
}
#endif
#endif /* MVGPU_H */
