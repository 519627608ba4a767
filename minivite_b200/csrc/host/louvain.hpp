// Host replacement for the reference's distLouvainMethod (dspl.hpp:1280-1283): same parameter list,
// same return value / `iters` semantics; the body marshals the Graph's arrays across the C ABI
// (include/mvgpu.h) to the CUDA library instead of running the OpenMP/MPI loops.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "graph.hpp"
#include "mvgpu.h"

// What MPI_Comm carried in the reference: which GPU to use and (for nprocs > 1) the communicator id
// every rank received from rank 0.
struct GpuRankContext {
  int device = 0;
  unsigned char unique_id[MVGPU_UNIQUE_ID_BYTES] = {0};
  bool trace = false;
  std::vector<GraphElem> *comm_out = nullptr;     // final currComm of this rank (the reference drops it, dspl.hpp:1432-1438)
  mvgpu_timings timings{};
  std::vector<mvgpu_iter_trace> iter_trace;
};

[[noreturn]] inline void mv_abort(const std::string &where) {
  // the reference's only error behaviour on this path is MPI_Abort(comm, -99) (main.cpp:90,251-276)
  std::fprintf(stderr, "[miniVite_b200] %s: %s\n", where.c_str(), mvgpu_last_error());
  std::exit(99);
}

// Developer knobs, not part of the reference's surface: MVGPU_OPTIONS="name=value,name=value" is handed to
// mvgpu_set_option (unknown names abort), MVGPU_REPEAT=n runs the phase n times (same result; timings of the last,
// warm, run are reported).
inline void mv_apply_env_options(mvgpu_ctx *ctx) {
  const char *e = std::getenv("MVGPU_OPTIONS");
  if (!e) return;
  std::string s(e);
  size_t pos = 0;
  while (pos < s.size()) {
    size_t end = s.find(',', pos);
    if (end == std::string::npos) end = s.size();
    const std::string kv = s.substr(pos, end - pos);
    pos = end + 1;
    if (kv.empty()) continue;
    const size_t eq = kv.find('=');
    if (eq == std::string::npos) { std::fprintf(stderr, "[miniVite_b200] MVGPU_OPTIONS: missing '=' in %s\n", kv.c_str()); std::exit(99); }
    if (mvgpu_set_option(ctx, kv.substr(0, eq).c_str(), std::atoll(kv.c_str() + eq + 1))) mv_abort("MVGPU_OPTIONS");
  }
}
inline int mv_env_repeat() {
  const char *e = std::getenv("MVGPU_REPEAT");
  const int n = e ? std::atoi(e) : 1;
  return n < 1 ? 1 : n;
}

// MVGPU_SCAN_TIMES=1: print the duration of every scan launch of the last phase (ms) on stderr
inline void mv_print_scan_times(mvgpu_ctx *ctx, int me) {
  if (me != 0 || !std::getenv("MVGPU_SCAN_TIMES")) return;
  int n = 0;
  mvgpu_get_scan_times(ctx, 0, nullptr, &n);
  std::vector<double> t(n);
  if (n) mvgpu_get_scan_times(ctx, n, t.data(), &n);
  std::fprintf(stderr, "SCAN_MS");
  for (double x : t) std::fprintf(stderr, " %.3f", x * 1e3);
  std::fprintf(stderr, "\n");
}

// The 6 scratch parameters (ssz ... rvdata) are caller-owned out-params that only the reference's
// exchangeVertexReqs filled (dspl.hpp:1106-1272); the ghost lists now live in device memory, so they
// are left empty.
inline GraphWeight distLouvainMethod(const int me, const int nprocs, const Graph &dg, size_t &ssz, size_t &rsz,
                                     std::vector<GraphElem> &ssizes, std::vector<GraphElem> &rsizes,
                                     std::vector<GraphElem> &svdata, std::vector<GraphElem> &rvdata,
                                     const GraphWeight lower, const GraphWeight thresh, int &iters,
                                     GpuRankContext &rc) {
  ssz = rsz = 0;
  ssizes.clear(); rsizes.clear(); svdata.clear(); rvdata.clear();
  mvgpu_ctx *ctx = nullptr;
  if (mvgpu_create(&ctx, rc.device, me, nprocs)) mv_abort("mvgpu_create");
  mv_apply_env_options(ctx);                       // before the communicator: host_transport decides how it is built
  if (nprocs > 1 && mvgpu_comm_init(ctx, rc.unique_id)) mv_abort("mvgpu_comm_init");
  if (rc.trace) mvgpu_set_option(ctx, "trace", 1);
  if (mvgpu_upload_shard(ctx, dg.get_nv(), dg.parts().data(), dg.get_lnv(), dg.get_lne(), dg.edge_indices_.data(),
                         dg.edge_list_.data()))
    mv_abort("mvgpu_upload_shard");
  double mod = 0.0;
  for (int rep = mv_env_repeat(); rep > 0; rep--)
    if (mvgpu_louvain(ctx, lower, thresh, &iters, &mod)) mv_abort("mvgpu_louvain");
  mvgpu_get_timings(ctx, &rc.timings);
  mv_print_scan_times(ctx, me);
  if (rc.trace) {
    int n = 0;
    mvgpu_get_trace(ctx, 0, nullptr, &n);
    rc.iter_trace.resize(n);
    if (n) mvgpu_get_trace(ctx, n, rc.iter_trace.data(), &n);
  }
  if (rc.comm_out) {
    rc.comm_out->resize(dg.get_lnv());
    if (mvgpu_get_communities(ctx, rc.comm_out->data())) mv_abort("mvgpu_get_communities");
  }
  mvgpu_destroy(ctx);
  return mod;
}

// Variant for `-D`: the shard never exists on the host -- GenerateRGG runs on the device (mvgpu_generate_rgg_shard,
// the reference's generator of graph.hpp:584-1213 bit for bit) and the Louvain phase consumes it in place.
// `gen_seconds` receives the generation time, `lne` the local edge count.
inline GraphWeight distLouvainMethodOnDeviceRGG(const int me, const int nprocs, const GraphElem nv, const bool unitEdgeWeight,
                                                const bool lcg, const GraphWeight randomEdgePercent, const GraphWeight lower, const GraphWeight thresh, int &iters,
                                                GpuRankContext &rc, GraphElem &lne, double &gen_seconds,
                                                const std::function<void()> &before_louvain) {
  mvgpu_ctx *ctx = nullptr;
  if (mvgpu_create(&ctx, rc.device, me, nprocs)) mv_abort("mvgpu_create");
  mv_apply_env_options(ctx);                       // before the communicator: host_transport decides how it is built
  if (nprocs > 1 && mvgpu_comm_init(ctx, rc.unique_id)) mv_abort("mvgpu_comm_init");
  if (rc.trace) mvgpu_set_option(ctx, "trace", 1);
  const auto t0 = std::chrono::steady_clock::now();
  int64_t lne64 = 0;
  if (mvgpu_generate_rgg_shard_ex(ctx, nv, unitEdgeWeight ? 1 : 0, lcg ? 1 : 0, randomEdgePercent, &lne64)) mv_abort("mvgpu_generate_rgg_shard");
  gen_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  lne = lne64;
  before_louvain();
  double mod = 0.0;
  for (int rep = mv_env_repeat(); rep > 0; rep--)
    if (mvgpu_louvain(ctx, lower, thresh, &iters, &mod)) mv_abort("mvgpu_louvain");
  mvgpu_get_timings(ctx, &rc.timings);
  mv_print_scan_times(ctx, me);
  if (rc.trace) {
    int n = 0;
    mvgpu_get_trace(ctx, 0, nullptr, &n);
    rc.iter_trace.resize(n);
    if (n) mvgpu_get_trace(ctx, n, rc.iter_trace.data(), &n);
  }
  if (rc.comm_out) {
    rc.comm_out->resize(nv / nprocs);
    if (mvgpu_get_communities(ctx, rc.comm_out->data())) mv_abort("mvgpu_get_communities");
  }
  mvgpu_destroy(ctx);
  return mod;
}
