// libmvhost.so: C entry points over the host-side graph tools (RGG generator,
// binary graph file reader/writer) so Python tests / bench.py can produce the
// same inputs the C++ driver does.  No CUDA here.
#include <cstring>
#include <string>

#include "binio.hpp"
#include "graph.hpp"
#include "rgg.hpp"

namespace {
thread_local std::string g_err;
struct ShardSet { std::vector<Graph *> shards; };
template <typename F> int guarded(F &&f) {
  try { f(); return 0; }
  catch (const std::exception &e) { g_err = e.what(); return 1; }
  catch (...) { g_err = "unknown error"; return 1; }
}
}  // namespace

extern "C" {

const char *mvh_last_error() { return g_err.c_str(); }

// OpenMP width of the host tools (launchers such as torchrun export OMP_NUM_THREADS=1)
void mvh_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int mvh_get_max_threads() { return omp_get_max_threads(); }

// Shards [r_begin, r_end) of the nprocs-strip RGG; r_end < 0 means all.
int mvh_rgg_generate(int64_t nv, int nprocs, int r_begin, int r_end, int is_lcg, int unit_weight,
                     double random_edge_percent, void **out) {
  return guarded([&] {
    mvhost::GenerateRGG gr(nv, nprocs);
    ShardSet *s = new ShardSet;
    s->shards = gr.generate(is_lcg != 0, unit_weight != 0, random_edge_percent, r_begin, r_end);
    *out = s;
  });
}

// utils.hpp:91-98 and the coordinates rank r of the reference draws (graph.hpp:680-700 / LCG utils.hpp:118-303)
int64_t mvh_reseeder(unsigned initseed) { return (int64_t)mvhost::reseeder(initseed); }
int mvh_rgg_points(int64_t nv, int nprocs, int r, int is_lcg, int64_t count, double *x, double *y) {
  return guarded([&] {
    mvhost::GenerateRGG gr(nv, nprocs);
    std::vector<GraphWeight> X, Y;
    gr.strip_points(r, is_lcg != 0, X, Y);
    for (int64_t i = 0; i < count && i < (int64_t)X.size(); i++) { x[i] = X[i]; y[i] = Y[i]; }
  });
}

double mvh_rgg_radius(int64_t nv, int nprocs) {
  try { return mvhost::GenerateRGG(nv, nprocs).get_d(); } catch (...) { return -1.0; }
}

int mvh_graph_read(const char *path, int me, int nprocs, int balanced, void **out) {
  return guarded([&] {
    mvhost::BinaryEdgeList rm;
    ShardSet *s = new ShardSet;
    s->shards.push_back(balanced ? rm.read_balanced(me, nprocs, 1, path) : rm.read(me, nprocs, 1, path));
    *out = s;
  });
}

int mvh_graph_count(void *h) { return (int)((ShardSet *)h)->shards.size(); }

// info[0..5] = base, bound, lnv, lne, nv, ne ; parts_out (size+1 entries) optional
int mvh_graph_shard(void *h, int idx, int64_t *info, const int64_t **rowptr, const void **edges,
                    const int64_t **parts, int *nparts) {
  return guarded([&] {
    ShardSet *s = (ShardSet *)h;
    if (idx < 0 || idx >= (int)s->shards.size()) throw std::out_of_range("shard index");
    const Graph *g = s->shards[idx];
    info[0] = g->get_base(g->get_rank());
    info[1] = g->get_bound(g->get_rank());
    info[2] = g->get_lnv();
    info[3] = g->get_lne();
    info[4] = g->get_nv();
    info[5] = g->get_ne();
    *rowptr = g->edge_indices_.data();
    *edges = g->edge_list_.data();
    if (parts) *parts = g->parts().data();
    if (nparts) *nparts = (int)g->parts().size();
  });
}

int mvh_graph_write(void *h, const char *path) {
  return guarded([&] { mvhost::BinaryEdgeList::write(path, ((ShardSet *)h)->shards); });
}

void mvh_graph_free(void *h) {
  ShardSet *s = (ShardSet *)h;
  if (!s) return;
  for (Graph *g : s->shards) delete g;
  delete s;
}

}  // extern "C"
