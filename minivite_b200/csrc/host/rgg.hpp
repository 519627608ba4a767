// Exact, fast random-geometric-graph generator (host side, OpenMP).
//
// Produces *the same graph* as the reference's GenerateRGG (graph.hpp:584-1213)
// run on `nprocs` ranks -- same RNG stream, same radius, same fp64 predicate,
// same strip layout and vertex numbering, same cross-strip quirk -- but in
// O(n) expected time with a uniform cell grid instead of the reference's
// O((n/p)^2) all-pairs loops (graph.hpp:759-760, 816-817, 848-849), which would
// need ~75 h for the 16M-vertex benchmark graph.  This is a measurement
// prerequisite, not part of the Louvain hot path.
//
// What "the same graph" rests on:
//   * coordinates: std::default_random_engine seeded with reseeder(1), one
//     uniform_real_distribution<double> draw for X in (0,1) then one for Y in
//     (lo,hi) per vertex, EVERY strip restarting from the same seed
//     (graph.hpp:680-700, utils.hpp:91-114); `-l` uses the 2x2-matrix LCG of
//     utils.hpp:118-303 including its wrap-around int64 arithmetic;
//   * radius: rn = (sqrt(ln(nv)/(3.14159 nv)) + sqrt(2.0736/nv))/2 (graph.hpp:629-631);
//   * predicate: sqrt(dx*dx + dy*dy) <= rn in fp64 without FMA contraction
//     (graph.hpp:763-767); it is symmetric in (i,j), so pair-discovery order is irrelevant;
//   * cross-strip pairs are discovered only for different local indices
//     (graph.hpp:817,849 start at j=i+1 in both directions) and only between adjacent strips;
//   * adjacency of each vertex sorted by global tail id (graph.hpp:1145-1153).
// Validated byte-for-byte against the reference generator in tests (via oracle/_ref dumps).
#pragma once
#include <omp.h>

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "graph.hpp"

#define MV_PI (3.14159)   // the reference's truncated constant (utils.hpp:44); part of the radius definition

namespace mvhost {

// utils.hpp:91-98: one 32-bit word out of seed_seq{initseed}.
inline GraphElem reseeder(unsigned initseed) {
  std::seed_seq seq({initseed});
  std::vector<std::uint32_t> seeds(1);
  seq.generate(seeds.begin(), seeds.end());
  return (GraphElem)seeds[0];
}

inline bool is_pwr2(int n) { return n != 0 && !(n & (n - 1)); }

class GenerateRGG {
 public:
  // Fixed seed for the `-p` random long edges: the reference seeds them with
  // time(0)^getpid() (graph.hpp:990), i.e. irreproducibly; we document a constant instead.
  static constexpr unsigned kRandomEdgeSeed = 20180912u;

  GenerateRGG(GraphElem nv, int nprocs) : nv_(nv), nprocs_(nprocs) {
    if (nprocs_ < 1 || nv_ < 1) throw std::invalid_argument("bad RGG size");
    if (nv_ % nprocs_ != 0)
      throw std::invalid_argument("[ERROR] Number of vertices must be perfectly divisible by number of processes.");
    if (!is_pwr2(nprocs_)) throw std::invalid_argument("[ERROR] Number of processes must be a power of 2.");
    n_ = nv_ / nprocs_;
    const GraphWeight rc = std::sqrt((GraphWeight)std::log((double)nv_) / (GraphWeight)(MV_PI * nv_));
    const GraphWeight rt = std::sqrt((GraphWeight)2.0736 / (GraphWeight)nv_);
    rn_ = (rc + rt) / (GraphWeight)2.0;
    if (!(((GraphWeight)1.0 / (GraphWeight)nprocs_) > rn_))
      throw std::invalid_argument("RGG radius does not fit the strip height (1/p > rn violated)");
  }

  GraphWeight get_d() const { return rn_; }
  GraphElem get_nv() const { return nv_; }

  // Coordinates of strip `r` exactly as rank r of the reference would draw them.
  void strip_points(int r, bool isLCG, std::vector<GraphWeight> &X, std::vector<GraphWeight> &Y) const {
    X.resize(n_);
    Y.resize(n_);
    const GraphWeight rec_np = (GraphWeight)(1.0 / (GraphWeight)nprocs_);
    const GraphWeight lo = r * rec_np;
    const GraphWeight hi = lo + rec_np;
    if (!isLCG) {
      std::default_random_engine gen((unsigned)reseeder(1));
      std::uniform_real_distribution<GraphWeight> utd;
      using P = std::uniform_real_distribution<GraphWeight>::param_type;
      for (GraphElem i = 0; i < n_; i++) {
        X[i] = utd(gen, P{0.0, 1.0});
        Y[i] = utd(gen, P{lo, hi});
      }
      return;
    }
    // LCG path: rank r owns entries [2 n_ r, 2 n_ (r+1)) of x[k] = 16807 x[k-1] mod (2^31-1);
    // the first one comes from a 2x2 matrix power evaluated in wrapping int64 (utils.hpp:146-218).
    const int64_t M = 2147483647LL, A = 16807LL;
    const int64_t x0 = reseeder(1);
    const GraphElem len = 2 * n_;
    int64_t first;
    if (r == 0) first = x0;
    else {
      uint64_t acc = 1, base = (uint64_t)A, e = (uint64_t)len * (uint64_t)r;
      while (e) { if (e & 1) acc *= base; base *= base; e >>= 1; }   // wrapping power, ring-equivalent to the loop
      first = (int64_t)((uint64_t)x0 * acc) % M;
    }
    const GraphWeight mult = 1.0 / (GraphWeight)(1.0 + (GraphWeight)(M - 1));
    int64_t x = first;
    for (GraphElem i = 0; i < len; i++) {
      if (i) x = (x * A) % M;
      const GraphWeight d = (GraphWeight)std::fabs((GraphWeight)x) * mult;
      if (i < n_) X[i] = d;
      else Y[i - n_] = lo + (GraphWeight)(rec_np * d);
    }
  }

  // Builds the CSR shards of ranks [r_begin, r_end).  With randomEdgePercent > 0 all
  // ranks have to be built together (random edges insert reverse edges on other ranks).
  std::vector<Graph *> generate(bool isLCG, bool unitEdgeWeight = true, GraphWeight randomEdgePercent = 0.0,
                                int r_begin = 0, int r_end = -1, unsigned rande_seed = kRandomEdgeSeed) const {
    if (r_end < 0) r_end = nprocs_;
    if (randomEdgePercent > 0.0 && (r_begin != 0 || r_end != nprocs_))
      throw std::invalid_argument("random edges need all shards generated together");
    std::vector<std::vector<GraphWeight>> PX(nprocs_), PY(nprocs_);
    const int s0 = std::max(0, r_begin - 1), s1 = std::min(nprocs_, r_end + 1);
    if (!isLCG) {
      // every strip draws the same X and the same canonical Y; still go through the
      // distribution object per strip so that lo + (hi-lo)*u rounds exactly like libstdc++ does.
#pragma omp parallel for schedule(dynamic, 1)
      for (int s = s0; s < s1; s++) strip_points(s, false, PX[s], PY[s]);
    } else {
#pragma omp parallel for schedule(dynamic, 1)
      for (int s = s0; s < s1; s++) strip_points(s, true, PX[s], PY[s]);
    }

    std::vector<Adj> adj(nprocs_);
    for (int r = r_begin; r < r_end; r++) build_strip(r, PX, PY, unitEdgeWeight, adj[r]);
    if (randomEdgePercent > 0.0) add_random_edges(adj, PX, PY, unitEdgeWeight, randomEdgePercent, rande_seed);

    GraphElem ne_global = -1;
    if (r_begin == 0 && r_end == nprocs_) {
      ne_global = 0;
      for (int r = 0; r < nprocs_; r++) ne_global += (GraphElem)adj[r].edges.size();
    }
    std::vector<Graph *> out;
    for (int r = r_begin; r < r_end; r++) {
      Graph *g = new Graph(n_, 0, nv_, ne_global, r, nprocs_);
      g->edge_indices_.swap(adj[r].rowptr);
      g->edge_list_.swap(adj[r].edges);
      g->set_nedges((GraphElem)g->edge_list_.size(), ne_global);
      out.push_back(g);
    }
    return out;
  }

 private:
  struct Adj {
    std::vector<GraphElem> rowptr;
    std::vector<Edge> edges;
  };

  // Neighbour search for all points of strip r against strips r-1, r, r+1.
  void build_strip(int r, const std::vector<std::vector<GraphWeight>> &PX,
                   const std::vector<std::vector<GraphWeight>> &PY, bool unit, Adj &out) const {
    // grid: square cells of width >= rn over the slab of rows this strip can reach
    int64_t ncell = (int64_t)std::floor(1.0 / rn_);
    while (ncell > 1 && 1.0 / (double)ncell < rn_ * 1.000001) ncell--;
    if (ncell < 1) ncell = 1;
    const GraphWeight rec_np = (GraphWeight)(1.0 / (GraphWeight)nprocs_);
    const double ylo = r * rec_np - rn_ * 1.01, yhi = (r + 1) * rec_np + rn_ * 1.01;
    auto cell_of = [&](double v) -> int64_t {
      int64_t c = (int64_t)std::floor(v * (double)ncell);
      return c < 0 ? 0 : (c >= ncell ? ncell - 1 : c);
    };
    const int64_t row0 = std::max<int64_t>(0, cell_of(ylo) - 1), row1 = std::min<int64_t>(ncell - 1, cell_of(yhi) + 1);
    const int64_t nrows = row1 - row0 + 1, ncells = nrows * ncell;

    struct Cand { GraphWeight x, y; GraphElem gid; };
    const int s0 = std::max(0, r - 1), s1 = std::min(nprocs_ - 1, r + 1);
    std::vector<int64_t> cstart(ncells + 1, 0);
    auto slab_cell = [&](double x, double y) -> int64_t {
      const int64_t row = cell_of(y);
      if (row < row0 || row > row1) return -1;
      return (row - row0) * ncell + cell_of(x);
    };
    for (int s = s0; s <= s1; s++)
      for (GraphElem i = 0; i < n_; i++) {
        if (PY[s][i] < ylo || PY[s][i] > yhi) continue;
        const int64_t c = slab_cell(PX[s][i], PY[s][i]);
        if (c >= 0) cstart[c + 1]++;
      }
    for (int64_t c = 0; c < ncells; c++) cstart[c + 1] += cstart[c];
    std::vector<Cand> cand(cstart[ncells]);
    {
      std::vector<int64_t> fill(cstart.begin(), cstart.end() - 1);
      for (int s = s0; s <= s1; s++)
        for (GraphElem i = 0; i < n_; i++) {
          if (PY[s][i] < ylo || PY[s][i] > yhi) continue;
          const int64_t c = slab_cell(PX[s][i], PY[s][i]);
          if (c >= 0) cand[fill[c]++] = Cand{PX[s][i], PY[s][i], (GraphElem)s * n_ + i};
        }
    }

    const GraphElem base = (GraphElem)r * n_;
    const std::vector<GraphWeight> &X = PX[r], &Y = PY[r];
    const GraphWeight rn = rn_;
    // visit(i, f): call f(gid, ed) for every neighbour of local vertex i
    auto visit = [&](GraphElem i, auto &&f) {
      const GraphWeight xi = X[i], yi = Y[i];
      const int64_t row = cell_of(yi), col = cell_of(xi);
      const int64_t c0 = std::max<int64_t>(0, col - 1), c1 = std::min<int64_t>(ncell - 1, col + 1);
      for (int64_t rr = std::max(row0, row - 1); rr <= std::min(row1, row + 1); rr++) {
        const int64_t b = cstart[(rr - row0) * ncell + c0], e = cstart[(rr - row0) * ncell + c1 + 1];
        for (int64_t k = b; k < e; k++) {
          const Cand &q = cand[k];
          const GraphWeight dx = xi - q.x, dy = yi - q.y;
          const GraphWeight ed = std::sqrt(dx * dx + dy * dy);
          if (!(ed <= rn)) continue;
          const GraphElem ql = q.gid % n_;
          const int qs = (int)(q.gid / n_);
          if (qs == r) { if (ql == i) continue; }          // no self pairs
          else if (ql == i) continue;                      // cross-strip equal-index pairs are never tested
          f(q.gid, ed);
        }
      }
    };

    out.rowptr.assign(n_ + 1, 0);
#pragma omp parallel for schedule(dynamic, 4096)
    for (GraphElem i = 0; i < n_; i++) {
      GraphElem d = 0;
      visit(i, [&](GraphElem, GraphWeight) { d++; });
      out.rowptr[i + 1] = d;
    }
    for (GraphElem i = 0; i < n_; i++) out.rowptr[i + 1] += out.rowptr[i];
    out.edges.resize(out.rowptr[n_]);
#pragma omp parallel for schedule(dynamic, 4096)
    for (GraphElem i = 0; i < n_; i++) {
      Edge *dst = out.edges.data() + out.rowptr[i];
      GraphElem d = 0;
      visit(i, [&](GraphElem gid, GraphWeight ed) {
        dst[d].tail_ = gid;
        dst[d].weight_ = unit ? 1.0 : ed;
        d++;
      });
      std::sort(dst, dst + d, [](const Edge &a, const Edge &b) { return a.tail_ < b.tail_; });
    }
    (void)base;
  }

  // graph.hpp:939-1122 with a fixed seed (rank r uses rande_seed + r).
  void add_random_edges(std::vector<Adj> &adj, const std::vector<std::vector<GraphWeight>> &PX,
                        const std::vector<std::vector<GraphWeight>> &PY, bool unit, GraphWeight pct,
                        unsigned rande_seed) const {
    GraphElem tot_pnedges = 0;
    for (int r = 0; r < nprocs_; r++) tot_pnedges += (GraphElem)(adj[r].edges.size() / 2);
    const GraphElem nrande = (((GraphElem)(pct * (GraphWeight)tot_pnedges)) / 100);
    struct Extra { GraphElem src_local; Edge e; };
    std::vector<std::vector<Extra>> extra(nprocs_);
    for (int r = 0; r < nprocs_; r++) {
      GraphElem pnrande = 0;
      if (nrande < nprocs_) { if (r == nprocs_ - 1) pnrande += nrande; }
      else {
        pnrande = nrande / nprocs_;
        if (r == nprocs_ - 1) pnrande += nrande % nprocs_;
      }
      std::default_random_engine re(rande_seed + (unsigned)r);
      std::uniform_int_distribution<GraphElem> IR, JR;
      std::uniform_real_distribution<GraphWeight> IJW;
      std::hash<GraphElem> reh;
      std::unordered_set<uint64_t> added;   // forward (i -> g_j) random edges of this rank so far
      const Adj &A = adj[r];
      for (GraphElem k = 0; k < pnrande; k++) {
        const GraphElem i = IR(re, std::uniform_int_distribution<GraphElem>::param_type{0, n_ - 1});
        const GraphElem g_j = JR(re, std::uniform_int_distribution<GraphElem>::param_type{0, nv_ - 1});
        const int target = (int)(g_j / n_);
        const GraphElem j = g_j - (GraphElem)target * n_;
        if (i == j) continue;
        const GraphElem g_i = (GraphElem)r * n_ + i;
        // duplicate check against this rank's current list: RGG adjacency + its own forward random edges
        const Edge *b = A.edges.data() + A.rowptr[i], *e = A.edges.data() + A.rowptr[i + 1];
        const bool in_rgg = std::binary_search(b, e, g_j, EdgeTailLess());
        const uint64_t key = (uint64_t)i * (uint64_t)nv_ + (uint64_t)g_j;
        if (in_rgg || added.count(key)) continue;
        GraphWeight weight = 1.0;
        if (!unit) {
          if (target == r || target == r - 1 || target == r + 1) {
            const GraphWeight dx = PX[r][i] - PX[target][j], dy = PY[r][i] - PY[target][j];
            weight = std::sqrt(dx * dx + dy * dy);
          } else {
            const unsigned randw_seed = (unsigned)reh((GraphElem)(g_i * nv_ + g_j));
            std::default_random_engine rew(randw_seed);
            weight = (GraphWeight)IJW(rew, std::uniform_real_distribution<GraphWeight>::param_type{0.01, 1.0});
          }
        }
        added.insert(key);
        Edge fwd; fwd.tail_ = g_j; fwd.weight_ = weight;
        Edge rev; rev.tail_ = g_i; rev.weight_ = weight;
        extra[r].push_back(Extra{i, fwd});
        extra[target].push_back(Extra{j, rev});
      }
    }
    // merge: per vertex, RGG list + extras, sorted by tail (stable: RGG entries first among equals)
    for (int r = 0; r < nprocs_; r++) {
      if (extra[r].empty()) continue;
      Adj &A = adj[r];
      std::vector<GraphElem> add(n_ + 1, 0);
      for (const Extra &x : extra[r]) add[x.src_local + 1]++;
      for (GraphElem i = 0; i < n_; i++) add[i + 1] += add[i];
      std::vector<GraphElem> nrow(n_ + 1);
      for (GraphElem i = 0; i <= n_; i++) nrow[i] = A.rowptr[i] + add[i];
      std::vector<Edge> ne(nrow[n_]);
      std::vector<GraphElem> pos(n_);
      for (GraphElem i = 0; i < n_; i++) {
        const GraphElem d = A.rowptr[i + 1] - A.rowptr[i];
        std::copy(A.edges.begin() + A.rowptr[i], A.edges.begin() + A.rowptr[i + 1], ne.begin() + nrow[i]);
        pos[i] = nrow[i] + d;
      }
      for (const Extra &x : extra[r]) ne[pos[x.src_local]++] = x.e;
#pragma omp parallel for schedule(dynamic, 4096)
      for (GraphElem i = 0; i < n_; i++)
        std::stable_sort(ne.begin() + nrow[i], ne.begin() + nrow[i + 1],
                         [](const Edge &a, const Edge &b) { return a.tail_ < b.tail_; });
      A.rowptr.swap(nrow);
      A.edges.swap(ne);
    }
  }

  struct EdgeTailLess {
    bool operator()(const Edge &a, GraphElem t) const { return a.tail_ < t; }
    bool operator()(GraphElem t, const Edge &a) const { return t < a.tail_; }
  };

  GraphElem nv_, n_;
  GraphWeight rn_;
  int nprocs_;
};

}  // namespace mvhost
