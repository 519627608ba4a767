// Host-side CSR shard with the reference's `Graph` accessor surface, MPI-free.
//
// Mirrors the *interface* of miniVite's per-process graph (reference graph.hpp:85-296:
// get_lnv/get_lne/get_nv/get_ne, get_base/get_bound/get_owner, edge_range, get_edge,
// set_edge, repart, public edge_indices_/edge_list_) so that host code written against
// the reference's Graph compiles against this one.  "rank" here means GPU shard.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <vector>

using GraphElem = int64_t;     // reference utils.hpp:78 (default 64-bit build)
using GraphWeight = double;    // reference utils.hpp:79

#ifndef MAX_PRINT_NEDGE
#define MAX_PRINT_NEDGE (10000000)
#endif

// 16-byte AoS edge record: identical memory layout to the reference's Edge
// (graph.hpp:60-66) and to the on-disk record of the binary graph format.
struct Edge {
  GraphElem tail_ = -1;
  GraphWeight weight_ = 0.0;
};
static_assert(sizeof(Edge) == 16, "Edge must be the 16-byte on-disk record");

class Graph {
 public:
  // `rank`/`size`: which shard of how many this object holds (the reference takes
  // these from its MPI communicator, graph.hpp:92-93).
  Graph(GraphElem lnv, GraphElem lne, GraphElem nv, GraphElem ne, int rank = 0, int size = 1)
      : edge_indices_(lnv + 1, 0), edge_list_(lne), lnv_(lnv), lne_(lne), nv_(nv), ne_(ne),
        parts_(size + 1), rank_(rank), size_(size) {
    for (int r = 0; r <= size_; r++) parts_[r] = (nv_ * r) / size_;   // graph.hpp:112-113
  }

  void repart(const std::vector<GraphElem> &parts) { parts_.assign(parts.begin(), parts.begin() + size_ + 1); }
  const std::vector<GraphElem> &parts() const { return parts_; }

  void set_edge_index(GraphElem vertex, GraphElem e0) { edge_indices_[vertex] = e0; }
  void edge_range(GraphElem vertex, GraphElem &e0, GraphElem &e1) const {
    e0 = edge_indices_[vertex];
    e1 = edge_indices_[vertex + 1];
  }
  // Local edge count only; the global count is supplied by the caller (no collective here).
  void set_nedges(GraphElem lne, GraphElem ne_global = -1) {
    lne_ = lne;
    edge_list_.resize(lne_);
    if (ne_global >= 0) ne_ = ne_global;
  }

  GraphElem get_base(int rank) const { return parts_[rank]; }
  GraphElem get_bound(int rank) const { return parts_[rank + 1]; }
  GraphElem get_range(int rank) const { return parts_[rank + 1] - parts_[rank] + 1; }
  int get_owner(GraphElem vertex) const {
    return int(std::upper_bound(parts_.begin(), parts_.end(), vertex) - parts_.begin()) - 1;
  }

  GraphElem get_lnv() const { return lnv_; }
  GraphElem get_lne() const { return lne_; }
  GraphElem get_nv() const { return nv_; }
  GraphElem get_ne() const { return ne_; }
  int get_rank() const { return rank_; }
  int get_size() const { return size_; }

  const Edge &get_edge(GraphElem index) const { return edge_list_[index]; }
  Edge &set_edge(GraphElem index) { return edge_list_[index]; }

  GraphElem local_to_global(GraphElem idx) const { return idx + parts_[rank_]; }
  GraphElem global_to_local(GraphElem idx) const { return idx - parts_[rank_]; }
  GraphElem local_to_global(GraphElem idx, int rank) const { return idx + parts_[rank]; }
  GraphElem global_to_local(GraphElem idx, int rank) const { return idx - parts_[rank]; }

  // "src dst [weight]" lines, as the reference's -s option prints them (graph.hpp:206-248).
  void print(bool print_weight = true) const {
    if (lne_ >= MAX_PRINT_NEDGE) {
      if (rank_ == 0)
        std::cout << "Graph size per process is {" << lnv_ << ", " << lne_ << "}, which will overwhelm STDOUT."
                  << std::endl;
      return;
    }
    std::cout << "###############" << std::endl;
    std::cout << "Process #" << rank_ << ": " << std::endl;
    std::cout << "###############" << std::endl;
    const GraphElem base = parts_[rank_];
    for (GraphElem v = 0; v < lnv_; v++)
      for (GraphElem e = edge_indices_[v]; e < edge_indices_[v + 1]; e++) {
        std::cout << v + base << " " << edge_list_[e].tail_;
        if (print_weight) std::cout << " " << edge_list_[e].weight_;
        std::cout << std::endl;
      }
  }

  std::vector<GraphElem> edge_indices_;
  std::vector<Edge> edge_list_;

 private:
  GraphElem lnv_, lne_, nv_, ne_;
  std::vector<GraphElem> parts_;
  int rank_, size_;
};

// The "Graph edge distribution characteristics" block of the reference
// (graph.hpp:251-286), computed over all shards held by this process.
inline void print_dist_stats(const std::vector<Graph *> &shards) {
  if (shards.empty()) return;
  long sumdeg = 0, maxdeg = 0;
  double sum_sq = 0;
  for (const Graph *g : shards) {
    long lne = (long)g->get_lne();
    sumdeg += lne;
    maxdeg = std::max(maxdeg, lne);
    sum_sq += (double)(lne * lne);
  }
  const double size = (double)shards.size();
  const double average = (double)sumdeg / size, avg_sq = sum_sq / size;
  const double var = avg_sq - average * average;
  std::cout << std::endl;
  std::cout << "-------------------------------------------------------" << std::endl;
  std::cout << "Graph edge distribution characteristics" << std::endl;
  std::cout << "-------------------------------------------------------" << std::endl;
  std::cout << "Number of vertices: " << shards[0]->get_nv() << std::endl;
  std::cout << "Number of edges: " << shards[0]->get_ne() << std::endl;
  std::cout << "Maximum number of edges: " << maxdeg << std::endl;
  std::cout << "Average number of edges: " << average << std::endl;
  std::cout << "Expected value of X^2: " << avg_sq << std::endl;
  std::cout << "Variance: " << var << std::endl;
  std::cout << "Standard deviation: " << std::sqrt(var) << std::endl;
  std::cout << "-------------------------------------------------------" << std::endl;
}
