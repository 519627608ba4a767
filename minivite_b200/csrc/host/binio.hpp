// Binary graph file I/O in the reference's on-disk format
//   { int64 nv; int64 ne; int64 rowptr[nv+1]; { int64 tail; double weight }[ne] }
// (reference graph.hpp:342-403).  Plain POSIX stdio instead of MPI-IO; each shard
// reads only its own slice, like rank r of the reference does.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "graph.hpp"

namespace mvhost {

class BinaryEdgeList {
 public:
  // Vertex-balanced split: shard r owns [nv*r/p, nv*(r+1)/p)  (graph.hpp:344,355).
  Graph *read(int me, int nprocs, int /*ranks_per_node*/, const std::string &file) {
    FILE *fp = open_header(file);
    std::vector<GraphElem> parts(nprocs + 1);
    for (int r = 0; r <= nprocs; r++) parts[r] = (M_ * r) / nprocs;
    Graph *g = read_slice(fp, me, nprocs, parts);
    fclose(fp);
    return g;
  }

  // Edge-balanced split (the reference's -b): greedy fill of ne/p edges per shard,
  // excess piling up on the last one (graph.hpp:416-461).
  void find_balanced_num_edges(int nprocs, const std::string &file, std::vector<GraphElem> &mbins) {
    FILE *fp = open_header(file);
    mbins.assign(nprocs + 1, 0);
    std::vector<GraphElem> nbins(nprocs, 0);
    const GraphElem nbcap = N_ / nprocs;
    GraphElem prev = 0, cur = 0;
    int p = 0;
    std::vector<GraphElem> buf(1 << 16);
    // the reference consumes rowptr[0..nv-1] as "ecount_idx" (graph.hpp:437-452)
    for (GraphElem m = 0; m < M_;) {
      const size_t want = (size_t)std::min<GraphElem>((GraphElem)buf.size(), M_ - m);
      if (fread(buf.data(), sizeof(GraphElem), want, fp) != want) { fclose(fp); throw std::runtime_error("short read"); }
      for (size_t k = 0; k < want; k++, m++) {
        cur = buf[k];
        if (nbins[p] < nbcap || p == nprocs - 1) nbins[p] += cur - prev;
        if (nbins[p] >= nbcap && p < nprocs - 1) p++;
        mbins[p + 1]++;
        prev = cur;
      }
    }
    fclose(fp);
    for (int k = 1; k <= nprocs; k++) mbins[k] += mbins[k - 1];
  }

  Graph *read_balanced(int me, int nprocs, int /*ranks_per_node*/, const std::string &file) {
    std::vector<GraphElem> mbins;
    find_balanced_num_edges(nprocs, file, mbins);
    FILE *fp = open_header(file);
    Graph *g = read_slice(fp, me, nprocs, mbins);
    fclose(fp);
    return g;
  }

  GraphElem nv() const { return M_; }
  GraphElem ne() const { return N_; }

  // Concatenate shards (in rank order) into one file of the same format.
  static void write(const std::string &file, const std::vector<Graph *> &shards) {
    FILE *fp = fopen(file.c_str(), "wb");
    if (!fp) throw std::runtime_error("cannot open " + file + " for writing");
    GraphElem nv = shards[0]->get_nv(), ne = 0;
    for (const Graph *g : shards) ne += g->get_lne();
    fwrite(&nv, sizeof nv, 1, fp);
    fwrite(&ne, sizeof ne, 1, fp);
    GraphElem off = 0;
    std::vector<GraphElem> tmp;
    for (size_t s = 0; s < shards.size(); s++) {
      const Graph *g = shards[s];
      const GraphElem lnv = g->get_lnv();
      tmp.resize(lnv + (s + 1 == shards.size() ? 1 : 0));
      for (size_t i = 0; i < tmp.size(); i++) tmp[i] = g->edge_indices_[i] + off;
      fwrite(tmp.data(), sizeof(GraphElem), tmp.size(), fp);
      off += g->get_lne();
    }
    for (const Graph *g : shards) fwrite(g->edge_list_.data(), sizeof(Edge), (size_t)g->get_lne(), fp);
    if (fclose(fp) != 0) throw std::runtime_error("write failed: " + file);
  }

 private:
  FILE *open_header(const std::string &file) {
    FILE *fp = fopen(file.c_str(), "rb");
    if (!fp) throw std::runtime_error(" Error opening file! ");
    if (fread(&M_, sizeof(GraphElem), 1, fp) != 1 || fread(&N_, sizeof(GraphElem), 1, fp) != 1) {
      fclose(fp);
      throw std::runtime_error("short header in " + file);
    }
    return fp;
  }

  Graph *read_slice(FILE *fp, int me, int nprocs, const std::vector<GraphElem> &parts) {
    const GraphElem v0 = parts[me], lnv = parts[me + 1] - parts[me];
    Graph *g = new Graph(lnv, 0, M_, N_, me, nprocs);
    g->repart(parts);
    fseeko(fp, (off_t)(2 * sizeof(GraphElem) + v0 * sizeof(GraphElem)), SEEK_SET);
    if (fread(g->edge_indices_.data(), sizeof(GraphElem), (size_t)lnv + 1, fp) != (size_t)lnv + 1)
      throw std::runtime_error("short read (rowptr)");
    const GraphElem e0 = g->edge_indices_[0], lne = g->edge_indices_[lnv] - e0;
    g->set_nedges(lne, N_);
    fseeko(fp, (off_t)(2 * sizeof(GraphElem) + (M_ + 1) * sizeof(GraphElem) + e0 * sizeof(Edge)), SEEK_SET);
    if (lne && fread(g->edge_list_.data(), sizeof(Edge), (size_t)lne, fp) != (size_t)lne)
      throw std::runtime_error("short read (edges)");
    for (GraphElem i = 0; i <= lnv; i++) g->edge_indices_[i] -= e0;   // graph.hpp:407-409
    return g;
  }

  GraphElem M_ = -1, N_ = -1;
};

}  // namespace mvhost
