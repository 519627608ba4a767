// oracle/mpi_shim/mpi.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A tiny stand-in for <mpi.h> so that the UNMODIFIED miniVite reference (main.cpp,
// dspl.hpp, graph.hpp, utils.hpp under /root/reference) can be compiled and run
// in a container that has no MPI installation.  Ranks are *processes*: rank 0
// fork()s MVSHIM_NP-1 children inside MPI_Init; point-to-point messages are
// eager copies into a MAP_SHARED arena, one singly linked mailbox per
// (source, destination) pair; collectives are layered on point-to-point.
//
// Only the subset of MPI that the default (non-RMA) build of the reference
// touches is provided (see SURVEY.md section 8(c) for the symbol list).
//
// Environment:
//   MVSHIM_NP       number of ranks (default 1, max 64)
//   MVSHIM_ARENA_GB size of the (lazily committed) shared arena in GiB (default 24)
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef int MPI_Info;
typedef FILE *MPI_File;
struct MPI_Status { int unused; };

#define MPI_COMM_WORLD 0
enum {
  MPI_BYTE = 1, MPI_INT = 2, MPI_FLOAT = 3, MPI_INT32_T = 4,
  MPI_INT64_T = 5, MPI_DOUBLE = 6, MPI_LONG = 7, MVSHIM_STRUCT24 = 100
};
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_SUCCESS 0
#define MPI_PROC_NULL (-2)
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_THREAD_MULTIPLE 3
#define MPI_MODE_RDONLY 0
#define MPI_INFO_NULL 0

namespace mvshim {

static size_t g_struct_bytes = 24;   // extent of the one struct type the reference creates (set by MPI_Type_create_struct)
static inline size_t type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_BYTE: return 1;
    case MPI_INT: case MPI_FLOAT: case MPI_INT32_T: return 4;
    case MVSHIM_STRUCT24: return g_struct_bytes;
    default: return 8;
  }
}

struct Node {                       // one message; payload follows the header
  std::atomic<Node *> next;
  int tag;
  int consumed;
  size_t bytes;
};
struct Header { std::atomic<size_t> bump; size_t cap; };
struct PendingRecv { void *buf; size_t bytes; int src; int tag; bool active; };

static const int kMaxRanks = 64;
static Header *g_hdr;
static char *g_arena;
static int g_np = 1, g_rank = 0;
static pid_t g_children[kMaxRanks];
static Node *g_head[kMaxRanks][kMaxRanks];   // [src][dst], consumer cursor
static Node *g_tail[kMaxRanks][kMaxRanks];   // [src][dst], producer cursor
static std::vector<PendingRecv> g_pending;

static inline Node *alloc_node(size_t bytes) {
  size_t need = (sizeof(Node) + bytes + 63) & ~size_t(63);
  size_t off = g_hdr->bump.fetch_add(need);
  if (off + need > g_hdr->cap) {
    fprintf(stderr, "mvshim: shared arena exhausted (raise MVSHIM_ARENA_GB)\n");
    abort();
  }
  Node *n = (Node *)(g_arena + off);
  new (&n->next) std::atomic<Node *>(nullptr);
  n->consumed = 0;
  n->bytes = bytes;
  return n;
}

static inline void send_bytes(const void *buf, size_t bytes, int dst, int tag) {
  Node *n = alloc_node(bytes);
  n->tag = tag;
  if (bytes) memcpy((char *)(n + 1), buf, bytes);
  Node *t = g_tail[g_rank][dst];
  g_tail[g_rank][dst] = n;
  t->next.store(n, std::memory_order_release);
}

static inline void recv_bytes(void *buf, size_t bytes, int src, int tag) {
  for (;;) {
    Node *h = g_head[src][g_rank];
    for (Node *n = h->next.load(std::memory_order_acquire); n;
         n = n->next.load(std::memory_order_acquire)) {
      if (!n->consumed && n->tag == tag) {
        if (n->bytes != bytes) {
          fprintf(stderr, "mvshim: size mismatch rank %d <- %d tag %d: sent %zu, expected %zu\n",
                  g_rank, src, tag, n->bytes, bytes);
          abort();
        }
        if (bytes) memcpy(buf, (char *)(n + 1), bytes);
        n->consumed = 1;
        Node *hh = g_head[src][g_rank];          // advance past the consumed prefix
        for (;;) {
          Node *nx = hh->next.load(std::memory_order_acquire);
          if (nx && nx->consumed) hh = nx; else break;
        }
        g_head[src][g_rank] = hh;
        return;
      }
    }
    sched_yield();
  }
}

static inline void combine(void *acc, const void *in, int n, MPI_Datatype t, MPI_Op op) {
  for (int i = 0; i < n; i++) {
    switch (t) {
      case MPI_DOUBLE: {
        double *a = (double *)acc; const double *b = (const double *)in;
        a[i] = (op == MPI_SUM) ? a[i] + b[i] : (a[i] > b[i] ? a[i] : b[i]);
      } break;
      case MPI_FLOAT: {
        float *a = (float *)acc; const float *b = (const float *)in;
        a[i] = (op == MPI_SUM) ? a[i] + b[i] : (a[i] > b[i] ? a[i] : b[i]);
      } break;
      case MPI_INT: case MPI_INT32_T: {
        int32_t *a = (int32_t *)acc; const int32_t *b = (const int32_t *)in;
        a[i] = (op == MPI_SUM) ? a[i] + b[i] : (a[i] > b[i] ? a[i] : b[i]);
      } break;
      default: {
        int64_t *a = (int64_t *)acc; const int64_t *b = (const int64_t *)in;
        a[i] = (op == MPI_SUM) ? (int64_t)((uint64_t)a[i] + (uint64_t)b[i]) : (a[i] > b[i] ? a[i] : b[i]);
      } break;
    }
  }
}

}  // namespace mvshim

static inline int MPI_Init(int *, char ***) {
  using namespace mvshim;
  const char *e = getenv("MVSHIM_NP");
  g_np = e ? atoi(e) : 1;
  if (g_np < 1 || g_np > kMaxRanks) { fprintf(stderr, "mvshim: bad MVSHIM_NP\n"); exit(2); }
  const char *g = getenv("MVSHIM_ARENA_GB");
  size_t cap = (size_t)(g ? atol(g) : 24) << 30;
  if (g_np == 1) cap = (size_t)1 << 20;
  g_arena = (char *)mmap(0, cap, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (g_arena == (char *)MAP_FAILED) { perror("mvshim: mmap"); abort(); }
  g_hdr = (Header *)g_arena;
  new (&g_hdr->bump) std::atomic<size_t>(4096);
  g_hdr->cap = cap;
  for (int s = 0; s < g_np; s++)
    for (int d = 0; d < g_np; d++) {
      Node *n = alloc_node(0);
      n->consumed = 1; n->tag = -12345;
      g_head[s][d] = g_tail[s][d] = n;
    }
  fflush(stdout); fflush(stderr);
  for (int r = 1; r < g_np; r++) {
    pid_t p = fork();
    if (p < 0) { perror("mvshim: fork"); abort(); }
    if (p == 0) { g_rank = r; break; }
    g_children[r] = p;
  }
  return 0;
}
static inline int MPI_Init_thread(int *a, char ***b, int req, int *prov) { *prov = req; return MPI_Init(a, b); }
static inline int MPI_Abort(MPI_Comm, int code) {
  using namespace mvshim;
  if (g_rank == 0) { for (int r = 1; r < g_np; r++) if (g_children[r] > 0) kill(g_children[r], SIGKILL); }
  else kill(getppid(), SIGTERM);
  _exit(code & 0xff);
}
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = mvshim::g_np; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = mvshim::g_rank; return 0; }
static inline double MPI_Wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm) {
  using namespace mvshim;
  size_t by = (size_t)n * type_size(t);
  if (g_rank == root) { for (int d = 0; d < g_np; d++) if (d != root) send_bytes(b, by, d, -2); }
  else recv_bytes(b, by, root, -2);
  return 0;
}
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm) {
  using namespace mvshim;
  size_t by = (size_t)n * type_size(t);
  if (g_rank == root) {
    std::vector<char> tmp(by), acc(by);
    bool first = true;
    for (int p = 0; p < g_np; p++) {            // combine in rank order (deterministic)
      const void *src;
      if (p == root) src = s; else { recv_bytes(tmp.data(), by, p, -3); src = tmp.data(); }
      if (first) { memcpy(acc.data(), src, by); first = false; }
      else combine(acc.data(), src, n, t, op);
    }
    memcpy(r, acc.data(), by);
  } else send_bytes(s, by, root, -3);
  return 0;
}
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  std::vector<char> tmp((size_t)n * mvshim::type_size(t));
  MPI_Reduce(s, tmp.data(), n, t, op, 0, c);
  if (mvshim::g_rank == 0) memcpy(r, tmp.data(), tmp.size());
  return MPI_Bcast(r, n, t, 0, c);
}
static inline int MPI_Barrier(MPI_Comm c) {
  int64_t a = 1, b = 0;
  return MPI_Allreduce(&a, &b, 1, MPI_INT64_T, MPI_SUM, c);
}
static inline int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype t, void *r,
                                const int *rc, const int *rd, MPI_Datatype, MPI_Comm) {
  using namespace mvshim;
  size_t z = type_size(t);
  for (int d = 0; d < g_np; d++) {
    if (d == g_rank) memcpy((char *)r + rd[d] * z, (const char *)s + sd[d] * z, sc[d] * z);
    else send_bytes((const char *)s + sd[d] * z, sc[d] * z, d, -4);
  }
  for (int p = 0; p < g_np; p++)
    if (p != g_rank) recv_bytes((char *)r + rd[p] * z, rc[p] * z, p, -4);
  return 0;
}
static inline int MPI_Alltoall(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm c) {
  std::vector<int> cnt(mvshim::g_np, n), dsp(mvshim::g_np);
  for (int i = 0; i < mvshim::g_np; i++) dsp[i] = i * n;
  return MPI_Alltoallv(s, cnt.data(), dsp.data(), t, r, cnt.data(), dsp.data(), t, c);
}
static inline int MPI_Ialltoall(const void *s, int n, MPI_Datatype t, void *r, int m, MPI_Datatype u,
                                MPI_Comm c, MPI_Request *q) {
  *q = MPI_REQUEST_NULL;
  return MPI_Alltoall(s, n, t, r, m, u, c);
}
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm, MPI_Request *q) {
  if (d != MPI_PROC_NULL) mvshim::send_bytes(b, (size_t)n * mvshim::type_size(t), d, tag);
  *q = MPI_REQUEST_NULL;
  return 0;
}
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm, MPI_Request *q) {
  if (s == MPI_PROC_NULL) { *q = MPI_REQUEST_NULL; return 0; }
  mvshim::g_pending.push_back({b, (size_t)n * mvshim::type_size(t), s, tag, true});
  *q = (int)mvshim::g_pending.size() - 1;
  return 0;
}
static inline int MPI_Wait(MPI_Request *q, MPI_Status *) {
  if (*q >= 0) {
    mvshim::PendingRecv &R = mvshim::g_pending[*q];
    if (R.active) { mvshim::recv_bytes(R.buf, R.bytes, R.src, R.tag); R.active = false; }
    *q = MPI_REQUEST_NULL;
  }
  return 0;
}
static inline int MPI_Waitall(int n, MPI_Request *q, MPI_Status *) {
  for (int i = 0; i < n; i++) MPI_Wait(&q[i], 0);
  return 0;
}
static inline int MPI_Sendrecv(const void *sb, int sn, MPI_Datatype st, int d, int stag, void *rb, int rn,
                               MPI_Datatype rt, int s, int rtag, MPI_Comm, MPI_Status *) {
  if (d != MPI_PROC_NULL) mvshim::send_bytes(sb, (size_t)sn * mvshim::type_size(st), d, stag);
  if (s != MPI_PROC_NULL) mvshim::recv_bytes(rb, (size_t)rn * mvshim::type_size(rt), s, rtag);
  return 0;
}
static inline int MPI_Finalize() {
  MPI_Barrier(0);
  fflush(stdout); fflush(stderr);
  if (mvshim::g_rank != 0) _exit(0);
  int st;
  while (wait(&st) > 0) {}
  return 0;
}
static inline int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)p; return 0; }
static inline int MPI_Type_create_struct(int n, const int *blens, const MPI_Aint *displ, const MPI_Datatype *types, MPI_Datatype *t) {
  // the only struct type the reference creates is CommInfo {GraphElem, GraphElem, GraphWeight}: 24 bytes in the default
  // build, 12 with -DUSE_32_BIT_GRAPH.  Extent = end of the last member, rounded up to the widest member.
  size_t end = 0, align = 1;
  for (int i = 0; i < n; i++) {
    const size_t sz = mvshim::type_size(types[i]);
    end = std::max(end, (size_t)displ[i] + sz * (size_t)blens[i]);
    align = std::max(align, sz);
  }
  mvshim::g_struct_bytes = (end + align - 1) / align * align;
  *t = MVSHIM_STRUCT24;
  return 0;
}
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_Type_free(MPI_Datatype *) { return 0; }
static inline int MPI_Info_create(MPI_Info *) { return 0; }
static inline int MPI_Info_set(MPI_Info, const char *, const char *) { return 0; }
static inline int MPI_Info_free(MPI_Info *) { return 0; }
static inline int MPI_File_open(MPI_Comm, const char *f, int, MPI_Info, MPI_File *fh) {
  *fh = fopen(f, "rb");
  return *fh ? MPI_SUCCESS : 1;
}
static inline int MPI_File_read_all(MPI_File fh, void *b, int n, MPI_Datatype t, MPI_Status *) {
  return fread(b, mvshim::type_size(t), n, fh) == (size_t)n ? 0 : 1;
}
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset o, void *b, int n, MPI_Datatype t, MPI_Status *) {
  fseeko(fh, o, SEEK_SET);
  return fread(b, mvshim::type_size(t), n, fh) == (size_t)n ? 0 : 1;
}
static inline int MPI_File_close(MPI_File *fh) { fclose(*fh); return 0; }
