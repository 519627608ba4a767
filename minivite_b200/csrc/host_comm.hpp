// Host-side bootstrap transport between the ranks of ONE box: a POSIX shared-memory segment named after the 128-byte
// communicator id, with a sense-reversing barrier and one outbox per rank.  It carries the setup-time exchanges of
// the Louvain phase (ghost lists, counts, IPC handles: a few KB to a few MB, once per run) when NCCL cannot --
// NCCL refuses two ranks on the same device, and that is exactly how the multi-rank code paths (ghost discovery,
// peer-memory exchange, remote community reads / delta atomics) are exercised on a box with a single GPU.
// The per-iteration data plane never goes through here: it is peer-memory stores and flags (comm_mode 1).
// Works for ranks that are processes and for ranks that are threads of one process alike.
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace mvhost {

struct HostComm {
  struct Header {
    std::atomic<uint32_t> magic;
    std::atomic<int> attached, arrived, sense;
    int nranks;
    int pad;
  };
  static constexpr uint32_t kMagic = 0x4d564743u;          // "MVGC"
  static constexpr size_t kSlot = 1u << 20;                // outbox bytes per rank
  Header *hdr = nullptr;
  unsigned char *base = nullptr;
  size_t total = 0;
  int rank = 0, nranks = 1, local_sense = 0;
  std::string name;

  bool is_open() const { return hdr != nullptr; }
  unsigned char *outbox(int r) const { return base + 4096 + (size_t)r * kSlot; }

  int open(const void *id128, int rank_, int nranks_, std::string &err) {
    rank = rank_; nranks = nranks_;
    char hex[40];
    const unsigned char *b = reinterpret_cast<const unsigned char *>(id128);
    unsigned long long h = 1469598103934665603ULL;
    for (int i = 0; i < 128; i++) { h ^= b[i]; h *= 1099511628211ULL; }
    snprintf(hex, sizeof hex, "/mvgpu_%016llx", h);
    name = hex;
    total = 4096 + (size_t)nranks * kSlot;
    int fd = -1;
    if (rank == 0) {
      shm_unlink(name.c_str());
      fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
      if (fd < 0 || ftruncate(fd, (off_t)total) != 0) { err = "host transport: cannot create " + name; if (fd >= 0) close(fd); return 1; }
    } else {
      const auto t0 = std::chrono::steady_clock::now();
      for (;;) {
        fd = shm_open(name.c_str(), O_RDWR, 0600);
        struct stat st;
        if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= total) break;
        if (fd >= 0) { close(fd); fd = -1; }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) { err = "host transport: rank 0 never created " + name; return 1; }
        usleep(1000);
      }
    }
    void *m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { err = "host transport: mmap failed"; return 1; }
    base = reinterpret_cast<unsigned char *>(m);
    hdr = reinterpret_cast<Header *>(m);
    if (rank == 0) {
      hdr->attached.store(0); hdr->arrived.store(0); hdr->sense.store(0); hdr->nranks = nranks;
      hdr->magic.store(kMagic, std::memory_order_release);
    } else {
      const auto t0 = std::chrono::steady_clock::now();
      while (hdr->magic.load(std::memory_order_acquire) != kMagic) {
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) { err = "host transport: segment never initialised"; return 1; }
        usleep(200);
      }
      if (hdr->nranks != nranks) { err = "host transport: ranks disagree on the communicator size"; return 1; }
    }
    hdr->attached.fetch_add(1);
    while (hdr->attached.load() < nranks) sched_yield();
    barrier();
    if (rank == 0) shm_unlink(name.c_str());               // the mapping lives on; the name is free again
    return 0;
  }
  void close_comm() {
    if (base) munmap(base, total);
    base = nullptr; hdr = nullptr;
  }
  void barrier() {
    local_sense ^= 1;
    if (hdr->arrived.fetch_add(1, std::memory_order_acq_rel) == nranks - 1) {
      hdr->arrived.store(0, std::memory_order_relaxed);
      hdr->sense.store(local_sense, std::memory_order_release);
    } else {
      int spins = 0;
      while (hdr->sense.load(std::memory_order_acquire) != local_sense)
        if (++spins > 2000) sched_yield();
    }
  }
  // every rank contributes `bytes` (<= kSlot); all[r*bytes ..] = rank r's contribution
  void allgather(const void *mine, void *all, size_t bytes) {
    memcpy(outbox(rank), mine, bytes);
    barrier();
    for (int r = 0; r < nranks; r++) memcpy(reinterpret_cast<unsigned char *>(all) + (size_t)r * bytes, outbox(r), bytes);
    barrier();
  }
  template <typename T, typename Op>
  void allreduce(T *vals, int n, Op op) {                  // combined in rank order on every rank: identical results
    std::vector<T> all((size_t)n * nranks);
    allgather(vals, all.data(), sizeof(T) * n);
    for (int k = 0; k < n; k++) {
      T acc = all[k];
      for (int r = 1; r < nranks; r++) acc = op(acc, all[(size_t)r * n + k]);
      vals[k] = acc;
    }
  }
  // all-to-all-v of byte strings: send + soff[r] .. + scount[r] goes to rank r; recv + roff[r] .. + rcount[r] comes from r
  // (counts in bytes, known on both sides).  Pieces larger than a share of the outbox go in several rounds.
  void alltoallv(const unsigned char *send, const size_t *scount, const size_t *soff, unsigned char *recv, const size_t *rcount,
                 const size_t *roff) {
    const size_t piece = (kSlot / (size_t)nranks) & ~(size_t)63;
    size_t mx = 0;
    for (int r = 0; r < nranks; r++) mx = std::max(mx, std::max(scount[r], rcount[r]));
    unsigned long long gmx = mx;
    allreduce(&gmx, 1, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; });
    for (size_t done = 0; done < gmx; done += piece) {
      for (int r = 0; r < nranks; r++)
        if (scount[r] > done) memcpy(outbox(rank) + (size_t)r * piece, send + soff[r] + done, std::min(piece, scount[r] - done));
      barrier();
      for (int r = 0; r < nranks; r++)
        if (rcount[r] > done) memcpy(recv + roff[r] + done, outbox(r) + (size_t)rank * piece, std::min(piece, rcount[r] - done));
      barrier();
    }
  }
};

}  // namespace mvhost
