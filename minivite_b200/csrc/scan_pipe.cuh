// Neighbour-scan kernel, third generation: persistent warps fed by the TMA engine.
//
// Same arithmetic as k_scan_ws (kernels.cuh) -- distExecuteLouvainIteration + distBuildLocalMapCounter +
// distGetMaxIndex, dspl.hpp:276-405, 230-274, 174-228 -- organised around what round 1's profile showed
// (profiles/r1_scan_ws_final_it12_summary.md): k_scan_ws was instruction-issue bound with a third of the issue
// slots empty, because the four warps of a CTA staged, met at a block barrier, reduced and met again, and every tile
// started with a dependent chain "load row offsets -> load tails -> gather cur[tail]" of three memory latencies.  Here
//   * the unit of work is a GROUP of 32 consecutive vertices owned by ONE warp; warps never meet at a block barrier
//     (only __syncwarp), so a warp that waits for its gathers leaves the issue slots to the others;
//   * the kernel is persistent: gridDim = SMs x resident CTAs, every warp strides over its groups;
//   * both streamed inputs arrive in shared memory through the TMA engine (cp.async.bulk + mbarrier complete_tx),
//     issued by one lane and never waited for in the common case: the row offsets of group k+2 and the tails of
//     group k+1 are in flight while the warp reduces group k (a ring of 3 row buffers and 2 tail buffers per warp);
//   * phase A reads four tails per 16-byte shared-memory load, gathers, and writes the four communities back in
//     place (one 4-byte slot per edge);
//   * iteration 1 of a simple graph has its own reduction (FIRST): every vertex is a singleton, so each neighbour is
//     a distinct community with one edge, and the best move is the neighbour of smallest degree.
// Groups whose edges do not fit one buffer are processed in sub-ranges with synchronous bulk copies; vertices with
// more than heavy_deg (< CAP) edges are left to k_scan_heavy.
#pragma once
#include "kernels.cuh"

namespace mv {

#ifndef MV_WCAP_UNIT
#define MV_WCAP_UNIT 512
#endif
#ifndef MV_WCAP_W
#define MV_WCAP_W 320
#endif
#ifndef MV_PW_WARPS
#define MV_PW_WARPS 8
#endif
#ifndef MV_PW_RES_WARPS
#define MV_PW_RES_WARPS 32             // resident warps per SM the register budget is sized for (unit path): 64 registers
#endif
constexpr int kPwWarps = MV_PW_WARPS;                 // warps per CTA (no block-level cooperation between them)
constexpr int kPwRows = 40;                           // row-offset slots per ring entry: 33 used, 36 copied (16-byte multiple)
template <bool UNIT> struct PwCap { static constexpr int value = UNIT ? MV_WCAP_UNIT : MV_WCAP_W; };   // edges per warp buffer

// bytes of one warp's carve-out, a multiple of 128 so that every warp's buffers keep the 16-byte alignment the bulk copies
// and the 16-byte shared-memory accesses need
template <bool UNIT> struct PwWarpBytes {
  static constexpr int value = ((2 * PwCap<UNIT>::value * (UNIT ? 4 : 12) + 3 * kPwRows * 4 + 5 * 8) + 127) / 128 * 128;
};
template <bool UNIT>
constexpr size_t pw_smem_bytes() { return (size_t)kPwWarps * PwWarpBytes<UNIT>::value; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// a bulk copy completes within microseconds; a wait that lasts seconds is a bug -> trap instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
// 1-D bulk copy global -> shared through the TMA engine; completion is counted in bytes on the mbarrier.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool UNIT, bool MULTI, bool TRACE, bool FIRST>
__global__ void __launch_bounds__(kPwWarps * 32, UNIT ? MV_PW_RES_WARPS / kPwWarps : 24 / kPwWarps)
k_scan_pw(const ScanParams p, int ngroups) {
  static_assert(!FIRST || UNIT, "the first-iteration reduction exists for the unit-weight path only");
  constexpr int CAP = PwCap<UNIT>::value;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // per-warp carve-out: tails/communities [2][CAP] int32 | (weights [2][CAP] double) | row offsets [3][kPwRows] | 5 mbarriers
  constexpr int kWarpBytes = PwWarpBytes<UNIT>::value;
  static_assert(CAP % 4 == 0 && kPwRows % 4 == 0, "buffers must be 16-byte multiples");
  unsigned char *wbase = smem_raw + (size_t)wid * kWarpBytes;
  int32_t *s_c = reinterpret_cast<int32_t *>(wbase);
  double *s_wt = reinterpret_cast<double *>(wbase + 2 * CAP * 4);
  uint32_t *s_rows = reinterpret_cast<uint32_t *>(wbase + 2 * CAP * (UNIT ? 4 : 12));
  const uint32_t bar0 = smem_u32(wbase + 2 * CAP * (UNIT ? 4 : 12) + 3 * kPwRows * 4);   // [0,1] tails, [2,3,4] rows
  __shared__ unsigned long long s_red[3][kPwWarps];
  __shared__ double s_redd[kPwWarps];

  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < 5; b++) mbar_init(bar0 + 8 * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int gstride = gridDim.x * kPwWarps;
  int g = blockIdx.x * kPwWarps + wid;

  unsigned long long acc_le_u = 0, acc_moved = 0, acc_hash = 0;
  double acc_le_d = 0.0;
  uint32_t phase_bits = 0;                       // bit b = parity the next wait on barrier b expects

  // ---- producers (lane 0) and consumers of the two rings
  auto issue_rows = [&](int gg, int slot) {      // row offsets rowptr[32*gg .. 32*gg + 36) -> ring slot
    if (lane == 0) {
      mbar_expect_tx(bar0 + 8 * (2 + slot), 36 * 4);
      bulk_g2s(smem_u32(s_rows + slot * kPwRows), p.rowptr + (size_t)gg * 32, 36 * 4, bar0 + 8 * (2 + slot));
    }
  };
  auto issue_tails = [&](int b, uint32_t lo4, uint32_t n4) {   // edges [lo4, lo4 + n4), both multiples of 4, n4 <= CAP
    if (lane == 0) {
      mbar_expect_tx(bar0 + 8 * b, n4 * (UNIT ? 4u : 12u));
      bulk_g2s(smem_u32(s_c + b * CAP), p.tails + lo4, n4 * 4u, bar0 + 8 * b);
      if (!UNIT) bulk_g2s(smem_u32(s_wt + b * CAP), p.weights + lo4, n4 * 8u, bar0 + 8 * b);
    }
  };
  auto wait_bar = [&](int b) {
    mbar_wait(bar0 + 8 * b, (phase_bits >> b) & 1u);
    phase_bits ^= (1u << b);
  };
  // extent of a group's edge block from its (arrived) row offsets: [lo4, lo4 + n4), clipped to one buffer
  auto extent = [&](int gg, int slot, uint32_t &lo4, uint32_t &n4) {
    const int nv = min(32, p.lnv - gg * 32);
    const uint32_t e0 = s_rows[slot * kPwRows], e1 = s_rows[slot * kPwRows + nv];
    lo4 = e0 & ~3u;
    n4 = (e1 > lo4) ? min((uint32_t)CAP, (e1 - lo4 + 3u) & ~3u) : 0u;
  };

  // ---- prologue: rows of the first two groups, tails of the first
  int rslot = 0, buf = 0;                        // ring positions of the current group
  uint32_t have_lo = 0, have_n = 0;              // what the current tail buffer holds: edges [have_lo, have_lo + have_n)
  if (g < ngroups) {
    issue_rows(g, 0);
    if (g + gstride < ngroups) issue_rows(g + gstride, 1);
    wait_bar(2);
    extent(g, 0, have_lo, have_n);
    if (have_n) issue_tails(0, have_lo, have_n);
  }

  for (; g < ngroups; g += gstride) {
    // ---- keep the rings full: rows of group k+2, tails of group k+1 (its rows arrived one iteration ago)
    const int rs1 = rslot == 2 ? 0 : rslot + 1, rs2 = rs1 == 2 ? 0 : rs1 + 1;
    if (g + 2 * gstride < ngroups) issue_rows(g + 2 * gstride, rs2);
    uint32_t nx_lo = 0, nx_n = 0;
    if (g + gstride < ngroups) {
      wait_bar(2 + rs1);
      extent(g + gstride, rs1, nx_lo, nx_n);
      if (nx_n) issue_tails(buf ^ 1, nx_lo, nx_n);
    }
    if (have_n) wait_bar(buf);

    const int v = g * 32 + lane;
    const int nvalid = min(32, p.lnv - g * 32);
    const uint32_t *rows = s_rows + rslot * kPwRows;
    uint32_t ra = 0, rb = 0;
    if (lane < nvalid) { ra = rows[lane]; rb = rows[lane + 1]; }
    const uint32_t deg = rb - ra;
    const bool is_heavy = p.has_heavy && deg > (uint32_t)p.heavy_deg;
    int32_t *sc = s_c + buf * CAP;
    double *sw = s_wt + buf * CAP;
    const uint32_t g_e0 = rows[0], g_e1 = rows[nvalid];
    // the common case: no vertex of the shard is heavy and the whole group sits in the prefetched buffer
    const bool whole = !p.has_heavy && (g_e1 - have_lo) <= have_n;

    int start = 0;
    while (start < nvalid) {
      // ---- sub-range [start, end): longest run of non-heavy vertices whose edges fit the buffer
      int end = nvalid;
      uint32_t e_lo = g_e0, e_end = g_e1;
      if (!whole) {
        const uint32_t heavy_mask = __ballot_sync(0xffffffffu, is_heavy);
        if ((heavy_mask >> start) & 1u) { start++; continue; }
        e_lo = __shfl_sync(0xffffffffu, ra, start);
        const uint32_t a_sub = e_lo & ~3u;
        const bool fits = lane >= start && lane < nvalid && !is_heavy && (rb - a_sub) <= (uint32_t)CAP;
        const uint32_t stop_mask = ~__ballot_sync(0xffffffffu, fits) & (0xffffffffu << start);
        end = stop_mask ? (__ffs(stop_mask) - 1) : 32;               // > start: the start vertex always fits
        e_end = __shfl_sync(0xffffffffu, rb, end - 1);
        if (!(a_sub >= have_lo && e_end <= have_lo + have_n)) {
          // not covered by what the buffer holds (oversize group): synchronous bulk copy of this sub-range
          have_lo = a_sub;
          have_n = (e_end - a_sub + 3u) & ~3u;
          if (have_n) { issue_tails(buf, have_lo, have_n); wait_bar(buf); }
        }
      }
      const bool mine = lane >= start && lane < end;
      const int d = mine ? (int)deg : 0;
      const int o0 = (int)(ra - have_lo);
      const int ne = (int)(e_end - e_lo), eoff = (int)(e_lo - have_lo);

      // ---- phase A: tails (already in shared memory) -> communities, in place; 16-byte shared-memory accesses,
      // all gathers of a pass in flight before the first store.  FIRST: the neighbour's degree instead (every
      // neighbour is its own community).
      {
        int4 *sc4 = reinterpret_cast<int4 *>(sc);
        const int q0 = eoff >> 2, q1 = (eoff + ne + 3) >> 2;
        constexpr int U = 3;
        for (int qb = q0; qb < q1; qb += U * 32) {
          int4 t[U];
#pragma unroll
          for (int u = 0; u < U; u++) {
            const int q = qb + u * 32 + lane;
            if (q < q1) t[u] = sc4[q];
          }
#pragma unroll
          for (int u = 0; u < U; u++) {
            const int q = qb + u * 32 + lane;
            if (q < q1) {
              const int i0 = 4 * q - eoff;               // position of the quad's first entry inside the sub-range
              auto gat = [&](int tl) -> int {
                if (!FIRST) return __ldg(p.cur + tl);
                if (!MULTI || tl < p.lnv) return (int)__ldg(p.loc_cdeg + tl);
                return (int)__ldg(at_cdeg<MULTI>(p, __ldg(p.cur + tl)));
              };
              if ((unsigned)(i0 + 0) < (unsigned)ne) t[u].x = gat(t[u].x);
              if ((unsigned)(i0 + 1) < (unsigned)ne) t[u].y = gat(t[u].y);
              if ((unsigned)(i0 + 2) < (unsigned)ne) t[u].z = gat(t[u].z);
              if ((unsigned)(i0 + 3) < (unsigned)ne) t[u].w = gat(t[u].w);
            }
          }
#pragma unroll
          for (int u = 0; u < U; u++) {
            const int q = qb + u * 32 + lane;
            if (q < q1) sc4[q] = t[u];
          }
        }
      }
      int cc = 0, best = 0;
      if (mine) { cc = __ldg(p.cur + v); best = cc; }
      double cc_deg = 0.0, vdeg = 0.0, sl = 0.0;
      if (d && !FIRST) {
        if (UNIT) {
          cc_deg = (double)__ldg(at_cdeg<MULTI>(p, cc));
          vdeg = (double)d;
          sl = p.has_self ? (double)__ldg(p.self_i + v) : 0.0;
        } else {
          cc_deg = __ldg(&at_cinfo_w<MULTI>(p, cc)->degree);
          vdeg = __ldg(p.vdeg + v);
          sl = p.has_self ? __ldg(p.self_d + v) : 0.0;
        }
      }
      __syncwarp();

      int lbest = kNoLabel;
      if (FIRST) {
        // ---- iteration 1 on a simple graph without self loops: cur[] is the identity, every community a singleton.
        // counter[0] = 0, eix = 0, ax = 0, and neighbour y offers gain_of(1, 0, d, deg(y), 0, c): the same expression
        // the general path evaluates, non-increasing in deg(y).  So the winner has the smallest degree; among equal
        // gains the smallest label wins (dspl.hpp:214-215).  sc[] holds the neighbours' degrees.
        int amin = 0x7fffffff;
        for (int k = 0; k < d; k++) amin = min(amin, sc[o0 + k]);
        if (d) {
          const double dd = (double)d;
          const double gmax = gain_of(1.0, 0.0, dd, (double)amin, 0.0, p.constant, p.f32);
          const double gnext = gain_of(1.0, 0.0, dd, (double)amin + 1.0, 0.0, p.constant, p.f32);
          if (gmax > 0.0) {                                   // maxGain starts at 0: only a positive gain moves (dspl.hpp:214)
            const bool collide = !(gnext < gmax);             // never seen: two degrees rounding to one gain -> compare gains
            int by = cc, lb = 0x7fffffff;
            for (int k = 0; k < d; k++) {
              const int a = sc[o0 + k];
              if (a == amin || (collide && gain_of(1.0, 0.0, dd, (double)a, 0.0, p.constant, p.f32) == gmax)) {
                const int tl = __ldg(p.tails + ra + k);
                const int y = (!MULTI || tl < p.lnv) ? (int)p.base + tl : __ldg(p.cur + tl);
                const int ly = label_of<MULTI>(p, y);
                if (ly < lb) { lb = ly; by = y; }
              }
            }
            best = by; lbest = lb;
            // singleton veto (dspl.hpp:224-225): both communities have size 1 in this iteration
            if (lbest > label_of<MULTI>(p, cc)) best = cc;
          }
        }
      } else {
      // ---- phase B, pass 0: weight towards the own community (counter[0], dspl.hpp:312-318), in edge order; the
      // other neighbours are compacted to the front of the segment (stable)
      double w0 = 0.0;
      int cnt0 = 0, m = 0;
      for (int k = 0; k < d; k++) {
        const int x = sc[o0 + k];
        if (x == cc) { if (UNIT) cnt0++; else w0 += sw[o0 + k]; }
        else {
          sc[o0 + m] = x;
          if (!UNIT) sw[o0 + m] = sw[o0 + k];
          m++;
        }
      }
      if (UNIT) w0 = (double)cnt0;
      const double eix = __dsub_rn(w0, sl), ax = __dsub_rn(cc_deg, vdeg);
      if (d) { if (UNIT) acc_le_u += (unsigned long long)cnt0; else acc_le_d += w0; }
      // ---- pass 1: two distinct neighbour communities per lane per round (see k_scan_ws)
      double bg = 0.0;
      for (;;) {
        const bool has = m > 0;
        if (!__any_sync(0xffffffffu, has)) break;
        int ck1 = 0;
        double ay1 = 0.0;
        if (has) {
          ck1 = sc[o0];
          if (UNIT) ay1 = (double)__ldg(at_cdeg<MULTI>(p, ck1));
          else ay1 = __ldg(&at_cinfo_w<MULTI>(p, ck1)->degree);
        }
        int ck2 = -1, c1 = 0, c2 = 0, m2 = 0;
        double sum1 = 0.0, sum2 = 0.0;
        for (int j = 0; j < m; j++) {
          const int x = sc[o0 + j];
          if (x == ck1) { if (UNIT) c1++; else sum1 += sw[o0 + j]; }
          else {
            if (ck2 < 0) ck2 = x;
            if (x == ck2) { if (UNIT) c2++; else sum2 += sw[o0 + j]; }
            else {
              sc[o0 + m2] = x;
              if (!UNIT) sw[o0 + m2] = sw[o0 + j];
              m2++;
            }
          }
        }
        m = m2;
        if (has) {
          if (UNIT) sum1 = (double)c1;
          const double g1 = gain_of(sum1, eix, vdeg, ay1, ax, p.constant, p.f32);
          if (better_l<MULTI>(p, g1, ck1, bg, best, lbest)) { bg = g1; best = ck1; }
          if (ck2 >= 0) {
            double ay2;
            if (UNIT) { ay2 = (double)__ldg(at_cdeg<MULTI>(p, ck2)); sum2 = (double)c2; }
            else ay2 = __ldg(&at_cinfo_w<MULTI>(p, ck2)->degree);
            const double g2 = gain_of(sum2, eix, vdeg, ay2, ax, p.constant, p.f32);
            if (better_l<MULTI>(p, g2, ck2, bg, best, lbest)) { bg = g2; best = ck2; }
          }
        }
      }
      if (mine && d && label_greater<MULTI>(p, best, lbest, cc)) {             // singleton veto, dspl.hpp:224-225
        long long sz_cc, sz_b;
        if (UNIT) {
          sz_cc = __ldg(at_csize<MULTI>(p, cc));
          sz_b = __ldg(at_csize<MULTI>(p, best));
        } else {
          sz_cc = __ldg(&at_cinfo_w<MULTI>(p, cc)->size);
          sz_b = __ldg(&at_cinfo_w<MULTI>(p, best)->size);
        }
        if (sz_cc == 1 && sz_b == 1) best = cc;
      }
      }
      if (mine) {
        if (best != cc) {                                                      // dspl.hpp:331-399
          if (UNIT) push_move_unit<MULTI>(p, cc, best, d);
          else push_move_w<MULTI>(p, cc, best, vdeg);
        }
        p.tgt[v] = best;                                                       // dspl.hpp:404
        if (TRACE) { acc_moved += (best != cc); acc_hash += vhash(label_of<MULTI>(p, (int)(p.base + v)), label_of<MULTI>(p, best)); }
      }
      start = end;
      fence_proxy_async_smem();                  // this lane's generic-proxy accesses to the buffer are ordered before
      __syncwarp();                              // the TMA write that reuses it (issued by lane 0 after the barrier)
    }
    // rotate the rings
    rslot = rs1;
    buf ^= 1;
    have_lo = nx_lo; have_n = nx_n;
  }

  if (UNIT) { const unsigned long long s = warp_sum(acc_le_u); if (lane == 0) s_red[0][wid] = s; }
  else { const double s = warp_sum(acc_le_d); if (lane == 0) s_redd[wid] = s; }
  if (TRACE) {
    const unsigned long long a = warp_sum(acc_moved), b = warp_sum(acc_hash);
    if (lane == 0) { s_red[1][wid] = a; s_red[2][wid] = b; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (UNIT) {
      unsigned long long s = 0;
      for (int w = 0; w < kPwWarps; w++) s += s_red[0][w];
      if (s) atomicAdd(&p.acc->le_u, s);
    } else {
      double s = 0;
      for (int w = 0; w < kPwWarps; w++) s += s_redd[w];
      if (s != 0.0) atomicAdd(&p.acc->le_d, s);
    }
    if (TRACE) {
      unsigned long long a = 0, b = 0;
      for (int w = 0; w < kPwWarps; w++) { a += s_red[1][w]; b += s_red[2][w]; }
      atomicAdd(&p.acc->moved, a);
      atomicAdd(&p.acc->hash, b);
    }
  }
}

}  // namespace mv
