// Host half of the compact upload (mvgpu_upload_shard, option compact_upload=1): narrows a run of the reference's
// 16-byte Edge records {int64 tail; double weight} (graph.hpp:60-66) to 4-byte tails, checking on the way that
// every weight is exactly 1.0 and every tail lies in [0, nv), and counting the tails outside [base, bound).
// Plain C++ (no CUDA): compiled by g++ and linked into libmvgpu.so.  An AVX2 body is picked at run time when the
// CPU has it (the generic x86-64 baseline only vectorises this loop poorly: ~4 GB/s per core).  An AVX-512 body with
// non-temporal stores was measured on the GPU box in round 2 and lost (upload 31 -> 44 ms at 16 threads): removed.
#include <stdint.h>
#include <stdlib.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

struct Rec { long long tail; double weight; };

void narrow_scalar(const Rec *src, long long n, long long nv, long long base, long long bound, int32_t *dst,
                   long long &nrem, int &bad) {
  long long r = 0;
  int b = 0;
  for (long long e = 0; e < n; e++) {
    const long long t = src[e].tail;
    b |= (src[e].weight != 1.0) | (t < 0) | (t >= nv);
    dst[e] = (int32_t)t;
    r += (t < base) | (t >= bound);
  }
  nrem += r;
  bad |= b;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void narrow_avx2(const Rec *src, long long n, long long nv, long long base, long long bound,
                                                 int32_t *dst, long long &nrem, int &bad) {
  const __m256i one = _mm256_set1_epi64x(0x3FF0000000000000LL);        // bit pattern of 1.0 (the only one that == 1.0)
  const __m256i zero = _mm256_setzero_si256();
  const __m256i nvm1 = _mm256_set1_epi64x(nv - 1), vbase = _mm256_set1_epi64x(base), vbm1 = _mm256_set1_epi64x(bound - 1);
  __m256i okw = _mm256_set1_epi64x(-1), badr = zero, cnt = zero;
  long long e = 0;
  for (; e + 4 <= n; e += 4) {
    const __m256i r0 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + e));       // t0 w0 t1 w1
    const __m256i r1 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + e + 2));   // t2 w2 t3 w3
    const __m256i t = _mm256_unpacklo_epi64(r0, r1);                                         // t0 t2 | t1 t3
    const __m256i w = _mm256_unpackhi_epi64(r0, r1);                                         // w0 w2 | w1 w3
    okw = _mm256_and_si256(okw, _mm256_cmpeq_epi64(w, one));
    badr = _mm256_or_si256(badr, _mm256_or_si256(_mm256_cmpgt_epi64(zero, t), _mm256_cmpgt_epi64(t, nvm1)));
    const __m256i rem = _mm256_or_si256(_mm256_cmpgt_epi64(vbase, t), _mm256_cmpgt_epi64(t, vbm1));
    cnt = _mm256_sub_epi64(cnt, rem);                                                        // mask is -1 per remote tail
    const __m256i lo = _mm256_shuffle_epi32(t, 0x88);                                        // per lane: (t0 t2 t0 t2) | (t1 t3 t1 t3), 32-bit
    const __m256i q = _mm256_permute4x64_epi64(lo, 0x08);                                    // low 128 bits: t0 t2 t1 t3
    const __m128i x = _mm_shuffle_epi32(_mm256_castsi256_si128(q), 0xD8);                    // t0 t1 t2 t3
    _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + e), x);
  }
  long long c4[4];
  _mm256_storeu_si256(reinterpret_cast<__m256i *>(c4), cnt);
  nrem += c4[0] + c4[1] + c4[2] + c4[3];
  if (_mm256_movemask_epi8(okw) != -1 || !_mm256_testz_si256(badr, badr)) bad |= 1;
  if (e < n) narrow_scalar(src + e, n - e, nv, base, bound, dst + e, nrem, bad);
}

#endif

}  // namespace

// returns through *nremote (added to) and *bad (or-ed into)
extern "C" void mv_narrow_edges(const void *edge_records, long long n, long long nv, long long base, long long bound,
                                int32_t *dst, long long *nremote, int *bad) {
  const Rec *src = static_cast<const Rec *>(edge_records);
  long long r = 0;
  int b = 0;
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) narrow_avx2(src, n, nv, base, bound, dst, r, b);
  else
#endif
    narrow_scalar(src, n, nv, base, bound, dst, r, b);
  *nremote += r;
  *bad |= b;
}
