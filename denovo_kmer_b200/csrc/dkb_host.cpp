// dkb_host.cpp — host side of the hot path's boundary: the pure k-mer
// primitives of src/kmer.rs, the read packer (decoded BAM records -> packed
// stream) and the variant k-mer builder (candidate alleles + flanks ->
// spanning k-mer entries for kernel 1).  The reference files are unmounted;
// the semantics are DESIGN.md §2.  No CUDA in this file.
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/dkb.h"

namespace {

// A=0 C=1 G=2 T=3 (either case); 4 = anything else
struct BaseLut {
  uint8_t t[256];
  BaseLut() {
    memset(t, 4, sizeof(t));
    t['A'] = t['a'] = 0;
    t['C'] = t['c'] = 1;
    t['G'] = t['g'] = 2;
    t['T'] = t['t'] = 3;
  }
};
const BaseLut LUT;
// BAM 4-bit code (=ACMGRSVTWYHKDBN: A=1 C=2 G=4 T=8) -> 2-bit code; 4 = not a plain base
const uint8_t NIB_LUT[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};

inline uint64_t kmask(int k) { return (1ull << (2 * k)) - 1; }

inline uint64_t revcomp(uint64_t fwd, int k) {
  // complement, then reverse the 2-bit groups of the 64-bit word and realign
  uint64_t x = ~fwd;
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = __builtin_bswap64(x);
  return x >> (64 - 2 * k);
}

// fn(0) .. fn(n - 1), one item per thread, the last on the calling thread.  Nothing escapes: an
// item that throws (out of memory) makes the call return false, and items for which no thread
// could be started run on the calling thread.
template <class F>
bool run_parallel(unsigned n, F &&fn) {
  std::atomic<bool> ok{true};
  auto guarded = [&](unsigned t) {
    try {
      fn(t);
    } catch (...) {
      ok = false;
    }
  };
  std::vector<std::thread> pool;
  unsigned started = 0;
  try {
    pool.reserve(n ? n - 1 : 0);
    for (; started + 1 < n; started++) pool.emplace_back(guarded, started);
  } catch (...) {  // no more threads (or no memory for the pool): the rest runs here
  }
  for (unsigned t = started; t < n; t++) guarded(t);
  for (auto &th : pool) th.join();
  return ok;
}

// The ABI promises that nothing throws across it: every entry point that allocates runs its
// body through this.
template <class F>
int no_throw(F &&body) {
  try {
    return body();
  } catch (const std::bad_alloc &) {
    return DKB_ENOMEM;
  } catch (...) {
    return DKB_EINVAL;
  }
}

}  // namespace

extern "C" {

int dkb_kmer_encode(const char *seq, int k, uint64_t *fwd_out) {
  if (!seq || !fwd_out || k < 1 || k > DKB_MAX_K) return DKB_EINVAL;
  uint64_t v = 0;
  for (int i = 0; i < k; i++) {
    const uint8_t c = LUT.t[(uint8_t)seq[i]];
    if (c > 3) return DKB_EINVAL;
    v = (v << 2) | c;
  }
  *fwd_out = v;
  return DKB_OK;
}

uint64_t dkb_kmer_revcomp(uint64_t fwd, int k) { return revcomp(fwd & kmask(k), k); }

uint64_t dkb_kmer_canonical(uint64_t fwd, int k) {
  fwd &= kmask(k);
  const uint64_t rc = revcomp(fwd, k);
  return fwd < rc ? fwd : rc;
}

uint64_t dkb_stream_positions(const uint64_t *offsets, size_t n_reads) {
  if (!offsets || n_reads == 0) return 0;
  return (offsets[n_reads] - offsets[0]) + n_reads;
}

// sizes are rounded up to whole 16-byte vectors so device loads stay aligned
size_t dkb_stream_bases_words(uint64_t n_positions) { return (size_t)((n_positions + 63) / 64 * 4); }
size_t dkb_stream_mask_words(uint64_t n_positions) { return (size_t)((n_positions + 127) / 128 * 4); }

// Pack stream positions [p0, p1) (p0 a multiple of 32, so no output word is shared
// between two calls).  Position of read r's first base: offsets[r] - offsets[0] + r.
// nib (4-bit input only, else nullptr): byte offset of every read's first pair of codes.
static void pack_range(const uint8_t *seq, const uint64_t *nib, const uint8_t *qual, const uint64_t *offsets,
                       size_t n_reads, int min_baseq, uint32_t *bases2, uint32_t *mask1,
                       uint64_t p0, uint64_t p1) {
  const uint64_t o0 = offsets[0];
  // last read whose first position is <= p0
  size_t lo = 0, hi = n_reads;
  while (hi - lo > 1) {
    const size_t mid = (lo + hi) / 2;
    if (offsets[mid] - o0 + mid <= p0) lo = mid; else hi = mid;
  }
  // words are built in registers and stored once (the range owns them exclusively)
  uint64_t p = p0;
  uint32_t cur_b = 0, cur_m = 0;
  auto advance = [&](uint64_t to) {  // move to position `to`, flushing completed words
    while ((p >> 4) != (to >> 4)) {
      bases2[p >> 4] = cur_b;
      cur_b = 0;
      p = ((p >> 4) + 1) << 4;
      if ((p & 31) == 0) {
        mask1[(p >> 5) - 1] = cur_m;
        cur_m = 0;
      }
    }
    p = to;
  };
  for (size_t r = lo; r < n_reads && p < p1; r++) {
    const uint64_t start = offsets[r] - o0 + r;  // stream position of the read's first base
    const uint64_t len = offsets[r + 1] - offsets[r];
    uint64_t i = p > start ? p - start : 0;       // first base of this read inside the range
    if (start + i > p) advance(start + i < p1 ? start + i : p1);
    if (p >= p1) break;
    const uint8_t *sp = nib ? seq + nib[r] : seq + offsets[r];
    const uint8_t *qp = qual ? qual + offsets[r] : nullptr;
    const uint64_t end = start + len < p1 ? len : p1 - start;  // bases of this read in range
    for (; i < end; i++) {
      const uint32_t c = nib ? NIB_LUT[(i & 1) ? (sp[i >> 1] & 15u) : (sp[i >> 1] >> 4)] : LUT.t[sp[i]];
      const uint32_t ok = (c <= 3) & (!qp || (int)qp[i] >= min_baseq);
      const uint32_t sh = (uint32_t)(p & 15);
      cur_b |= (ok ? c : 0u) << (2 * sh);
      cur_m |= ok << (p & 31);
      p++;
      if (sh == 15) {
        bases2[(p >> 4) - 1] = cur_b;
        cur_b = 0;
        if ((p & 31) == 0) {
          mask1[(p >> 5) - 1] = cur_m;
          cur_m = 0;
        }
      }
    }
    if (i == len) advance(start + len + 1 < p1 ? start + len + 1 : p1);  // the separator (flag 0)
  }
  advance(p1);
  // p1 is either a multiple of 128 (all words flushed) or the end of the stream
  if (p1 & 15) bases2[p1 >> 4] = cur_b;
  if (p1 & 31) mask1[p1 >> 5] = cur_m;
}

// ---- the same range, 32 or 64 bases at a time (AVX2 / AVX-512 BW, + BMI2; chosen at run time) --
// A chunk of bases becomes three bit masks - usable = (a|c|g|t) & (qual >= threshold), and the two
// bit planes (c|t), (g|t) of the 2-bit codes, zero where the base is unusable - and the planes
// are interleaved into stream bits with two bit deposits per 32 bases.  The range's positions are
// produced strictly in order through two bit writers that keep the unfinished word in a register
// and store every 64-bit word exactly once: OR-ing chunks into memory at arbitrary bit offsets
// instead made each chunk's load wait for the previous chunk's narrower store (no store
// forwarding) and ran at a third of this speed.
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define DKB_HAVE_SIMD_PACK 1

extern "C++" {  // (templates; this part of the file sits inside extern "C")
namespace {
struct BitWriter {
  uint8_t *out;  // next 64-bit word (little-endian: two consecutive uint32 stream words)
  uint64_t acc = 0;
  unsigned nb = 0;  // bits held in acc, < 64
  // append the low n bits of v (1 <= n <= 64; v is zero above them)
  inline void put(uint64_t v, unsigned n) {
    acc |= v << nb;
    if (nb + n >= 64) {
      memcpy(out, &acc, 8);
      out += 8;
      acc = nb ? v >> (64 - nb) : 0;
      nb = nb + n - 64;
    } else {
      nb += n;
    }
  }
  inline void flush() {  // the stream's last, partial word (the buffers are padded to 16 bytes)
    if (nb) memcpy(out, &acc, 8);
  }
};

// Body of both SIMD packers (a macro, not a template: the bit deposits and the chunk lambda must
// be compiled with their function's target options).  Walks the reads of positions [p0, p1) in
// order; chunk(sp, qp, n, ok, pl, ph) turns n <= CH bases into their masks.
#define DKB_PACK_ORDERED_LOOP(CH)                                                                        \
  const uint64_t o0 = offsets[0];                                                                        \
  size_t lo = 0, hi = n_reads;                                                                           \
  while (hi - lo > 1) { /* last read whose first position is <= p0 */                                    \
    const size_t mid = (lo + hi) / 2;                                                                    \
    if (offsets[mid] - o0 + mid <= p0) lo = mid; else hi = mid;                                          \
  }                                                                                                      \
  /* p0 is a multiple of 128: both writers start on a 64-bit word */                                     \
  BitWriter wb{reinterpret_cast<uint8_t *>(bases2) + p0 / 4}, wm{reinterpret_cast<uint8_t *>(mask1) + p0 / 8}; \
  for (size_t r = lo; r < n_reads; r++) {                                                                \
    const uint64_t start = offsets[r] - o0 + r, len = offsets[r + 1] - offsets[r];                       \
    if (start >= p1) break;                                                                              \
    uint64_t i = p0 > start ? p0 - start : 0;                 /* first base of the read inside the range */ \
    const uint64_t end = start + len < p1 ? len : p1 - start; /* one past its last base inside the range */ \
    const uint8_t *sp = nib ? seq + nib[r] : seq + offsets[r];                                           \
    const uint8_t *qp = qual ? qual + offsets[r] : nullptr;                                              \
    if (nib && (i & 1) && i < end) { /* 4-bit input cut inside a byte: its low code on its own */        \
      const uint32_t c = NIB_LUT[sp[i >> 1] & 15u];                                                      \
      const uint32_t okb = (c <= 3) & (!qp || (int)qp[i] >= min_baseq);                                  \
      wb.put(okb ? c : 0u, 2);                                                                           \
      wm.put(okb, 1);                                                                                    \
      i++;                                                                                               \
    }                                                                                                    \
    while (i < end) {                                                                                    \
      const unsigned n = end - i < (uint64_t)(CH) ? (unsigned)(end - i) : (unsigned)(CH);                \
      uint64_t ok, pl, ph;                                                                               \
      chunk(nib ? sp + (i >> 1) : sp + i, qp ? qp + i : nullptr, n, ok, pl, ph);                         \
      const unsigned n0 = n < 32 ? n : 32;                                                               \
      wb.put(_pdep_u64(pl & 0xFFFFFFFFull, 0x5555555555555555ull) |                                      \
                 _pdep_u64(ph & 0xFFFFFFFFull, 0xAAAAAAAAAAAAAAAAull), 2 * n0);                          \
      if ((CH) > 32 && n > 32)                                                                           \
        wb.put(_pdep_u64(pl >> 32, 0x5555555555555555ull) | _pdep_u64(ph >> 32, 0xAAAAAAAAAAAAAAAAull),  \
               2 * (n - 32));                                                                            \
      wm.put(ok, n);                                                                                     \
      i += n;                                                                                            \
    }                                                                                                    \
    if (start + len >= p0 && start + len < p1) { /* the separator after the read: flag 0, code 0 */      \
      wb.put(0, 2);                                                                                      \
      wm.put(0, 1);                                                                                      \
    }                                                                                                    \
  }                                                                                                      \
  wb.flush();                                                                                            \
  wm.flush();
}  // namespace
}  // extern "C++"

__attribute__((target("avx2,bmi2"))) static void pack_range_avx2(
    const uint8_t *seq, const uint64_t *nib, const uint8_t *qual, const uint64_t *offsets, size_t n_reads,
    int min_baseq, uint32_t *bases2, uint32_t *mask1, uint64_t p0, uint64_t p1) {
  const __m256i lower = _mm256_set1_epi8(0x20), cA = _mm256_set1_epi8('a'), cC = _mm256_set1_epi8('c'),
                cG = _mm256_set1_epi8('g'), cT = _mm256_set1_epi8('t');
  const int mq = min_baseq < 0 ? 0 : min_baseq > 255 ? 255 : min_baseq;
  const __m256i thr = _mm256_set1_epi8((char)mq);
  const bool never = qual && min_baseq > 255;  // no quality byte can reach the threshold
  auto chunk = [&](const uint8_t *sp, const uint8_t *qp, unsigned n, uint64_t &ok, uint64_t &pl, uint64_t &ph)
                   __attribute__((target("avx2,bmi2"))) {
        alignas(32) uint8_t ts[32], tq[32];
        const unsigned sbytes = nib ? (n + 1) / 2 : n;  // input bytes of this chunk
        if (n < 32) {  // a read's last bases: through a zero-padded copy (a zero byte / code is no base)
          memset(ts, 0, 32);
          memcpy(ts, sp, sbytes);
          sp = ts;
          if (qp) {
            memset(tq, 0, 32);
            memcpy(tq, qp, n);
            qp = tq;
          }
        }
        __m256i isA, isC, isG, isT;
        if (nib) {
          // 16 bytes = 32 codes, high nibble first: widen every byte to 16 bits and put its high
          // nibble in the low byte, its low nibble in the high byte - bytes in base order
          const __m256i w = _mm256_cvtepu8_epi16(_mm_loadu_si128(reinterpret_cast<const __m128i *>(sp)));
          const __m256i c4 = _mm256_or_si256(_mm256_and_si256(_mm256_srli_epi16(w, 4), _mm256_set1_epi16(0x000F)),
                                             _mm256_slli_epi16(_mm256_and_si256(w, _mm256_set1_epi16(0x000F)), 8));
          isA = _mm256_cmpeq_epi8(c4, _mm256_set1_epi8(1));
          isC = _mm256_cmpeq_epi8(c4, _mm256_set1_epi8(2));
          isG = _mm256_cmpeq_epi8(c4, _mm256_set1_epi8(4));
          isT = _mm256_cmpeq_epi8(c4, _mm256_set1_epi8(8));
        } else {
          const __m256i s = _mm256_or_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(sp)), lower);
          isA = _mm256_cmpeq_epi8(s, cA);
          isC = _mm256_cmpeq_epi8(s, cC);
          isG = _mm256_cmpeq_epi8(s, cG);
          isT = _mm256_cmpeq_epi8(s, cT);
        }
        __m256i okv = _mm256_or_si256(_mm256_or_si256(isA, isC), _mm256_or_si256(isG, isT));
        if (qp) {
          const __m256i q = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(qp));
          okv = _mm256_and_si256(okv, _mm256_cmpeq_epi8(_mm256_max_epu8(q, thr), q));  // q >= threshold, unsigned
        }
        // (an odd chunk's last byte carries a padding code: nothing beyond base n counts)
        const uint32_t m = never ? 0u : (uint32_t)_mm256_movemask_epi8(okv) & (n == 32 ? ~0u : (1u << n) - 1u);
        ok = m;
        pl = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(isC, isT)) & m;
        ph = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(isG, isT)) & m;
      };
  DKB_PACK_ORDERED_LOOP(32)
}

__attribute__((target("avx512f,avx512bw,avx512vl,bmi2"))) static void pack_range_avx512(
    const uint8_t *seq, const uint64_t *nib, const uint8_t *qual, const uint64_t *offsets, size_t n_reads,
    int min_baseq, uint32_t *bases2, uint32_t *mask1, uint64_t p0, uint64_t p1) {
  const __m512i lower = _mm512_set1_epi8(0x20), cA = _mm512_set1_epi8('a'), cC = _mm512_set1_epi8('c'),
                cG = _mm512_set1_epi8('g'), cT = _mm512_set1_epi8('t');
  const int mq = min_baseq < 0 ? 0 : min_baseq > 255 ? 255 : min_baseq;
  const __m512i thr = _mm512_set1_epi8((char)mq);
  const bool never = qual && min_baseq > 255;
  auto chunk = [&](const uint8_t *sp, const uint8_t *qp, unsigned n, uint64_t &ok, uint64_t &pl, uint64_t &ph)
                   __attribute__((target("avx512f,avx512bw,avx512vl,bmi2"))) {
        // a read's last bases are loaded under a byte mask (masked-out bytes are not touched)
        const __mmask64 km = n == 64 ? ~0ull : (1ull << n) - 1;
        __mmask64 isA, isC, isG, isT;
        if (nib) {
          // 32 bytes = 64 codes, high nibble first (see the AVX2 path); bytes beyond the read's
          // last pair are not touched
          const unsigned sb = (n + 1) / 2;
          const __mmask32 kb = sb == 32 ? ~0u : (1u << sb) - 1u;
          const __m512i w = _mm512_cvtepu8_epi16(_mm256_maskz_loadu_epi8(kb, sp));
          const __m512i c4 = _mm512_or_si512(_mm512_and_si512(_mm512_srli_epi16(w, 4), _mm512_set1_epi16(0x000F)),
                                             _mm512_slli_epi16(_mm512_and_si512(w, _mm512_set1_epi16(0x000F)), 8));
          isA = _mm512_cmpeq_epi8_mask(c4, _mm512_set1_epi8(1));
          isC = _mm512_cmpeq_epi8_mask(c4, _mm512_set1_epi8(2));
          isG = _mm512_cmpeq_epi8_mask(c4, _mm512_set1_epi8(4));
          isT = _mm512_cmpeq_epi8_mask(c4, _mm512_set1_epi8(8));
        } else {
          const __m512i s = _mm512_or_si512(_mm512_maskz_loadu_epi8(km, sp), lower);
          isA = _mm512_cmpeq_epi8_mask(s, cA);
          isC = _mm512_cmpeq_epi8_mask(s, cC);
          isG = _mm512_cmpeq_epi8_mask(s, cG);
          isT = _mm512_cmpeq_epi8_mask(s, cT);
        }
        uint64_t m = (uint64_t)(isA | isC | isG | isT) & (uint64_t)km;  // (an odd chunk's padding code)
        if (qp) m &= (uint64_t)_mm512_cmpge_epu8_mask(_mm512_maskz_loadu_epi8(km, qp), thr);
        if (never) m = 0;
        ok = m;
        pl = (uint64_t)(isC | isT) & m;
        ph = (uint64_t)(isG | isT) & m;
      };
  DKB_PACK_ORDERED_LOOP(64)
}
#endif

int dkb_pack_reads(const uint8_t *seq, const uint8_t *qual, const uint64_t *offsets,
                   size_t n_reads, int min_baseq, uint32_t *bases2, uint32_t *mask1,
                   uint64_t *n_positions_out) {
  return dkb_pack_reads_fmt(seq, 0, qual, offsets, n_reads, min_baseq, bases2, mask1, n_positions_out);
}

int dkb_pack_reads_fmt(const uint8_t *seq, int seq_format, const uint8_t *qual, const uint64_t *offsets,
                       size_t n_reads, int min_baseq, uint32_t *bases2, uint32_t *mask1,
                       uint64_t *n_positions_out) {
  return no_throw([&]() -> int {
  if (!offsets && n_reads) return DKB_EINVAL;
  if (!bases2 || !mask1) return DKB_EINVAL;
  if (n_reads && !seq) return DKB_EINVAL;
  if (seq_format != 0 && seq_format != 1) return DKB_EINVAL;
  const uint64_t n_pos = dkb_stream_positions(offsets, n_reads);
  const size_t bw = dkb_stream_bases_words(n_pos), mw = dkb_stream_mask_words(n_pos);
  // 0 scalar, 1 AVX2 + BMI2 (32 bases per step), 2 AVX-512 BW + BMI2 (64 per step); the best the
  // CPU has, DKB_PACK_SCALAR=1 / DKB_PACK_ISA=0|1|2 lower it (tests, timing)
  int simd = 0;
#ifdef DKB_HAVE_SIMD_PACK
  if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2")) simd = 1;
  if (simd && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
      __builtin_cpu_supports("avx512vl"))
    simd = 2;
  if (getenv("DKB_PACK_SCALAR")) simd = 0;
  if (const char *e = getenv("DKB_PACK_ISA")) simd = atoi(e) < simd ? (atoi(e) < 0 ? 0 : atoi(e)) : simd;
#endif
  (void)simd;
  // output words are split between threads on 128-position boundaries: no sharing
  // every hardware thread up to 32, at least 1 M positions each; DKB_PACK_THREADS=n means exactly n
  // (up to 64), whatever the batch size
  unsigned n_thr = std::thread::hardware_concurrency();
  if (n_thr > 32) n_thr = 32;
  if (n_thr > (n_pos >> 20)) n_thr = (unsigned)(n_pos >> 20);
  if (const char *e = getenv("DKB_PACK_THREADS"))
    if (atoi(e) > 0) n_thr = atoi(e) > 64 ? 64u : (unsigned)atoi(e);
  if (n_thr < 1) n_thr = 1;
  // The offsets are checked before anything is read through them, and - BAM keeps every read's
  // 4-bit codes byte-aligned, (len + 1) / 2 bytes per read, back to back - the byte offset of
  // every read is their prefix sum: both by chunks of reads, on the same threads as the packing
  // (at 150-base reads a serial pass over the offsets cost a third of the packing time).
  std::vector<uint64_t> nib_of;
  if (seq_format == 1) nib_of.resize(n_reads);
  {
    const size_t rper = (n_reads + n_thr - 1) / n_thr;
    std::vector<uint64_t> chunk_bytes(n_thr, 0);
    std::atomic<bool> bad{false};
    if (!run_parallel(n_thr, [&](unsigned t) {
          const size_t a = (size_t)t * rper, b = a + rper < n_reads ? a + rper : n_reads;
          uint64_t bytes = 0;
          bool dec = false;
          for (size_t r = a; r < b; r++) {
            dec |= offsets[r + 1] < offsets[r];
            bytes += (offsets[r + 1] - offsets[r] + 1) / 2;
          }
          if (dec) bad = true;
          chunk_bytes[t] = bytes;
        }))
      return DKB_ENOMEM;
    if (bad) return DKB_EINVAL;
    if (seq_format == 1) {
      uint64_t acc = 0;
      for (unsigned t = 0; t < n_thr; t++) {
        const uint64_t c = chunk_bytes[t];
        chunk_bytes[t] = acc;
        acc += c;
      }
      if (!run_parallel(n_thr, [&](unsigned t) {
            const size_t a = (size_t)t * rper, b = a + rper < n_reads ? a + rper : n_reads;
            uint64_t at = chunk_bytes[t];
            for (size_t r = a; r < b; r++) {
              nib_of[r] = at;
              at += (offsets[r + 1] - offsets[r] + 1) / 2;
            }
          }))
        return DKB_ENOMEM;
    }
  }
  const uint64_t *nib = seq_format == 1 ? nib_of.data() : nullptr;
  const uint64_t per = ((n_pos + n_thr - 1) / n_thr + 127) / 128 * 128;
  auto work = [&](unsigned t) {
    const uint64_t p0 = (uint64_t)t * per, p1 = p0 + per < n_pos ? p0 + per : n_pos;
    // this thread's words.  The packers store every word that lies wholly inside [p0, p1) and the
    // one p1 falls into; what is cleared here - before they run - is the rest: from the 64-bit word
    // p1 falls into to the end of the thread's words (the stream's padding, for the tail thread),
    // or everything when the range holds no position.
    const bool packs = p0 < p1 && n_reads;
    size_t b0 = (size_t)(p0 / 16), m0 = (size_t)(p0 / 32);
    const size_t b1 = t + 1 == n_thr ? bw : (size_t)((p0 + per) / 16);
    const size_t m1 = t + 1 == n_thr ? mw : (size_t)((p0 + per) / 32);
    if (packs) {
      b0 = (size_t)(p1 / 32) * 2;
      m0 = (size_t)(p1 / 64) * 2;
    }
    if (b0 < bw && b0 < b1) memset(bases2 + b0, 0, ((b1 < bw ? b1 : bw) - b0) * 4);
    if (m0 < mw && m0 < m1) memset(mask1 + m0, 0, ((m1 < mw ? m1 : mw) - m0) * 4);
    if (packs) {
#ifdef DKB_HAVE_SIMD_PACK
      if (simd == 2) return pack_range_avx512(seq, nib, qual, offsets, n_reads, min_baseq, bases2, mask1, p0, p1);
      if (simd == 1) return pack_range_avx2(seq, nib, qual, offsets, n_reads, min_baseq, bases2, mask1, p0, p1);
#endif
      pack_range(seq, nib, qual, offsets, n_reads, min_baseq, bases2, mask1, p0, p1);
    }
  };
  if (!run_parallel(n_thr, work)) return DKB_ENOMEM;
  if (n_positions_out) *n_positions_out = n_pos;
  return DKB_OK;
  });
}

// ---- dense flags -> zero list (include/dkb.h) ---------------------------------------------
namespace {
constexpr uint64_t ZL_BLOCK = 2048;
constexpr uint32_t ZL_RAW_BIT = 0x80000000u;


// Gap coding of block [p0, p1) (p0 a multiple of 2048) into out (nullptr: only count); returns
// its length in bytes.  Positions >= p1 - the end of the stream - are not coded.  64 flags at a
// time (the stream's mask words come in groups of four, so the last 64-bit load is in bounds);
// the zeros of a word are found with ctz.  out must have room for 2048 + 16 bytes.
size_t zl_block_bytes(const uint32_t *mask1, uint64_t p0, uint64_t p1, uint8_t *out) {
  size_t n = 0;
  uint64_t gap = 0;
  for (uint64_t wp = p0; wp < p1; wp += 64) {
    const unsigned nvalid = p1 - wp >= 64 ? 64u : (unsigned)(p1 - wp);  // positions of this word in range
    uint64_t w;
    memcpy(&w, mask1 + (wp >> 5), 8);
    uint64_t zeros = ~w & (nvalid == 64 ? ~0ull : (1ull << nvalid) - 1);
    unsigned done = 0;  // positions of the word already accounted for
    while (zeros) {
      const unsigned t = (unsigned)__builtin_ctzll(zeros);
      zeros &= zeros - 1;
      gap += t - done;
      done = t + 1;
      while (gap >= 255) {
        if (out) out[n] = 255;
        n++;
        gap -= 255;
      }
      if (out) out[n] = (uint8_t)gap;
      n++;
      gap = 0;
    }
    gap += nvalid - done;
  }
  return n;
}
}  // namespace

size_t dkb_zero_list_blocks(uint64_t n_positions) { return (size_t)((n_positions + ZL_BLOCK - 1) / ZL_BLOCK); }

int dkb_mask_to_zero_list(const uint32_t *mask1, uint64_t n_positions, uint32_t *zoff, uint8_t *zbytes,
                          size_t zbytes_cap, size_t *zbytes_used) {
  return no_throw([&]() -> int {
  if (!zbytes_used || (n_positions && (!mask1 || !zoff))) return DKB_EINVAL;
  const size_t nb = dkb_zero_list_blocks(n_positions);
  if (nb >= 0x7FFFFFFFu / 256) return DKB_EINVAL;
  // ONE pass over the flags: every thread codes a contiguous run of blocks into its own scratch
  // (gap coding, or the block's plain 256 bytes when that is shorter) and notes the lengths; the
  // offsets are their prefix sum, and each thread's scratch is then copied to its place in one piece.
  std::vector<uint32_t> len(nb);
  unsigned n_thr = std::thread::hardware_concurrency();
  if (n_thr > 32) n_thr = 32;
  if (n_thr > nb / 512) n_thr = (unsigned)(nb / 512);  // >= 1 M positions per thread
  if (const char *e = getenv("DKB_PACK_THREADS"))  // exactly n (up to 64), whatever the size
    if (atoi(e) > 0) n_thr = atoi(e) > 64 ? 64u : (unsigned)atoi(e);
  if (n_thr < 1) n_thr = 1;
  const size_t per = (nb + n_thr - 1) / n_thr;
  const bool fill = zbytes != nullptr;
  std::vector<std::vector<uint8_t>> scratch(n_thr);
  auto code = [&](unsigned t) {
    const size_t a = (size_t)t * per, b = a + per < nb ? a + per : nb;
    std::vector<uint8_t> &buf = scratch[t];
    size_t used = 0;
    uint8_t tmp[2048 + 16];
    if (fill && a < b) buf.resize((b - a) * 96 + 4096);
    for (size_t i = a; i < b; i++) {
      const uint64_t p0 = (uint64_t)i * ZL_BLOCK, p1 = p0 + ZL_BLOCK < n_positions ? p0 + ZL_BLOCK : n_positions;
      if (!fill) {
        const size_t g = zl_block_bytes(mask1, p0, p1, nullptr);
        len[i] = g > 256 ? (256u | ZL_RAW_BIT) : (uint32_t)g;
        continue;
      }
      if (buf.size() < used + 2048 + 16) buf.resize(buf.size() * 2);
      const size_t g = zl_block_bytes(mask1, p0, p1, tmp);
      uint8_t *dst = buf.data() + used;
      if (g > 256) {  // the block's 64 mask words, little-endian; flags past the stream are 0
        for (uint32_t wi = 0; wi < 64; wi++) {
          const uint64_t m = p0 / 32 + wi;
          uint32_t v = m * 32 < n_positions ? mask1[m] : 0u;
          if (m * 32 + 32 > n_positions && m * 32 < n_positions) v &= (1u << (n_positions - m * 32)) - 1u;
          for (int k = 0; k < 4; k++) dst[4 * wi + k] = (uint8_t)(v >> (8 * k));
        }
        len[i] = 256u | ZL_RAW_BIT;
        used += 256;
      } else {
        memcpy(dst, tmp, g);
        len[i] = (uint32_t)g;
        used += g;
      }
    }
  };
  if (!run_parallel(n_thr, code)) return DKB_ENOMEM;
  size_t total = 0;
  for (size_t i = 0; i < nb; i++) {
    zoff[i] = (uint32_t)total | (len[i] & ZL_RAW_BIT);
    total += len[i] & ~ZL_RAW_BIT;
  }
  if (n_positions) zoff[nb] = (uint32_t)total;
  *zbytes_used = total;
  if (!fill) return DKB_OK;  // sizing call
  if (total > zbytes_cap) return DKB_EINVAL;
  run_parallel(n_thr, [&](unsigned t) {
    const size_t a = (size_t)t * per, b = a + per < nb ? a + per : nb;
    if (a >= b) return;
    const size_t first = zoff[a] & ~ZL_RAW_BIT, last = b < nb ? (zoff[b] & ~ZL_RAW_BIT) : total;
    memcpy(zbytes + first, scratch[t].data(), last - first);
  });
  return DKB_OK;
  });
}

int dkb_variant_kmers(const char *const *left, const char *const *ref, const char *const *alt,
                      const char *const *right, size_t n_variants, int k, int drop_shared,
                      uint64_t *keys, uint32_t *variant_ids, uint8_t *allele_ids,
                      uint16_t *win_index, uint16_t *win_count, size_t *n_out) {
  return no_throw([&]() -> int {
  if (!n_out || k < DKB_MIN_K || k > DKB_MAX_K) return DKB_EINVAL;
  if (n_variants && (!left || !ref || !alt || !right)) return DKB_EINVAL;
  size_t n = 0;
  struct Win {
    uint64_t key;
    uint16_t idx;  // bit 15: haplotype window is the reverse complement of the key
  };
  std::vector<Win> per[2];
  std::unordered_set<uint64_t> seen[2];
  uint16_t run[2];
  std::string hap;
  for (size_t v = 0; v < n_variants; v++) {
    if (!left[v] || !ref[v] || !alt[v] || !right[v]) return DKB_EINVAL;
    const size_t ll = strlen(left[v]), rl = strlen(right[v]);
    const size_t lt = ll < (size_t)(k - 1) ? ll : (size_t)(k - 1);
    const size_t rt = rl < (size_t)(k - 1) ? rl : (size_t)(k - 1);
    for (int a = 0; a < 2; a++) {
      per[a].clear();
      seen[a].clear();
      run[a] = 0;
      const char *allele = a ? alt[v] : ref[v];
      const size_t al = strlen(allele);
      if (al > 30000) return DKB_EINVAL;
      hap.assign(left[v] + (ll - lt), lt);
      hap.append(allele, al);
      hap.append(right[v], rt);
      const long long hl = (long long)hap.size();
      if (hl < k) continue;
      // windows overlapping the allele; an empty allele needs both junction bases
      long long lo = (long long)lt - k + 1;
      long long hi = al ? (long long)(lt + al) - 1 : (long long)lt - 1;
      if (lo < 0) lo = 0;
      if (hi > hl - k) hi = hl - k;
      if (hi < lo) continue;
      run[a] = (uint16_t)(hi - lo + 1);
      // rolling encode over hap[lo .. hi+k)
      uint64_t fwd = 0;
      int good = 0;
      for (long long i = lo; i < hi + k; i++) {
        const uint8_t c = LUT.t[(uint8_t)hap[(size_t)i]];
        if (c > 3) {
          good = 0;
          fwd = 0;
          continue;
        }
        fwd = ((fwd << 2) | c) & kmask(k);
        if (++good >= k) {
          const long long w = i - k + 1;
          const uint64_t rc = revcomp(fwd, k);
          const uint64_t key = fwd < rc ? fwd : rc;
          if (seen[a].insert(key).second)
            per[a].push_back(Win{key, (uint16_t)((w - lo) | (fwd <= rc ? 0 : 0x8000))});
        }
      }
    }
    for (int a = 0; a < 2; a++) {
      for (const Win &w : per[a]) {
        if (drop_shared && seen[1 - a].count(w.key)) continue;
        if (keys) {
          keys[n] = w.key;
          if (variant_ids) variant_ids[n] = (uint32_t)v;
          if (allele_ids) allele_ids[n] = (uint8_t)a;
          if (win_index) win_index[n] = w.idx;
          if (win_count) win_count[n] = run[a];
        }
        n++;
      }
    }
  }
  *n_out = n;
  return DKB_OK;
  });
}

}  // extern "C"
