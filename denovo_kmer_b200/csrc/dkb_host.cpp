// dkb_host.cpp — host side of the hot path's boundary: the pure k-mer
// primitives of src/kmer.rs, the read packer (decoded BAM records -> packed
// stream) and the variant k-mer builder (candidate alleles + flanks ->
// spanning k-mer entries for kernel 1).  The reference files are unmounted;
// the semantics are DESIGN.md §2.  No CUDA in this file.
#include <cstdint>
#include <cstring>
#include <string>
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

inline uint64_t kmask(int k) { return (1ull << (2 * k)) - 1; }

inline uint64_t revcomp(uint64_t fwd, int k) {
  // complement, then reverse the 2-bit groups of the 64-bit word and realign
  uint64_t x = ~fwd;
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = __builtin_bswap64(x);
  return x >> (64 - 2 * k);
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

int dkb_pack_reads(const uint8_t *seq, const uint8_t *qual, const uint64_t *offsets,
                   size_t n_reads, int min_baseq, uint32_t *bases2, uint32_t *mask1,
                   uint64_t *n_positions_out) {
  if (!offsets && n_reads) return DKB_EINVAL;
  if (!bases2 || !mask1) return DKB_EINVAL;
  const uint64_t n_pos = dkb_stream_positions(offsets, n_reads);
  const size_t bw = dkb_stream_bases_words(n_pos), mw = dkb_stream_mask_words(n_pos);
  memset(bases2, 0, bw * 4);
  memset(mask1, 0, mw * 4);
  uint64_t p = 0;
  for (size_t r = 0; r < n_reads; r++) {
    if (offsets[r + 1] < offsets[r]) return DKB_EINVAL;
    const uint8_t *s = seq + offsets[r];
    const uint8_t *q = qual ? qual + offsets[r] : nullptr;
    const size_t len = (size_t)(offsets[r + 1] - offsets[r]);
    for (size_t i = 0; i < len; i++, p++) {
      const uint8_t c = LUT.t[s[i]];
      if (c > 3 || (q && (int)q[i] < min_baseq)) continue;  // flag stays 0, base stays A
      bases2[p >> 4] |= (uint32_t)c << (2 * (p & 15));
      mask1[p >> 5] |= 1u << (p & 31);
    }
    p++;  // separator: flag 0
  }
  if (n_positions_out) *n_positions_out = n_pos;
  return DKB_OK;
}

int dkb_variant_kmers(const char *const *left, const char *const *ref, const char *const *alt,
                      const char *const *right, size_t n_variants, int k, int drop_shared,
                      uint64_t *keys, uint32_t *variant_ids, uint8_t *allele_ids,
                      uint16_t *win_index, uint16_t *win_count, size_t *n_out) {
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
}

}  // extern "C"
