/*
 * dnk_oracle.c — CPU restatement of the de novo k-mer hot path.  TEST
 * INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it; nothing under
 * denovo_kmer_b200/ may import, link or call it.
 *
 * PARITY UNPINNED.  The files this should follow line by line —
 * jlanej/denovo_kmer src/kmer.rs (k-mer extraction + canonical hashing) and
 * src/counter.rs (spanning k-mer sets + membership counting) — are not in the
 * /root/reference mount (only .github/workflows/ci.yml:1-50 and
 * .gitignore:1 are), and there is no Rust toolchain here, so neither golden
 * vectors nor reference outputs exist to pin it.  It restates the semantics in
 * DESIGN.md §2, which follow BASELINE.json `north_star`; each function names
 * the reference component it stands in for.  It is written the plain way a
 * CPU tool does it (per read, rolling window, hash map) and shares no code or
 * data structure with the CUDA path.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_NO_ENTRY 0xFFFFFFFFu

/* ---- kmer.rs: base encoding ------------------------------------------- */
/* A=0 C=1 G=2 T=3, either case; everything else (N, IUPAC, '=') is -1. */
int orc_base_code(uint8_t ch) {
  switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

static inline uint64_t kmask(int k) { return (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1); }

/* kmer.rs: reverse complement of a packed k-mer (first base most significant). */
uint64_t orc_revcomp(uint64_t fwd, int k) {
  uint64_t rc = 0;
  for (int i = 0; i < k; i++) {
    rc = (rc << 2) | (3 - (fwd & 3));
    fwd >>= 2;
  }
  return rc;
}

/* kmer.rs: canonical form = numeric minimum of the k-mer and its reverse complement. */
uint64_t orc_canonical(uint64_t fwd, int k) {
  uint64_t rc = orc_revcomp(fwd, k);
  return fwd < rc ? fwd : rc;
}

/* kmer.rs: all canonical k-mers of one read.  A base that is not A/C/G/T or
 * whose quality is below min_bq resets the window, so no emitted k-mer covers
 * it.  out/out_pos (either may be NULL) receive the key and window start.
 * Returns the number of k-mers. */
size_t orc_read_kmers(const uint8_t *seq, const uint8_t *qual, size_t len, int k, int min_bq,
                      uint64_t *out, uint32_t *out_pos) {
  uint64_t fwd = 0, rc = 0;
  const uint64_t mask = kmask(k);
  int run = 0; /* consecutive usable bases ending here */
  size_t n = 0;
  for (size_t i = 0; i < len; i++) {
    int c = orc_base_code(seq[i]);
    if (c < 0 || (qual && (int)qual[i] < min_bq)) {
      run = 0;
      fwd = rc = 0;
      continue;
    }
    fwd = ((fwd << 2) | (uint64_t)c) & mask;
    rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
    if (++run >= k) {
      if (out) out[n] = fwd < rc ? fwd : rc;
      if (out_pos) out_pos[n] = (uint32_t)(i + 1 - (size_t)k);
      n++;
    }
  }
  return n;
}

/* ---- counter.rs: the k-mer -> owners map -------------------------------- */
typedef struct orc_slot {
  uint64_t key; /* valid when ent != ORC_NO_ENTRY */
  uint32_t ent; /* entry index or ORC_NO_ENTRY for empty */
  uint32_t pad;
} orc_slot;

typedef struct orc_set {
  size_t n_entries, cap; /* cap is a power of two */
  orc_slot *slots;
  uint8_t *live; /* [n_entries] 0 for a repeated (key, owner) triple */
} orc_set;

static inline size_t orc_hash(uint64_t x, size_t cap) {
  x ^= x >> 31;
  x *= 0x9E3779B97F4A7C15ull;
  x ^= x >> 29;
  return (size_t)x & (cap - 1);
}

/* counter.rs: build the map from (key, variant, allele) entries.  One key may
 * have several owners; a repeated triple keeps only its first entry live. */
orc_set *orc_set_build(const uint64_t *keys, const uint32_t *variant, const uint8_t *allele,
                       size_t n) {
  orc_set *s = (orc_set *)calloc(1, sizeof(orc_set));
  if (!s) return NULL;
  size_t cap = 16;
  while (cap < 2 * n + 2) cap <<= 1;
  s->n_entries = n;
  s->cap = cap;
  s->slots = (orc_slot *)malloc(cap * sizeof(orc_slot));
  s->live = (uint8_t *)malloc(n ? n : 1);
  if (!s->slots || !s->live) return NULL;
  for (size_t i = 0; i < cap; i++) s->slots[i].ent = ORC_NO_ENTRY;
  for (size_t i = 0; i < n; i++) {
    size_t h = orc_hash(keys[i], cap);
    int dup = 0;
    while (s->slots[h].ent != ORC_NO_ENTRY) {
      uint32_t e = s->slots[h].ent;
      if (s->slots[h].key == keys[i] && variant[e] == variant[i] && allele[e] == allele[i]) {
        dup = 1;
        break;
      }
      h = (h + 1) & (cap - 1);
    }
    s->live[i] = (uint8_t)!dup;
    if (!dup) {
      s->slots[h].key = keys[i];
      s->slots[h].ent = (uint32_t)i;
    }
  }
  return s;
}

void orc_set_free(orc_set *s) {
  if (!s) return;
  free(s->slots);
  free(s->live);
  free(s);
}

void orc_set_live(const orc_set *s, uint8_t *live_out) { memcpy(live_out, s->live, s->n_entries); }

/* counter.rs: count, for every entry, how many read k-mers equal its key.
 * counts[n_entries] is ADDED to.  Reads are split over n_threads threads; matches are a
 * fraction of a percent of the k-mers, so the threads add into the one shared array with
 * atomic increments (no per-thread copy of the counters to allocate and merge per call). */
void orc_count_reads(const orc_set *s, const uint8_t *seq, const uint8_t *qual,
                     const uint64_t *offsets, size_t n_reads, int k, int min_bq, uint64_t *counts,
                     int n_threads) {
  const uint64_t mask = kmask(k);
  const size_t cap = s->cap;
#ifdef _OPENMP
  if (n_threads < 1) n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1024) num_threads(n_threads)
  for (long long r = 0; r < (long long)n_reads; r++) {
    const uint8_t *rs = seq + offsets[r];
    const uint8_t *rq = qual ? qual + offsets[r] : NULL;
    size_t len = (size_t)(offsets[r + 1] - offsets[r]);
    uint64_t fwd = 0, rc = 0;
    int run = 0;
    for (size_t i = 0; i < len; i++) {
      int c = orc_base_code(rs[i]);
      if (c < 0 || (rq && (int)rq[i] < min_bq)) {
        run = 0;
        fwd = rc = 0;
        continue;
      }
      fwd = ((fwd << 2) | (uint64_t)c) & mask;
      rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
      if (++run >= k) {
        uint64_t key = fwd < rc ? fwd : rc;
        size_t h = orc_hash(key, cap);
        while (s->slots[h].ent != ORC_NO_ENTRY) {
          if (s->slots[h].key == key) {
#pragma omp atomic
            counts[s->slots[h].ent]++;
          }
          h = (h + 1) & (cap - 1);
        }
      }
    }
  }
}

/* The same count taken from a PACKED read stream (include/dkb.h: 2-bit bases, 16 per
 * uint32 word; 1-bit usable flags, 32 per word; one flag-0 separator after every read) -
 * what the host packer hands to the device.  A flag-0 position resets the window exactly
 * like an N or a low-quality base does in orc_read_kmers, so for a stream packed from reads
 * this equals orc_count_reads on those reads (tests/test_oracle.py checks that).  Lets the
 * full-size parity tests and bench.py check the GPU on the very streams it scanned.
 * Positions are cut into blocks; a block rolls in from k-1 positions before its start. */
void orc_count_stream(const orc_set *s, const uint32_t *bases2, const uint32_t *mask1,
                      uint64_t n_positions, int k, uint64_t *counts, int n_threads) {
  const uint64_t mask = kmask(k);
  const size_t cap = s->cap;
  const uint64_t block = 1u << 20;
  const long long n_blocks = (long long)((n_positions + block - 1) / block);
#ifdef _OPENMP
  if (n_threads < 1) n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
  for (long long b = 0; b < n_blocks; b++) {
    const uint64_t lo = (uint64_t)b * block;
    const uint64_t hi = lo + block < n_positions ? lo + block : n_positions;
    uint64_t p = lo >= (uint64_t)(k - 1) ? lo - (uint64_t)(k - 1) : 0; /* roll-in */
    uint64_t fwd = 0, rc = 0;
    int run = 0;
    for (; p < hi; p++) {
      if (!((mask1[p >> 5] >> (p & 31)) & 1u)) {
        run = 0;
        fwd = rc = 0;
        continue;
      }
      const uint64_t c = (bases2[p >> 4] >> (2 * (p & 15))) & 3u;
      fwd = ((fwd << 2) | c) & mask;
      rc = (rc >> 2) | ((3 - c) << (2 * (k - 1)));
      if (++run >= k && p >= lo) { /* the window ENDING at p belongs to this block */
        uint64_t key = fwd < rc ? fwd : rc;
        size_t h = orc_hash(key, cap);
        while (s->slots[h].ent != ORC_NO_ENTRY) {
          if (s->slots[h].key == key) {
#pragma omp atomic
            counts[s->slots[h].ent]++;
          }
          h = (h + 1) & (cap - 1);
        }
      }
    }
  }
}

/* counter.rs: per-variant, per-allele, per-sample summary of entry counts.
 * counts: [3][n_entries]; hits/distinct: [n_variants][2][3]; n_kmers: [n_variants][2]. */
void orc_variant_stats(const orc_set *s, const uint32_t *variant, const uint8_t *allele,
                       const uint64_t *counts, size_t n_variants, uint64_t *hits,
                       uint64_t *distinct, uint32_t *n_kmers) {
  memset(hits, 0, n_variants * 6 * sizeof(uint64_t));
  memset(distinct, 0, n_variants * 6 * sizeof(uint64_t));
  memset(n_kmers, 0, n_variants * 2 * sizeof(uint32_t));
  for (size_t e = 0; e < s->n_entries; e++) {
    if (!s->live[e]) continue;
    size_t va = (size_t)variant[e] * 2 + allele[e];
    n_kmers[va]++;
    for (int smp = 0; smp < 3; smp++) {
      uint64_t c = counts[(size_t)smp * s->n_entries + e];
      hits[va * 3 + smp] += c;
      distinct[va * 3 + smp] += (c > 0);
    }
  }
}

/* caller: de novo support thresholds (include/dkb.h dkb_thresholds). */
void orc_calls(const uint64_t *hits, const uint64_t *distinct, size_t n_variants,
               uint32_t min_child_alt_hits, uint32_t min_child_alt_distinct,
               uint32_t max_parent_alt_hits, uint32_t min_parent_ref_hits, uint8_t *calls) {
  for (size_t v = 0; v < n_variants; v++) {
    const uint64_t *h = hits + v * 6, *d = distinct + v * 6;
    /* h[allele*3 + sample] */
    uint8_t c = 0;
    if (h[3 + 0] < min_child_alt_hits || d[3 + 0] < min_child_alt_distinct) c |= 0x02;
    if (h[3 + 1] > max_parent_alt_hits) c |= 0x04;
    if (h[3 + 2] > max_parent_alt_hits) c |= 0x08;
    if (h[0 + 1] < min_parent_ref_hits || h[0 + 2] < min_parent_ref_hits) c |= 0x10;
    if (c == 0) c = 0x01;
    calls[v] = c;
  }
}

/* counter.rs: spanning k-mers of one allele of one variant.
 * hap = last (k-1) of left + allele + first (k-1) of right; windows that
 * overlap the allele (or straddle the junction if it is empty); windows with
 * a non-ACGT base skipped.  out_keys / out_win (may be NULL) receive canonical
 * keys and the window's index inside the spanning run.  *run_len gets the run
 * length.  Returns the number of keys (duplicates included). */
size_t orc_allele_kmers(const char *left, const char *allele, const char *right, int k,
                        uint64_t *out_keys, uint16_t *out_win, uint16_t *run_len) {
  size_t ll = strlen(left), al = strlen(allele), rl = strlen(right);
  size_t lt = ll < (size_t)(k - 1) ? ll : (size_t)(k - 1);
  size_t rt = rl < (size_t)(k - 1) ? rl : (size_t)(k - 1);
  size_t hl = lt + al + rt;
  char *hap = (char *)malloc(hl + 1);
  memcpy(hap, left + (ll - lt), lt);
  memcpy(hap + lt, allele, al);
  memcpy(hap + lt + al, right, rt);
  hap[hl] = 0;
  size_t n = 0;
  if (run_len) *run_len = 0;
  if (hl >= (size_t)k) {
    /* window [w, w+k) overlaps allele [lt, lt+al): w < lt+al and w+k > lt.
     * empty allele: must contain both hap[lt-1] and hap[lt]: w <= lt-1 and w+k > lt. */
    long long lo, hi;
    if (al > 0) {
      lo = (long long)lt - k + 1;
      hi = (long long)(lt + al) - 1;
    } else {
      lo = (long long)lt - k + 1;
      hi = (long long)lt - 1;
    }
    if (lo < 0) lo = 0;
    if (hi > (long long)hl - k) hi = (long long)hl - k;
    if (hi >= lo && run_len) *run_len = (uint16_t)(hi - lo + 1);
    for (long long w = lo; w <= hi; w++) {
      uint64_t fwd = 0;
      int ok = 1;
      for (int i = 0; i < k; i++) {
        int c = orc_base_code((uint8_t)hap[w + i]);
        if (c < 0) {
          ok = 0;
          break;
        }
        fwd = (fwd << 2) | (uint64_t)c;
      }
      if (!ok) continue;
      if (out_keys) out_keys[n] = orc_canonical(fwd, k);
      if (out_win) out_win[n] = (uint16_t)(w - lo);
      n++;
    }
  }
  free(hap);
  return n;
}
