/*
 * dkb.h — C ABI of the B200-native de novo k-mer hot path (libdkb.so).
 *
 * Drop-in boundary for jlanej/denovo_kmer's one data-parallel path:
 *   src/kmer.rs     k-mer extraction + canonical hashing from child/parent reads
 *   src/counter.rs  membership counting against each candidate allele's
 *                   spanning k-mer set, per sample (child / mother / father)
 *   (caller)        de novo support thresholds
 * Those files are NOT in the /root/reference mount (only
 * .github/workflows/ci.yml:1-50 and .gitignore:1 are); the semantics this
 * ABI implements are therefore the ones written down in DESIGN.md §2
 * ("spec"), which follows BASELINE.json `north_star`.  Parity is UNPINNED
 * against the real reference until its source is mounted.
 *
 * Conventions
 *   - every function returns DKB_OK (0) or a DKB_E* code; nothing throws or
 *     aborts across the ABI; dkb_last_error(ctx) has the detail string.
 *   - caller owns every host buffer; the library owns all device memory.
 *   - one context per GPU, one submitting thread per context.
 *   - there is NO CPU fallback: without a CUDA device dkb_ctx_create fails
 *     with DKB_ENODEV.
 *   - base codes: A=0 C=1 G=2 T=3.  A k-mer value ("key") packs the first
 *     base in the most significant of its 2k bits; canonical = min(fwd, rc).
 *
 * Packed read batch ("stream") — what src/kmer.rs would iterate, as bits:
 *   bases2 : uint32 words, 16 bases each, base p in bits 2*(p%16) of word p/16
 *   mask1  : uint32 words, 32 flags each, flag p in bit  p%32     of word p/32
 *            flag = 1 iff base p is A/C/G/T and its base quality >= min_baseq
 *   every read is followed by ONE separator position with flag 0, so the
 *   rolling window resets at read ends exactly like an N does.
 *   n_positions = sum(read lengths) + n_reads.  A k-mer at position w counts
 *   iff flags w..w+k-1 are all 1.
 */
#ifndef DKB_H
#define DKB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DKB_ABI_VERSION 1
#define DKB_MIN_K 8
#define DKB_MAX_K 31
#define DKB_N_SAMPLES 3 /* 0 child, 1 mother, 2 father */
#define DKB_N_ALLELES 2 /* 0 ref, 1 alt (multi-allelic sites are split by the host) */

enum {
  DKB_OK = 0,
  DKB_EINVAL = 1, /* bad argument */
  DKB_ECUDA = 2,  /* CUDA runtime error (see dkb_last_error) */
  DKB_ENOMEM = 3, /* host or device allocation failed */
  DKB_ESTATE = 4, /* call out of order (e.g. submit before table_build) */
  DKB_ENODEV = 5, /* no usable sm_100 device; there is no CPU fallback */
  DKB_ENCCL = 6   /* NCCL error, or libnccl.so.2 cannot be loaded (multi-GPU calls only) */
};

typedef struct dkb_ctx dkb_ctx;

/* De novo support rule (integer only).  A variant is called (bit 0 of the
 * call byte) iff
 *   child  alt hits     >= min_child_alt_hits      and
 *   child  alt distinct >= min_child_alt_distinct  and
 *   mother alt hits     <= max_parent_alt_hits     and
 *   father alt hits     <= max_parent_alt_hits     and
 *   mother ref hits     >= min_parent_ref_hits     and
 *   father ref hits     >= min_parent_ref_hits.
 * Failing clauses set bits 1..4 (see DKB_CALL_*). */
typedef struct dkb_thresholds {
  uint32_t min_child_alt_hits;
  uint32_t min_child_alt_distinct;
  uint32_t max_parent_alt_hits;
  uint32_t min_parent_ref_hits;
} dkb_thresholds;

#define DKB_CALL_DENOVO 0x01u         /* all clauses hold */
#define DKB_CALL_CHILD_LOW 0x02u      /* child alt support below minimum */
#define DKB_CALL_MOTHER_ALT 0x04u     /* mother carries alt k-mers */
#define DKB_CALL_FATHER_ALT 0x08u     /* father carries alt k-mers */
#define DKB_CALL_PARENT_UNCOVERED 0x10u /* a parent lacks ref k-mer coverage */

/* Scan tuning, normally left to the library (dkb_ctx_set_tuning(ctx, NULL)). */
typedef struct dkb_tuning {
  int seed_len;     /* s: 8..15 and <= k - stride + 1; 0 = auto */
  int stride;       /* D: probe every D-th stream position (1, 2, 4, 8 or 16); 0 = auto */
  int bloom_hashes; /* 1..4 bits per seed in the seed filter; 0 = auto */
  int filter_mode;  /* 1 = filter in shared memory (small candidate sets), 2 = filter in L2
                       (strides 2..16, 1..2 hashes; the library decides whether a 145 KB
                       (37 120-word) shared-memory pre-filter goes in front of it, env
                       DKB_PREFILTER_WORDS overrides); 0 = auto */
} dkb_tuning;

/* ---- library ---------------------------------------------------------- */
int dkb_abi_version(void);
const char *dkb_strerror(int code);
const char *dkb_last_error(const dkb_ctx *ctx);

/* ---- context ---------------------------------------------------------- */
int dkb_ctx_create(int device, int k, dkb_ctx **out);
int dkb_ctx_destroy(dkb_ctx *ctx);
int dkb_ctx_set_tuning(dkb_ctx *ctx, const dkb_tuning *tuning);
int dkb_ctx_get_tuning(const dkb_ctx *ctx, dkb_tuning *out);

/* ---- kmer.rs host-side primitives (pure functions, no device) ----------- */
/* seq[0..k) -> forward key; DKB_EINVAL if a base is not A/C/G/T (either case) */
int dkb_kmer_encode(const char *seq, int k, uint64_t *fwd_out);
uint64_t dkb_kmer_revcomp(uint64_t fwd, int k);
uint64_t dkb_kmer_canonical(uint64_t fwd, int k);

/* ---- host packer: decoded reads -> packed stream ------------------------ */
/* reads are concatenated in seq/qual; read r occupies [offsets[r], offsets[r+1]).
 * qual may be NULL (all qualities pass).  Sizes: */
uint64_t dkb_stream_positions(const uint64_t *offsets, size_t n_reads);
size_t dkb_stream_bases_words(uint64_t n_positions);
size_t dkb_stream_mask_words(uint64_t n_positions);
int dkb_pack_reads(const uint8_t *seq, const uint8_t *qual, const uint64_t *offsets,
                   size_t n_reads, int min_baseq, uint32_t *bases2, uint32_t *mask1,
                   uint64_t *n_positions_out);
/* The same from either form the BAM layer holds (seq_format as in dkb_batch_submit_reads): 0 =
 * ASCII bases, read r at seq[offsets[r]..offsets[r+1]); 1 = BAM 4-bit codes (=ACMGRSVTWYHKDBN,
 * high nibble first: record.seq().encoded of rust-htslib, no decode pass), every read starting
 * on a byte boundary with (len + 1) / 2 bytes, reads back to back from seq[0].  qual is one byte
 * per base at qual[offsets[r]..] either way.  Same stream, bit for bit, for the same reads. */
int dkb_pack_reads_fmt(const uint8_t *seq, int seq_format, const uint8_t *qual, const uint64_t *offsets,
                       size_t n_reads, int min_baseq, uint32_t *bases2, uint32_t *mask1,
                       uint64_t *n_positions_out);

/* ---- the flags as a ZERO LIST (a third less PCIe traffic) ----------------------- */
/* mask1 costs 1 bit per position; with ~4 % of the positions unusable (N, low quality, read
 * separators) it is cheaper to send WHERE the zeros are.  Per block of 2048 positions: a byte
 * string, each byte g < 255 = "g usable positions, then one unusable one", 255 = "255 usable
 * positions" (no zero); positions after the last byte's zero are usable.  A block whose string
 * would exceed 256 bytes is stored as its 256 bytes of plain flag bits instead.  zoff[b] =
 * byte offset of block b's string in zbytes, bit 31 set for a plain-bits block; zoff[n_blocks]
 * = total bytes.  Flags at positions >= n_positions are not coded (the scan ignores them).
 * dkb_mask_to_zero_list converts dense flags (zbytes == NULL: only sizes: *zbytes_used and
 * zoff are filled in); zoff holds dkb_zero_list_blocks(n) + 1 entries. */
size_t dkb_zero_list_blocks(uint64_t n_positions);
int dkb_mask_to_zero_list(const uint32_t *mask1, uint64_t n_positions, uint32_t *zoff, uint8_t *zbytes,
                          size_t zbytes_cap, size_t *zbytes_used);

/* ---- host variant k-mer builder (counter.rs "spanning k-mer set") ------- */
/* For variant v with left flank L, right flank R (reference bases either side
 * of REF, at least k-1 each unless the contig ends) and alleles REF/ALT,
 * haplotype(a) = L[-(k-1):] + allele(a) + R[:k-1]; its spanning k-mers are the
 * windows that overlap the allele, or that straddle the junction when the
 * allele is empty.  Windows with a non-ACGT base are skipped.  Keys present
 * in both alleles of one variant are dropped when drop_shared != 0.  Within one
 * (variant, allele) a key is emitted once (first window wins).
 * Pass keys == NULL to size the output; *n_out returns the entry count.
 * win_index/win_count (may be NULL) give each entry's window index within its
 * haplotype's spanning run and that run's length (seed-ladder hints for
 * dkb_table_build). */
int dkb_variant_kmers(const char *const *left, const char *const *ref,
                      const char *const *alt, const char *const *right, size_t n_variants,
                      int k, int drop_shared, uint64_t *keys, uint32_t *variant_ids,
                      uint8_t *allele_ids, uint16_t *win_index, uint16_t *win_count,
                      size_t *n_out);

/* ---- kernel 1: table build ------------------------------------------------ */
/* Entry i = (canonical key, variant id, allele id); counts are reported per
 * entry index.  A (key, variant, allele) triple repeated in the input keeps
 * its first entry live; later repeats stay at 0.  One key may belong to many
 * (variant, allele) owners: every owner is counted.  win_index/win_count may
 * be NULL. */
int dkb_table_build(dkb_ctx *ctx, const uint64_t *keys, const uint32_t *variant_ids,
                    const uint8_t *allele_ids, const uint16_t *win_index,
                    const uint16_t *win_count, size_t n_entries, uint32_t n_variants);

/* ---- kernel 2: streaming extract-and-probe ------------------------------- */
/* Host buffers (pinned for full speed); H2D copy + scan are queued on the
 * context's streams and overlap with the next submit.  The buffers must stay
 * valid until dkb_sync. */
int dkb_batch_submit(dkb_ctx *ctx, const uint32_t *bases2, const uint32_t *mask1,
                     uint64_t n_positions, int sample);
/* Same as dkb_batch_submit with the flags as a zero list (above): bases2 + zoff + zbytes cross
 * PCIe - 0.29 instead of 0.38 bytes per base at 4 % unusable positions - and the library
 * expands the list into mask1 on the device before the scan. */
int dkb_batch_submit_sparse(dkb_ctx *ctx, const uint32_t *bases2, const uint32_t *zoff,
                            const uint8_t *zbytes, size_t zbytes_used, uint64_t n_positions, int sample);
/* Decoded reads as the BAM layer holds them: the packing (what dkb_pack_reads does on the
 * host) runs on the GPU.  seq_format 0: ASCII bases, read r at seq[offsets[r]..offsets[r+1]);
 * 1: BAM 4-bit codes (=ACMGRSVTWYHKDBN, high nibble first), every read starting on a byte
 * boundary with (len + 1) / 2 bytes, reads back to back from seq[0].  qual (may be NULL):
 * one byte per base at qual[offsets[r]..].  Same stream, bit for bit, as dkb_pack_reads. */
int dkb_batch_submit_reads(dkb_ctx *ctx, const uint8_t *seq, int seq_format, const uint8_t *qual,
                           const uint64_t *offsets, size_t n_reads, int min_baseq, int sample);
/* Device-resident buffers: d_bases2 16-byte aligned, d_mask1 4-byte aligned, each
 * dkb_stream_*_words long.  Flags at positions >= n_positions are ignored. */
int dkb_batch_submit_device(dkb_ctx *ctx, const uint32_t *d_bases2, const uint32_t *d_mask1,
                            uint64_t n_positions, int sample);
/* Up to DKB_MAX_MULTI device-resident streams (e.g. the three samples of a trio) in ONE kernel
 * launch; batch i goes into the counters of samples[i].  Same alignment rules; a batch of 0
 * positions is skipped.  Worth it for short batches, where a launch costs as much as the scan. */
#define DKB_MAX_MULTI 4
int dkb_batch_submit_device_multi(dkb_ctx *ctx, int n_batches, const uint32_t *const *d_bases2,
                                  const uint32_t *const *d_mask1, const uint64_t *n_positions,
                                  const int *samples);
int dkb_sync(dkb_ctx *ctx);
int dkb_counts_reset(dkb_ctx *ctx);

/* per-entry counts, layout [DKB_N_SAMPLES][n_entries] */
int dkb_entry_counts_fetch(dkb_ctx *ctx, uint32_t *out);
/* device pointer of the same array, for the multi-GPU sum (NCCL allreduce) */
int dkb_entry_counts_device(dkb_ctx *ctx, void **d_ptr, size_t *n_u32);

/* ---- kernel 3: finalise ---------------------------------------------------- */
int dkb_finalise(dkb_ctx *ctx, const dkb_thresholds *thr);
/* Same, from a caller-owned DEVICE copy of the counters ([DKB_N_SAMPLES][n_entries], the
   layout of dkb_entry_counts_device) - e.g. a copy that is being summed over ranks while
   the context's own counters already take the next batch.  Runs on the scan stream: order
   it after the copy is complete (dkb_scan_stream). */
int dkb_finalise_from(dkb_ctx *ctx, const dkb_thresholds *thr, const uint32_t *d_counts);
/* hits / distinct: [n_variants][DKB_N_ALLELES][DKB_N_SAMPLES]; n_kmers:
 * [n_variants][DKB_N_ALLELES]; calls: [n_variants].  Any pointer may be NULL. */
int dkb_results_fetch(dkb_ctx *ctx, uint32_t *hits, uint32_t *distinct, uint32_t *n_kmers,
                      uint8_t *calls);

/* ---- pinned staging next to the GPU ------------------------------------------ */
/* Page-locked host memory for the packed batches, allocated on the NUMA node the GPU's PCIe
 * slot belongs to (the calling thread is moved there for the allocation and moved back).  With
 * eight ranks pushing batches at once, buffers on the wrong socket halve the H2D rate.
 * dkb_thread_bind_near_gpu pins the CALLING thread (the one that packs and submits) to that
 * node's CPUs for good; *numa_node_out (may be NULL) gets the node, -1 if unknown. */
int dkb_host_alloc(dkb_ctx *ctx, size_t bytes, void **out);
int dkb_host_free(dkb_ctx *ctx, void *p);
int dkb_thread_bind_near_gpu(dkb_ctx *ctx, int *numa_node_out);

/* ---- multi-GPU: read batches sharded over ranks, ONE sum of the counters ------ */
/* One process (or thread) per GPU, each with its own context and its own slice of the read
 * batches; the table is built identically on every rank.  The per-entry counters are summed
 * over ranks with a single NCCL allreduce (libnccl.so.2 is loaded on first use; without it
 * these calls return DKB_ENCCL and everything else works).
 *   rank 0: dkb_comm_unique_id(id)  -> send the DKB_COMM_ID_BYTES to every rank (MPI, TCP, a file ...)
 *   all   : dkb_comm_init(ctx, id, rank, world)
 * Simple form: scan, then dkb_counts_allreduce (in place, on the scan stream), dkb_finalise.
 * Overlapped form, for a sequence of independent batches (trios, regions): after a batch's
 * scans dkb_reduce_push snapshots the counters and starts the sum on a side stream; the
 * caller resets the counters and submits the next batch at once; the push of batch i also
 * queues kernel 3 for batch i-1 behind the scans just submitted, by which time that sum has
 * long finished.  dkb_reduce_flush finalises the last batch; dkb_results_fetch then returns
 * its results and dkb_reduced_counts_fetch its summed counters.  With no communicator (one
 * GPU) both forms run without NCCL and equal plain dkb_finalise. */
#define DKB_COMM_ID_BYTES 128
int dkb_comm_unique_id(void *id_out);
int dkb_comm_init(dkb_ctx *ctx, const void *id, int rank, int world);
int dkb_comm_destroy(dkb_ctx *ctx);
/* rank, communicator size as NCCL reports it, NCCL version (0 without a communicator) */
int dkb_comm_info(const dkb_ctx *ctx, int *rank, int *world, int *nccl_version);
int dkb_counts_allreduce(dkb_ctx *ctx);
int dkb_reduce_push(dkb_ctx *ctx, const dkb_thresholds *thr);
int dkb_reduce_flush(dkb_ctx *ctx, const dkb_thresholds *thr);
int dkb_reduced_counts_fetch(dkb_ctx *ctx, uint32_t *out);

/* ---- introspection (bench / tests) --------------------------------------- */
typedef struct dkb_stats {
  uint64_t n_entries, n_live_entries, table_slots;
  uint64_t n_seeds, seed_slots, bloom_words, bloom_bits_set;
  uint64_t scan_launches;    /* scan kernels launched since create */
  uint64_t positions_scanned;
  uint64_t bloom_hits, seed_hits, windows_probed, window_hits; /* if profiling counters on */
  uint64_t scan_launches_timed; /* launches folded into scan_ms_total */
  double scan_ms_total;      /* sum of per-launch CUDA-event times of the scan kernel */
  float last_scan_ms;        /* CUDA-event time of the most recent finished scan kernel */
  uint32_t prefilter_words;  /* L2 filter mode: words of the shared-memory pre-filter, 0 = none */
  uint32_t gated_lookups;    /* 1: the scan also reads the flag stream in full and skips the lookups of
                                seeds that hold an unusable base (env DKB_GATE=0/1 overrides) */
} dkb_stats;
int dkb_stats_get(dkb_ctx *ctx, dkb_stats *out);
int dkb_profile_counters(dkb_ctx *ctx, int enable);
/* CUDA stream the scan kernels are launched on (cudaStream_t as void*). */
int dkb_scan_stream(dkb_ctx *ctx, void **stream_out);

#ifdef __cplusplus
}
#endif
#endif /* DKB_H */
