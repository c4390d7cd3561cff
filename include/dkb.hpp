// dkb.hpp — header-only C++17 convenience layer over the C ABI in dkb.h: RAII context,
// std::vector in/out, exceptions instead of error codes.  Nothing here computes; every
// call forwards to libdkb.so.  (The reference is a Rust crate — INTEGRATION.md shows its
// binding; this is the same surface for C++ hosts and for this repo's own C++ tests.)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "dkb.h"

namespace dkbxx {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc, const dkb_ctx *ctx = nullptr) {
  if (rc != DKB_OK)
    throw Error(rc, std::string(dkb_strerror(rc)) + ": " + dkb_last_error(ctx));
}

// Spanning k-mer entries of a candidate list (counter.rs: per-allele k-mer sets).
struct Entries {
  std::vector<uint64_t> keys;
  std::vector<uint32_t> variant;
  std::vector<uint8_t> allele;
  std::vector<uint16_t> win_index, win_count;
  uint32_t n_variants = 0;
};

struct Candidate {
  std::string left, ref, alt, right;  // flanks hold at least k-1 reference bases each
};

inline Entries variant_kmers(const std::vector<Candidate> &c, int k, bool drop_shared = true) {
  std::vector<const char *> l, r, a, rt;
  for (const auto &v : c) {
    l.push_back(v.left.c_str());
    r.push_back(v.ref.c_str());
    a.push_back(v.alt.c_str());
    rt.push_back(v.right.c_str());
  }
  size_t n = 0;
  check(dkb_variant_kmers(l.data(), r.data(), a.data(), rt.data(), c.size(), k, drop_shared, nullptr,
                          nullptr, nullptr, nullptr, nullptr, &n));
  Entries e;
  e.keys.resize(n);
  e.variant.resize(n);
  e.allele.resize(n);
  e.win_index.resize(n);
  e.win_count.resize(n);
  e.n_variants = (uint32_t)c.size();
  check(dkb_variant_kmers(l.data(), r.data(), a.data(), rt.data(), c.size(), k, drop_shared,
                          e.keys.data(), e.variant.data(), e.allele.data(), e.win_index.data(),
                          e.win_count.data(), &n));
  return e;
}

// Packed read stream (kmer.rs: what the per-read k-mer iteration consumes).
struct Stream {
  std::vector<uint32_t> bases2, mask1;
  uint64_t n_positions = 0;
};

// four_bit: seq holds BAM 4-bit codes (high nibble first, every read on a byte boundary)
inline Stream pack_reads(const std::vector<uint8_t> &seq, const std::vector<uint8_t> &qual,
                         const std::vector<uint64_t> &offsets, int min_baseq, bool four_bit = false) {
  const size_t n_reads = offsets.empty() ? 0 : offsets.size() - 1;
  Stream s;
  s.n_positions = dkb_stream_positions(offsets.data(), n_reads);
  s.bases2.resize(dkb_stream_bases_words(s.n_positions) + 1);
  s.mask1.resize(dkb_stream_mask_words(s.n_positions) + 1);
  check(dkb_pack_reads_fmt(seq.data(), four_bit ? 1 : 0, qual.empty() ? nullptr : qual.data(), offsets.data(),
                           n_reads, min_baseq, s.bases2.data(), s.mask1.data(), &s.n_positions));
  return s;
}

struct Results {
  std::vector<uint32_t> hits, distinct;  // [n_variants][2 alleles][3 samples]
  std::vector<uint32_t> n_kmers;         // [n_variants][2]
  std::vector<uint8_t> calls;            // [n_variants], DKB_CALL_* bits
};

// One GPU context.  Throws Error{DKB_ENODEV} when no sm_100 device exists: there is no
// CPU fallback.
class Counter {
 public:
  Counter(int k, int device = 0) { check(dkb_ctx_create(device, k, &ctx_)); }
  ~Counter() { dkb_ctx_destroy(ctx_); }
  Counter(const Counter &) = delete;
  Counter &operator=(const Counter &) = delete;

  void set_tuning(const dkb_tuning &t) { check(dkb_ctx_set_tuning(ctx_, &t), ctx_); }
  void build_table(const Entries &e) {
    check(dkb_table_build(ctx_, e.keys.data(), e.variant.data(), e.allele.data(),
                          e.win_index.empty() ? nullptr : e.win_index.data(),
                          e.win_count.empty() ? nullptr : e.win_count.data(), e.keys.size(),
                          e.n_variants),
          ctx_);
    n_entries_ = e.keys.size();
    n_variants_ = e.n_variants;
  }
  // the stream must outlive the next sync()
  void submit(const Stream &s, int sample) {
    check(dkb_batch_submit(ctx_, s.bases2.data(), s.mask1.data(), s.n_positions, sample), ctx_);
  }
  // decoded reads packed on the GPU (ASCII bases, or BAM 4-bit codes when four_bit)
  void submit_reads(const std::vector<uint8_t> &seq, const std::vector<uint8_t> &qual,
                    const std::vector<uint64_t> &offsets, int min_baseq, int sample,
                    bool four_bit = false) {
    check(dkb_batch_submit_reads(ctx_, seq.data(), four_bit ? 1 : 0, qual.empty() ? nullptr : qual.data(),
                                 offsets.data(), offsets.empty() ? 0 : offsets.size() - 1, min_baseq,
                                 sample),
          ctx_);
  }
  // the stream with its flags as a zero list (dkb_mask_to_zero_list): less PCIe traffic;
  // the vectors must outlive the next sync()
  struct SparseStream {
    std::vector<uint32_t> bases2, zoff;
    std::vector<uint8_t> zbytes;
    uint64_t n_positions = 0;
  };
  static SparseStream to_sparse(const Stream &s) {
    SparseStream z;
    z.bases2 = s.bases2;
    z.n_positions = s.n_positions;
    z.zoff.resize(dkb_zero_list_blocks(s.n_positions) + 1);
    size_t used = 0;
    check(dkb_mask_to_zero_list(s.mask1.data(), s.n_positions, z.zoff.data(), nullptr, 0, &used));
    z.zbytes.resize(used + 1);
    check(dkb_mask_to_zero_list(s.mask1.data(), s.n_positions, z.zoff.data(), z.zbytes.data(), used, &used));
    z.zbytes.resize(used);
    return z;
  }
  void submit(const SparseStream &z, int sample) {
    check(dkb_batch_submit_sparse(ctx_, z.bases2.data(), z.zoff.data(), z.zbytes.data(), z.zbytes.size(),
                                  z.n_positions, sample),
          ctx_);
  }
  // multi-GPU: one context per GPU and rank; id from dkb_comm_unique_id on rank 0
  void comm_init(const void *id, int rank, int world) { check(dkb_comm_init(ctx_, id, rank, world), ctx_); }
  void counts_allreduce() { check(dkb_counts_allreduce(ctx_), ctx_); }
  void reduce_push(const dkb_thresholds &t) { check(dkb_reduce_push(ctx_, &t), ctx_); }
  void reduce_flush(const dkb_thresholds &t) { check(dkb_reduce_flush(ctx_, &t), ctx_); }
  void sync() { check(dkb_sync(ctx_), ctx_); }
  void reset_counts() { check(dkb_counts_reset(ctx_), ctx_); }
  std::vector<uint32_t> entry_counts() {  // [3 samples][n_entries]
    std::vector<uint32_t> out(3 * n_entries_ + 1);
    check(dkb_entry_counts_fetch(ctx_, out.data()), ctx_);
    out.resize(3 * n_entries_);
    return out;
  }
  // d_counts: optional caller-owned DEVICE copy of the counters (dkb_finalise_from)
  Results finalise(const dkb_thresholds &t, const uint32_t *d_counts = nullptr) {
    check(d_counts ? dkb_finalise_from(ctx_, &t, d_counts) : dkb_finalise(ctx_, &t), ctx_);
    Results r;
    r.hits.resize(6 * (size_t)n_variants_ + 1);
    r.distinct.resize(6 * (size_t)n_variants_ + 1);
    r.n_kmers.resize(2 * (size_t)n_variants_ + 1);
    r.calls.resize((size_t)n_variants_ + 1);
    check(dkb_results_fetch(ctx_, r.hits.data(), r.distinct.data(), r.n_kmers.data(), r.calls.data()),
          ctx_);
    r.hits.resize(6 * (size_t)n_variants_);
    r.distinct.resize(6 * (size_t)n_variants_);
    r.n_kmers.resize(2 * (size_t)n_variants_);
    r.calls.resize(n_variants_);
    return r;
  }
  dkb_ctx *raw() { return ctx_; }

 private:
  dkb_ctx *ctx_ = nullptr;
  size_t n_entries_ = 0;
  uint32_t n_variants_ = 0;
};

}  // namespace dkbxx
