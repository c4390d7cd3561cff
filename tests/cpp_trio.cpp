// C++ host-side use of the boundary (include/dkb.hpp): a three-read toy trio.  Exits 0
// with "no-gpu" when the context cannot be created (CPU-only test run), otherwise checks
// the counts of the known-answer case of tests/test_oracle.py on the GPU.
#include <cstdio>
#include <cstring>

#include "dkb.hpp"

int main() {
  using namespace dkbxx;
  const int k = 8;
  // one SNV: ...ACGTACGT [G>T] TTGACCAT...
  std::vector<Candidate> cands = {{"CCATGACGTACGT", "G", "T", "TTGACCATGGCA"}};
  Entries e = variant_kmers(cands, k);
  if (e.keys.size() != 16 || e.n_variants != 1) return 2;
  const std::string ref = "CCATGACGTACGTGTTGACCATGGCA", alt = "CCATGACGTACGTTTTGACCATGGCA";
  std::vector<uint8_t> seq(ref.begin(), ref.end());
  seq.insert(seq.end(), alt.begin(), alt.end());
  std::vector<uint64_t> off = {0, ref.size(), ref.size() + alt.size()};
  Stream s = pack_reads(seq, {}, off, 0);
  if (s.n_positions != ref.size() + alt.size() + 2) return 3;
  try {
    Counter c(k);
    c.build_table(e);
    c.submit(s, 0);
    std::vector<uint32_t> counts = c.entry_counts();
    // child read 1 carries every ref k-mer once, read 2 every alt k-mer once
    for (size_t i = 0; i < e.keys.size(); i++)
      if (counts[i] != 1) return 4;
    Results r = c.finalise({1, 1, 0, 0});
    if (r.hits[0] != 8 || r.hits[3] != 8 || r.calls[0] != DKB_CALL_DENOVO) return 5;
    // the same reads with the flags as a zero list, into the mother's counters; no communicator:
    // the overlapped reduction runs without NCCL and must equal plain finalise
    Counter::SparseStream z = Counter::to_sparse(s);
    if (z.zbytes.size() != 2) return 7;  // two separators, nothing else unusable
    c.submit(z, 1);
    c.reduce_push({1, 1, 0, 0});
    c.reduce_flush({1, 1, 0, 0});
    std::vector<uint32_t> both = c.entry_counts();
    for (size_t i = 0; i < e.keys.size(); i++)
      if (both[i] != 1 || both[e.keys.size() + i] != 1) return 8;
    std::puts("gpu-ok");
  } catch (const Error &err) {
    if (err.code != DKB_ENODEV) {
      std::fprintf(stderr, "%s\n", err.what());
      return 6;
    }
    std::puts("no-gpu");
  }
  return 0;
}
