// Test infrastructure: the host packer (scalar, AVX2, AVX-512; ASCII and BAM 4-bit input; one and
// three threads) and the zero-list coder on EXACT-SIZE heap buffers under AddressSanitizer +
// UBSan - any out-of-bounds read or write, or undefined shift, aborts - and all packer paths
// compared word for word.  Built and run by tests/test_host.py::test_host_code_under_asan:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined ../denovo_kmer_b200/csrc/dkb_host.cpp asan_host.cpp -pthread
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "../include/dkb.h"
int main(int argc, char **argv) {
  const int n_iter = argc > 1 ? atoi(argv[1]) : 400;
  std::mt19937_64 g(7);
  const char *isas[] = {"0", "1", "2"};
  const char *thr[] = {"1", "3", "16"};
  for (int it = 0; it < n_iter; it++) {
    size_t n_reads = g() % 3000;
    int maxlen = 1 + g() % 300;
    std::vector<uint64_t> off(n_reads + 1);
    uint64_t base = g() % 5;  // offsets[0] != 0
    off[0] = base;
    for (size_t r = 0; r < n_reads; r++) off[r + 1] = off[r] + (g() % 10 == 0 ? 0 : g() % (maxlen + 1));
    if (it % 7 == 0) {  // a big one for the threaded split
      n_reads = 12000; off.resize(n_reads + 1); off[0] = 0;
      for (size_t r = 0; r < n_reads; r++) off[r + 1] = off[r] + 100 + g() % 52;
    }
    uint64_t total = off[n_reads];
    // exact-size heap buffers
    uint8_t *seq = (uint8_t *)malloc(total ? total : 1), *qual = (uint8_t *)malloc(total ? total : 1);
    for (uint64_t i = 0; i < total; i++) { seq[i] = "ACGTNacgtRY="[g() % 12]; qual[i] = g() % 256; }
    // 4-bit form
    std::vector<uint64_t> nb(n_reads + 1, 0);
    for (size_t r = 0; r < n_reads; r++) nb[r + 1] = nb[r] + (off[r + 1] - off[r] + 1) / 2;
    uint8_t *b4 = (uint8_t *)malloc(nb[n_reads] ? nb[n_reads] : 1);
    memset(b4, 0, nb[n_reads] ? nb[n_reads] : 1);
    for (size_t r = 0; r < n_reads; r++)
      for (uint64_t i = 0; i < off[r + 1] - off[r]; i++) {
        uint8_t ch = seq[off[r] + i] & 0xDF, c = ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 4 : ch == 'T' ? 8 : 15;
        b4[nb[r] + i / 2] |= (i & 1) ? c : (uint8_t)(c << 4);
      }
    uint64_t np = dkb_stream_positions(off.data(), n_reads);
    size_t bw = dkb_stream_bases_words(np), mw = dkb_stream_mask_words(np);
    uint32_t *ref_b = nullptr, *ref_m = nullptr;
    for (const char *t : thr) for (const char *isa : isas) for (int fmt = 0; fmt < 2; fmt++) for (int uq = 0; uq < 2; uq++) {
      setenv("DKB_PACK_THREADS", t, 1); setenv("DKB_PACK_ISA", isa, 1);
      uint32_t *b = (uint32_t *)malloc(bw ? bw * 4 : 4), *m = (uint32_t *)malloc(mw ? mw * 4 : 4);
      memset(b, 0xAB, bw * 4); memset(m, 0xCD, mw * 4);
      uint64_t out = 0;
      int mq = (int)(g() % 3 == 0 ? 0 : 20);
      if (!uq) mq = 20;
      // seq buffer for ASCII must be addressed from offsets[0]: the API indexes seq[offsets[r]]
      int rc = dkb_pack_reads_fmt(fmt ? b4 : seq - 0, fmt, uq ? qual : nullptr, off.data(), n_reads, uq ? 20 : mq, b, m, &out);
      if (rc != 0 || out != np) { printf("rc %d\n", rc); return 1; }
      if (uq) {
        if (!ref_b) { ref_b = b; ref_m = m; continue; }
        if (memcmp(ref_b, b, bw * 4) || memcmp(ref_m, m, mw * 4)) { printf("MISMATCH it %d thr %s isa %s fmt %d\n", it, t, isa, fmt); return 1; }
      }
      free(b); free(m);
    }
    // zero list on exact-size buffers
    if (ref_m) {
      size_t nblk = dkb_zero_list_blocks(np);
      uint32_t *zo = (uint32_t *)malloc((nblk + 1) * 4);
      size_t used = 0;
      dkb_mask_to_zero_list(ref_m, np, zo, nullptr, 0, &used);
      uint8_t *zb = (uint8_t *)malloc(used ? used : 1);
      size_t used2 = 0;
      int rc = dkb_mask_to_zero_list(ref_m, np, zo, zb, used, &used2);
      if (rc != 0 || used2 != used) { printf("zl rc %d\n", rc); return 1; }
      free(zo); free(zb);
    }
    free(ref_b); free(ref_m); free(seq); free(qual); free(b4);
  }
  // the variant k-mer builder: random SNVs / indels, short flanks, N in a flank, every k
  for (int it = 0; it < n_iter; it++) {
    const int k = 8 + (int)(g() % 24);
    const size_t nv = g() % 40;
    std::vector<std::string> L(nv), R(nv), A(nv), B(nv);
    auto rnd = [&](size_t n, bool with_n) {
      std::string x(n, 'A');
      for (auto &c : x) c = (with_n && g() % 50 == 0) ? 'N' : "ACGTacgt"[g() % 8];
      return x;
    };
    for (size_t v = 0; v < nv; v++) {
      L[v] = rnd(g() % 4 == 0 ? g() % (size_t)k : (size_t)k - 1 + g() % 5, true);
      R[v] = rnd(g() % 4 == 0 ? g() % (size_t)k : (size_t)k - 1 + g() % 5, true);
      A[v] = rnd(g() % 3 == 0 ? g() % 9 : 1, false);
      B[v] = rnd(g() % 3 == 0 ? g() % 9 : 1, false);
    }
    std::vector<const char *> lp(nv), rp(nv), ap(nv), bp(nv);
    for (size_t v = 0; v < nv; v++) { lp[v] = L[v].c_str(); rp[v] = R[v].c_str(); ap[v] = A[v].c_str(); bp[v] = B[v].c_str(); }
    size_t n = 0;
    const int drop = (int)(g() % 2);
    int rc = dkb_variant_kmers(lp.data(), ap.data(), bp.data(), rp.data(), nv, k, drop, nullptr, nullptr, nullptr, nullptr,
                               nullptr, &n);
    if (rc != 0) { printf("variant_kmers sizing rc %d\n", rc); return 1; }
    uint64_t *keys = (uint64_t *)malloc(n ? n * 8 : 1);
    uint32_t *var = (uint32_t *)malloc(n ? n * 4 : 1);
    uint8_t *al = (uint8_t *)malloc(n ? n : 1);
    uint16_t *wi = (uint16_t *)malloc(n ? n * 2 : 1), *wc = (uint16_t *)malloc(n ? n * 2 : 1);
    size_t n2 = 0;
    rc = dkb_variant_kmers(lp.data(), ap.data(), bp.data(), rp.data(), nv, k, drop, keys, var, al, wi, wc, &n2);
    if (rc != 0 || n2 != n) { printf("variant_kmers fill rc %d %zu %zu\n", rc, n, n2); return 1; }
    for (size_t i = 0; i < n; i++)
      if (var[i] >= nv || al[i] > 1 || (keys[i] >> (2 * k)) != 0 || (wi[i] & 0x7FFF) >= wc[i]) { printf("bad entry\n"); return 1; }
    free(keys); free(var); free(al); free(wi); free(wc);
  }
  printf("asan ok\n");
}
