// dkb_scan.cuh — kernel 2: streaming extract-and-probe over a packed read
// stream (the GPU form of src/kmer.rs's per-read k-mer iteration feeding
// src/counter.rs's membership counting; both unmounted, DESIGN.md §2 is the spec).
//
// Every stream position p with p % D == 0 has its s-mer tested against a
// seed filter held in shared memory — or, for very large candidate tables, in
// L2 — (stage A, the only per-position work).
// Filter hits are compacted per warp tile and verified, 32 at a time, against
// the exact seed table in L2 (stage B, one 16-byte bucket load per hit, issued
// one tile ahead of its use).  A verified seed carries the offsets j at which
// some table key designates it; each window w = p - j is rebuilt from the
// stream, validated against the mask, canonicalised and probed in the key
// table (stage C).  A slot is counted only when its own designated offset for
// class (j % D) equals j, so a matching window is counted exactly once however
// many seeds it contains (proof in DESIGN.md §4).
#pragma once
#include "dkb_device.cuh"

namespace dkb {

template <int D, int NH, bool GF, bool PROF>
struct ScanWarp {
  // Strides 8 and 16 leave 8 / 4 lookups per lane in a 2048-position tile; SUB such tiles
  // form a macro tile with 32 lookups per lane so that hit handling and loop overhead are
  // paid once per macro tile.
  static constexpr bool MACRO = D >= 8;
  static constexpr int LPT = 64 / D;                 // lookups per lane and (sub-)tile
  static constexpr int SUB = MACRO ? 32 / LPT : 1;   // sub-tiles per macro tile
  const ScanParams &P;
  const uint32_t *filt;
  uint16_t *hl;  // filter-hit ids of the current tile: lane << 6 | lookup index
  uint64_t *cq;  // ring of verified seeds: seed-table slot << 32 | position
  uint32_t ch = 0, ct = 0;
  int lane;
  uint32_t lt_mask;
  uint64_t keep = l2_policy_evict_last();  // cache policy of every table load
  // stage B probes in flight (issued at the end of one tile, consumed in the next)
  uint32_t pend_n = 0, pend_x = 0, pend_p = 0, pend_b = 0;
  uint4 pend_bucket = {0, 0, 0, 0};
  unsigned long long n_bloom = 0, n_seed = 0, n_probe = 0, n_hit = 0;

  __device__ __forceinline__ ScanWarp(const ScanParams &p, const uint32_t *f, uint16_t *h,
                                      uint64_t *c, int l)
      : P(p), filt(f), hl(h), cq(c), lane(l), lt_mask((1u << l) - 1) {}

  __device__ __forceinline__ uint32_t ld_bases(uint32_t wi) const {
    return wi < P.n_bwords ? __ldg(P.bases + wi) : 0u;
  }
  __device__ __forceinline__ uint32_t ld_mask(uint32_t wi) const {
    return wi < P.n_mwords ? __ldg(P.mask + wi) : 0u;
  }

  // ---- stage C: every candidate window of up to n verified seeds ----------------
  // One lane per verified seed.  All windows of a seed at p lie in
  // [p - (k - s), p + k): their bases and mask bits are fetched once (5 + 3
  // words, issued together with the offset bitmap), then each window costs one
  // key-bucket load.
  __device__ __forceinline__ void stage_c(uint32_t n) {
    if ((uint32_t)lane < n) {
      const uint64_t e = cq[(ch + lane) & (CQ_CAP - 1)];
      const uint32_t p = (uint32_t)e;
      const int k = P.k, E = k - P.s;
      const uint32_t start = p > (uint32_t)E ? p - (uint32_t)E : 0u;
      const uint32_t bw0 = start >> 4, mw0 = start >> 5;
      uint32_t info = ldg_u32_hint(P.st.sinfo + (uint32_t)(e >> 32), keep);
      uint32_t b[5], m[3];
#pragma unroll
      for (int i = 0; i < 5; i++) b[i] = ld_bases(bw0 + i);
#pragma unroll
      for (int i = 0; i < 3; i++) m[i] = ld_mask(mw0 + i);
      const uint64_t km = kmer_mask(k);
      const uint32_t vm = (1u << k) - 1;  // k <= 31
      while (info) {
        const int j = __ffs(info) - 1;
        info &= info - 1;
        if (p < (uint32_t)j) continue;
        const uint32_t w = p - (uint32_t)j;
        if (w + (uint32_t)k > P.n_pos) continue;
        if (PROF) n_probe++;
        // validity: mask bits w .. w+k-1 must all be set
        const uint32_t mo = w - (mw0 << 5);  // < 64
        const uint32_t mbits = mo < 32 ? __funnelshift_r(m[0], m[1], mo)
                                       : __funnelshift_r(m[1], m[2], mo - 32);
        if ((mbits & vm) != vm) continue;
        // the window's bases in stream order (first base least significant)
        const uint32_t bo = w - (bw0 << 4);  // < 48
        uint32_t x0 = b[0], x1 = b[1], x2 = b[2];
        if (bo >= 16) { x0 = b[1]; x1 = b[2]; x2 = b[3]; }
        if (bo >= 32) { x0 = b[2]; x1 = b[3]; x2 = b[4]; }
        const uint32_t sh = 2 * (bo & 15);
        const uint64_t v =
            ((uint64_t)__funnelshift_r(x1, x2, sh) << 32 | __funnelshift_r(x0, x1, sh)) & km;
        const uint64_t fwd = base_reverse(v, k);
        const uint64_t rc = ~v & km;  // complement of the stream-order value IS the rc key
        const uint64_t key = fwd <= rc ? fwd : rc;
        const int ori = fwd <= rc ? 0 : 1;
        uint32_t bk = key_bucket(key, P.kt.bucket_mask);
        while (true) {
          const uint4 *bp = P.kt.slots + bk * KBUCKET;
          const uint4 s0 = ldg_v4_hint(bp, keep), s1 = ldg_v4_hint(bp + 1, keep);
          const uint64_t k0 = slot_key(s0), k1 = slot_key(s1);
          if (k0 == key && s0.z != ENTRY_DEAD && slot_offset(s0.w, ori, j % D, D) == (uint32_t)j) {
            atomicAdd(P.counts + s0.z, 1u);
            if (PROF) n_hit++;
          }
          if (k1 == key && s1.z != ENTRY_DEAD && slot_offset(s1.w, ori, j % D, D) == (uint32_t)j) {
            atomicAdd(P.counts + s1.z, 1u);
            if (PROF) n_hit++;
          }
          if (k0 == KEY_EMPTY || k1 == KEY_EMPTY) break;  // bucket not full: nothing spilled
          bk = (bk + 1) & P.kt.bucket_mask;
        }
      }
    }
    ch += n;
    __syncwarp();
  }

  // ---- stage B, second half: use the seed buckets loaded one tile ago -----------
  __device__ __forceinline__ void consume_pending() {
    if (pend_n == 0) return;
    const bool act = (uint32_t)lane < pend_n;
    const uint4 bk = pend_bucket;
    const uint32_t x = pend_x;
    int q = -1;
    if (bk.w == x) q = 3;
    if (bk.z == x) q = 2;
    if (bk.y == x) q = 1;
    if (bk.x == x) q = 0;
    bool found = act && q >= 0;
    uint32_t slot = pend_b * BUCKET + q;
    // Full home bucket without the seed (about 1 probe in 600): it may have
    // spilled into the following buckets.
    const bool full = bk.x != SEED_EMPTY && bk.y != SEED_EMPTY && bk.z != SEED_EMPTY &&
                      bk.w != SEED_EMPTY;
    if (__any_sync(FULL_MASK, act && !found && full)) {
      if (act && !found && full) {
        uint32_t b = pend_b;
        while (true) {
          b = (b + 1) & P.st.bucket_mask;
          const uint4 nb = ldg_v4_hint(reinterpret_cast<const uint4 *>(P.st.seeds) + b, keep);
          const uint32_t v[4] = {nb.x, nb.y, nb.z, nb.w};
          bool open = false;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            if (v[i] == x) { found = true; slot = b * BUCKET + i; }
            if (v[i] == SEED_EMPTY) open = true;
          }
          if (found || open) break;
        }
      }
    }
    pend_n = 0;
    const uint32_t bal = __ballot_sync(FULL_MASK, found);
    if (bal) {
      if (found) cq[(ct + __popc(bal & lt_mask)) & (CQ_CAP - 1)] = (uint64_t)slot << 32 | pend_p;
      ct += __popc(bal);
      if (PROF && found) n_seed++;
      __syncwarp();
      if (ct - ch >= 32) stage_c(32);
    }
  }

  // ---- stage B, first half: start the exact check of hits [first, first + n) ------
  // Each lane takes one hit id, pulls the 16 bases at that position out of the
  // owning lane's registers with shuffles (no trip back to L2) and issues the
  // load of the seed's bucket.  tile_base = stream position of lane 0's chunk.
  __device__ __forceinline__ void issue_probes(uint32_t first, uint32_t n, const uint32_t (&w)[5],
                                               uint32_t tile_base) {
    const bool act = (uint32_t)lane < n;
    const uint32_t id = act ? hl[first + lane] : (uint32_t)lane << 6;
    const int src = id >> 6;
    if constexpr (MACRO) {
      // macro tiles (strides 8, 16): the lookup index runs over SUB sub-tiles whose words
      // are no longer in registers; hits are rare here, so re-read them from L2
      const uint32_t idx = id & 63;
      pend_p = tile_base + (idx / LPT) * WTILE + src * CHUNK + (idx % LPT) * D;
      const uint32_t wi = pend_p >> 4;
      pend_x = __funnelshift_r(ld_bases(wi), ld_bases(wi + 1), 2 * (pend_p & 15)) & P.seed_mask;
      pend_b = seed_bucket(pend_x, P.st.shift);
      if (act) pend_bucket = ldg_v4_hint(reinterpret_cast<const uint4 *>(P.st.seeds) + pend_b, keep);
      pend_n = n;
      return;
    }
    const uint32_t q = (id & 63) * D;  // offset inside the owning lane's chunk
    const uint32_t c = q >> 4;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; i++) v[i] = __shfl_sync(FULL_MASK, w[i], src);
    uint32_t lo = v[0], hi = v[1];
    if (c == 1) { lo = v[1]; hi = v[2]; }
    if (c == 2) { lo = v[2]; hi = v[3]; }
    if (c == 3) { lo = v[3]; hi = v[4]; }
    pend_x = __funnelshift_r(lo, hi, 2 * (q & 15)) & P.seed_mask;
    pend_p = tile_base + src * CHUNK + q;
    pend_b = seed_bucket(pend_x, P.st.shift);
    if (act) pend_bucket = ldg_v4_hint(reinterpret_cast<const uint4 *>(P.st.seeds) + pend_b, keep);
    pend_n = n;
  }

  // ---- compact this tile's filter hits and start their verification ---------------
  // bit 31 of acc0 is lookup 0; acc1 (D == 1 only) continues at lookup 32.
  __device__ __forceinline__ void handle_hits(uint32_t acc0, uint32_t acc1,
                                              const uint32_t (&w)[5], uint32_t tile_base) {
    const uint32_t cnt = __popc(acc0) + (D == 1 ? __popc(acc1) : 0);
    if (PROF) n_bloom += cnt;
    // Exclusive prefix sum of the per-lane hit counts.  Counts are almost always
    // below 8, so three ballots of their bit planes replace five dependent shuffles.
    const uint32_t b0 = __ballot_sync(FULL_MASK, cnt & 1), b1 = __ballot_sync(FULL_MASK, cnt & 2),
                   b2 = __ballot_sync(FULL_MASK, cnt & 4), bhi = __ballot_sync(FULL_MASK, cnt > 7);
    if ((b0 | b1 | b2 | bhi) == 0) return;
    uint32_t total, excl;
    if (bhi == 0) {
      total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
      excl = __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
    } else {
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      total = __shfl_sync(FULL_MASK, incl, 31);
      excl = incl - cnt;
    }
    const uint32_t tag = (uint32_t)lane << 6;
    if constexpr (!MACRO) {
      if (total <= (uint32_t)HL_CAP) {
        // common case: the whole tile fits the id list, no bounds checks
        uint32_t idx = excl;
        uint32_t a = acc0;
        while (a) {
          const int i = __clz(a);
          a ^= 0x80000000u >> i;
          hl[idx++] = (uint16_t)(tag + i);
        }
        if (D == 1) {
          a = acc1;
          while (a) {
            const int i = __clz(a);
            a ^= 0x80000000u >> i;
            hl[idx++] = (uint16_t)(tag + 32 + i);
          }
        }
        __syncwarp();
        for (uint32_t r = 0; r < total; r += 32) {
          consume_pending();  // at most one batch of probes in flight
          issue_probes(r, min(total - r, 32u), w, tile_base);
        }
        __syncwarp();
        return;
      }
    }
    // the id list holds HL_CAP hits; denser tiles (low-complexity sequence) take more passes
    for (uint32_t base = 0; base < total; base += HL_CAP) {
      uint32_t idx = excl - base;  // wraps below zero for hits of earlier passes
      uint32_t a = acc0;
      while (a) {
        const int i = __clz(a);
        a ^= 0x80000000u >> i;
        if (idx < (uint32_t)HL_CAP) hl[idx] = (uint16_t)(tag + i);
        idx++;
      }
      if (D == 1) {
        a = acc1;
        while (a) {
          const int i = __clz(a);
          a ^= 0x80000000u >> i;
          if (idx < (uint32_t)HL_CAP) hl[idx] = (uint16_t)(tag + 32 + i);
          idx++;
        }
      }
      __syncwarp();
      const uint32_t here = min(total - base, (uint32_t)HL_CAP);
      for (uint32_t r = 0; r < here; r += 32) {
        consume_pending();  // at most one batch of probes in flight
        issue_probes(r, min(here - r, 32u), w, tile_base);
      }
      __syncwarp();
    }
  }

  __device__ __forceinline__ void drain() {
    consume_pending();
    while (ct != ch) stage_c(min(ct - ch, 32u));
  }

  // ---- stage A, macro-tile form: LPT lookups of one sub-tile, hit bits in the low bits,
  // first lookup highest.  At stride 16 every seed lies inside one word (s <= 15): no
  // cut, no halo.
  __device__ __forceinline__ uint32_t stage_a_sub(const uint32_t (&w)[5]) const {
    const uint32_t mult = P.seed_mult;
    const uint32_t fbase = (uint32_t)__cvta_generic_to_shared(filt);
    uint32_t a = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
      for (int t = 0; t < 16; t += D) {
        const uint32_t x = t ? __funnelshift_r(w[c], w[c + 1], 2 * t) : w[c];
        const uint32_t h = x * mult;
        uint32_t word;
        if constexpr (GF) {
          word = ldg_u32_hint(P.bloom + __umulhi(h, P.bloom_words), keep);
        } else {
          const uint32_t addr = __umulhi(h, (uint32_t)BLOOM_WORDS) * P.four + fbase;
          asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(addr));
        }
        uint32_t bit = word << (x & 31);
        if (NH >= 2) bit &= word << (__umulhi(h, SEED_MULT2) & 31);
        if (NH >= 3) bit &= word << (__umulhi(h, SEED_MULT3) & 31);
        if (NH >= 4) bit &= word << (__umulhi(h, SEED_MULT4) & 31);
        a = __funnelshift_l(bit, a, 1);
      }
    }
    return a;
  }

  // ---- stage A: the shared-memory seed filter over one lane chunk ---------------
  // w[0..3] hold this lane's 64 positions, w[4] the next 16 (halo for s-mers
  // that start in the chunk and end beyond it).
  __device__ __forceinline__ void stage_a(const uint32_t (&w)[5], uint32_t &acc0,
                                          uint32_t &acc1) const {
    const uint32_t mult = P.seed_mult;
    const uint32_t fbase = (uint32_t)__cvta_generic_to_shared(filt);
    uint32_t part[4];  // per-word hit bits (16 / D each): four short dependency chains
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t a = 0;
#pragma unroll
      for (int t = 0; t < 16; t += D) {
        const uint32_t x = t ? __funnelshift_r(w[c], w[c + 1], 2 * t) : w[c];
        const uint32_t h = x * mult;
        uint32_t word;
        if constexpr (GF) {
          // large candidate sets: the filter does not fit in shared memory; probe it in L2
          word = ldg_u32_hint(P.bloom + __umulhi(h, P.bloom_words), keep);
        } else {
          // byte address = idx * 4 + base as a multiply-add (P.four is opaque), not an ALU-pipe LEA
          const uint32_t addr = __umulhi(h, (uint32_t)BLOOM_WORDS) * P.four + fbase;
          asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(addr));
        }
        uint32_t bit = word << (x & 31);
        if (NH >= 2) bit &= word << (__umulhi(h, SEED_MULT2) & 31);
        if (NH >= 3) bit &= word << (__umulhi(h, SEED_MULT3) & 31);
        if (NH >= 4) bit &= word << (__umulhi(h, SEED_MULT4) & 31);
        a = __funnelshift_l(bit, a, 1);
      }
      part[c] = a;
    }
    // lookup 0 ends up in bit 31 of acc0
    constexpr int NB = 16 / D;  // hit bits per word
    if constexpr (D == 1) {
      acc0 = part[0] << 16 | part[1];
      acc1 = part[2] << 16 | part[3];
    } else {
      acc0 = (part[0] << (3 * NB) | part[1] << (2 * NB) | part[2] << NB | part[3]) << (32 - 4 * NB);
      acc1 = 0;
    }
  }
};

// tile -> registers.  Streaming (evict-first) loads: the stream is read once
// and must not push the seed / key tables out of L2.
__device__ __forceinline__ void load_tile(const ScanParams &P, uint32_t tile, int lane,
                                          uint32_t (&w)[5]) {
  const uint32_t wi = tile * WTILE_WORDS + lane * 4;
  if (wi + 4 <= P.n_bwords) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(P.bases + wi));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = wi + i < P.n_bwords ? P.bases[wi + i] : 0u;
  }
  // w[4] of lane 31 = first word of the next warp tile; the other lanes get
  // theirs from lane + 1 in halo_finish(), AFTER the loads have landed, so the
  // prefetch of the next tile never stalls on a shuffle.
  w[4] = 0;
  if (lane == 31) w[4] = wi + 4 < P.n_bwords ? __ldg(P.bases + wi + 4) : 0u;
}

__device__ __forceinline__ void halo_finish(int lane, uint32_t (&w)[5]) {
  const uint32_t up = __shfl_down_sync(FULL_MASK, w[0], 1);
  if (lane != 31) w[4] = up;
}

template <int D, int NH, bool GF, bool PROF>
__global__ void __launch_bounds__(SCAN_THREADS, 1) k_scan(const ScanParams P) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t *filt = smem;
  uint64_t *cq_all = reinterpret_cast<uint64_t *>(smem + BLOOM_WORDS);
  uint16_t *hl_all = reinterpret_cast<uint16_t *>(cq_all + SCAN_WARPS * CQ_CAP);

  if constexpr (!GF) {
    for (int i = threadIdx.x; i < BLOOM_WORDS / 4; i += SCAN_THREADS)
      reinterpret_cast<uint4 *>(filt)[i] = __ldg(reinterpret_cast<const uint4 *>(P.bloom) + i);
    __syncthreads();
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ScanWarp<D, NH, GF, PROF> W(P, filt, hl_all + warp * HL_CAP, cq_all + warp * CQ_CAP, lane);

  const uint32_t n_warps = gridDim.x * SCAN_WARPS;
  if constexpr (ScanWarp<D, NH, GF, PROF>::MACRO) {
    // Strides 8 / 16: a macro tile = SUB sub-tiles of 2048 positions (32 lookups per lane),
    // read in groups of GS sub-tiles (GS LDG.128 per lane in flight, 1-2 KB per warp), with
    // the next group prefetched while the current one is filtered.
    constexpr int SUB = ScanWarp<D, NH, GF, PROF>::SUB, LPT = ScanWarp<D, NH, GF, PROF>::LPT;
    constexpr int GS = 2, GPM = SUB / GS;  // sub-tiles per group, groups per macro tile
    constexpr bool HALO = D < 16;          // at stride 16 every seed lies inside one word
    const uint32_t n_macro = (P.n_tiles + SUB - 1) / SUB;
    auto load_group = [&](uint32_t t0, uint4 (&v)[GS], uint32_t &edge) {
      const uint32_t wb = t0 * WTILE_WORDS + lane * 4;
      if ((t0 + GS) * WTILE_WORDS + 4 <= P.n_bwords) {
#pragma unroll
        for (int j = 0; j < GS; j++)  // normal L2 priority: filter hits re-read these sectors soon
          v[j] = __ldg(reinterpret_cast<const uint4 *>(P.bases + wb + j * WTILE_WORDS));
        edge = 0;
        if (HALO && lane == 31) edge = __ldg(P.bases + (t0 + GS) * WTILE_WORDS);
      } else {  // the stream ends inside this group
#pragma unroll
        for (int j = 0; j < GS; j++) {
          const uint32_t q = wb + j * WTILE_WORDS;
          v[j].x = q < P.n_bwords ? P.bases[q] : 0u;
          v[j].y = q + 1 < P.n_bwords ? P.bases[q + 1] : 0u;
          v[j].z = q + 2 < P.n_bwords ? P.bases[q + 2] : 0u;
          v[j].w = q + 3 < P.n_bwords ? P.bases[q + 3] : 0u;
        }
        const uint32_t e = (t0 + GS) * WTILE_WORDS;
        edge = HALO && e < P.n_bwords ? P.bases[e] : 0u;
      }
    };
    uint32_t macro = warp * gridDim.x + blockIdx.x;  // consecutive units on different SMs
    int g = 0;
    uint4 nxt[GS];
    uint32_t nxt_edge = 0;
    if (macro < n_macro) load_group(macro * SUB, nxt, nxt_edge);
    uint32_t acc = 0;
    while (macro < n_macro) {
      uint4 v[GS];
#pragma unroll
      for (int j = 0; j < GS; j++) v[j] = nxt[j];
      const uint32_t edge = nxt_edge;
      const uint32_t cur_macro = macro;
      const bool last_group = g == GPM - 1;
      if (last_group) { macro += n_warps; g = 0; } else { g++; }
      if (macro < n_macro) load_group(macro * SUB + g * GS, nxt, nxt_edge);
#pragma unroll
      for (int j = 0; j < GS; j++) {
        uint32_t w[5] = {v[j].x, v[j].y, v[j].z, v[j].w, 0};
        if constexpr (HALO) {
          // lanes 0..30 take lane + 1's first word; lane 31 takes the first word of the
          // next sub-tile (lane 0's, or the group's edge word)
          const uint32_t up = __shfl_down_sync(FULL_MASK, w[0], 1);
          const uint32_t wrap = j + 1 < GS ? __shfl_sync(FULL_MASK, v[j + 1 < GS ? j + 1 : j].x, 0)
                                           : edge;
          w[4] = lane == 31 ? wrap : up;
        }
        acc = acc << LPT | W.stage_a_sub(w);
      }
      if (last_group) {
        W.consume_pending();
        const uint32_t w0[5] = {0, 0, 0, 0, 0};
        W.handle_hits(acc, 0, w0, cur_macro * SUB * WTILE);
        acc = 0;
      }
    }
  } else {
  uint32_t tile = warp * gridDim.x + blockIdx.x;  // consecutive tiles on different SMs
  uint32_t nxt[5];
  if (tile < P.n_tiles) load_tile(P, tile, lane, nxt);
  for (; tile < P.n_tiles; tile += n_warps) {
    uint32_t w[5];
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = nxt[i];
    if (tile + n_warps < P.n_tiles) load_tile(P, tile + n_warps, lane, nxt);
    halo_finish(lane, w);
    uint32_t acc0, acc1;
    W.stage_a(w, acc0, acc1);
    W.consume_pending();  // the previous tile's seed buckets have had a whole stage A to arrive
    W.handle_hits(acc0, acc1, w, tile * WTILE);
  }
  }
  W.drain();

  if (PROF) {
    unsigned long long v[4] = {W.n_bloom, W.n_seed, W.n_probe, W.n_hit};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      unsigned long long s = v[i];
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
      if (lane == 0 && s) atomicAdd(P.prof + i, s);
    }
  }
}

}  // namespace dkb
