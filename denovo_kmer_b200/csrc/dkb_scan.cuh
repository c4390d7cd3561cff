// dkb_scan.cuh — kernel 2: streaming extract-and-probe over a packed read
// stream (the GPU form of src/kmer.rs's per-read k-mer iteration feeding
// src/counter.rs's membership counting; both unmounted, DESIGN.md §2 is the spec).
//
// Every stream position p with p % D == 0 has its s-mer tested against a
// seed filter held in shared memory (stage A, the only per-position work).
// Filter hits are compacted per warp and verified, 32 at a time, against the
// exact seed table in L2 (stage B); a verified seed carries the offsets j at
// which some table key designates it, and each window w = p - j is then
// rebuilt from the stream, validated against the mask, canonicalised and
// probed in the key table (stage C).  A slot is counted only when its own
// designated offset for class (j % D) equals j, so a matching window is
// counted exactly once however many seeds it contains (proof in DESIGN.md §4).
#pragma once
#include "dkb_device.cuh"

namespace dkb {

template <int D, int NH, bool PROF>
struct ScanWarp {
  const ScanParams &P;
  const uint32_t *filt;
  uint16_t *hl;  // ring of filter-hit ids of the current tile: lane << 6 | lookup index
  uint64_t *cq;  // ring of verified seeds: offset bitmap << 32 | position
  uint32_t hh = 0, ht = 0, ch = 0, ct = 0;
  int lane;
  uint32_t lt_mask;
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

  // ---- stage C: windows of up to n verified seeds ---------------------------
  __device__ __forceinline__ void stage_c(uint32_t n) {
    if ((uint32_t)lane < n) {
      const uint64_t e = cq[(ch + lane) & (CQ_CAP - 1)];
      const uint32_t p = (uint32_t)e;
      uint32_t info = (uint32_t)(e >> 32);
      const int k = P.k;
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
        const uint32_t mi = w >> 5, ms = w & 31;
        const uint64_t m64 = ((uint64_t)ld_mask(mi + 1) << 32 | ld_mask(mi)) >> ms;
        if (((uint32_t)m64 & vm) != vm) continue;
        // the window's bases in stream order (first base least significant)
        const uint32_t wi = w >> 4, sh = 2 * (w & 15);
        const uint64_t lo = (uint64_t)ld_bases(wi + 1) << 32 | ld_bases(wi);
        uint64_t v = lo >> sh;
        if (sh) v |= (uint64_t)ld_bases(wi + 2) << (64 - sh);
        v &= km;
        const uint64_t fwd = base_reverse(v, k);
        const uint64_t rc = ~v & km;  // complement of the stream-order value IS the rc key
        const uint64_t key = fwd <= rc ? fwd : rc;
        const int sel = (fwd <= rc ? 0 : 32) + 5 * (j % D);
        uint32_t slot = (uint32_t)mix64(key) & P.table_mask;
        while (true) {
          const uint64_t tk = __ldg(P.tkeys + slot);
          if (tk == KEY_EMPTY) break;
          if (tk == key) {
            const uint32_t want = (uint32_t)(__ldg(P.toffs + slot) >> sel) & 31u;
            const uint32_t ent = __ldg(P.tentry + slot);
            if (want == (uint32_t)j && ent != ENTRY_DEAD) {
              atomicAdd(P.counts + ent, 1u);
              if (PROF) n_hit++;
            }
          }
          slot = (slot + 1) & P.table_mask;
        }
      }
    }
    ch += n;
    __syncwarp();
  }

  // ---- stage B: exact check of up to n filter hits of the current tile ------------
  // Each lane takes one hit id, pulls the 16 bases at that position out of the
  // owning lane's registers with shuffles (no trip back to L2) and probes the
  // seed table.  tile_base = stream position of lane 0's chunk.
  __device__ __forceinline__ void stage_b(uint32_t n, const uint32_t (&w)[5], uint32_t tile_base) {
    const bool act = (uint32_t)lane < n;
    const uint32_t id = act ? hl[(hh + lane) & (HL_CAP - 1)] : (uint32_t)lane << 6;
    const int src = id >> 6;
    const uint32_t q = (id & 63) * D;  // offset inside the owning lane's chunk
    const uint32_t c = q >> 4;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; i++) v[i] = __shfl_sync(FULL_MASK, w[i], src);
    uint32_t lo = v[0], hi = v[1];
    if (c == 1) { lo = v[1]; hi = v[2]; }
    if (c == 2) { lo = v[2]; hi = v[3]; }
    if (c == 3) { lo = v[3]; hi = v[4]; }
    const uint32_t x = __funnelshift_r(lo, hi, 2 * (q & 15)) & P.seed_mask;
    const uint32_t p = tile_base + src * CHUNK + q;
    bool found = false;
    uint32_t info = 0;
    if (act) {
      uint32_t slot = seed_slot(x, P.seedtab_shift);
      while (true) {
        const uint64_t e = __ldg(P.seedtab + slot);
        if (e == 0) break;
        if ((uint32_t)e == x) {
          found = true;
          info = (uint32_t)(e >> 32);
          break;
        }
        slot = (slot + 1) & P.seedtab_mask;
      }
    }
    hh += n;
    const uint32_t b = __ballot_sync(FULL_MASK, found);
    if (b) {
      if (found) cq[(ct + __popc(b & lt_mask)) & (CQ_CAP - 1)] = (uint64_t)info << 32 | p;
      ct += __popc(b);
      if (PROF && found) n_seed++;
      __syncwarp();
      if (ct - ch >= 32) stage_c(32);
    }
  }

  // ---- compact the set bits of one 32-lookup hit mask into the id ring -----------
  // bit 31 of acc is lookup idx0, bit 30 lookup idx0 + 1, ...
  __device__ __forceinline__ void push_hits(uint32_t acc, uint32_t idx0, const uint32_t (&w)[5],
                                            uint32_t tile_base) {
    if (PROF) n_bloom += __popc(acc);
    const uint32_t tag = (uint32_t)lane << 6 | idx0;
    while (true) {
      const bool has = acc != 0;
      const uint32_t b = __ballot_sync(FULL_MASK, has);
      if (b == 0) break;
      if (has) {
        const int i = __clz(acc);
        acc ^= 0x80000000u >> i;
        hl[(ht + __popc(b & lt_mask)) & (HL_CAP - 1)] = (uint16_t)(tag + i);
      }
      ht += __popc(b);
      __syncwarp();
      if (ht - hh >= 32) stage_b(32, w, tile_base);
    }
  }

  __device__ __forceinline__ void flush_tile(const uint32_t (&w)[5], uint32_t tile_base) {
    if (ht != hh) stage_b(ht - hh, w, tile_base);  // fewer than 32 left
  }

  __device__ __forceinline__ void drain() {
    while (ct != ch) stage_c(min(ct - ch, 32u));
  }

  // ---- stage A: the shared-memory seed filter over one lane chunk ---------------
  // w[0..3] hold this lane's 64 positions, w[4] the next 16 (halo for s-mers
  // that start in the chunk and end beyond it).
  __device__ __forceinline__ void stage_a(const uint32_t (&w)[5], uint32_t &acc0,
                                          uint32_t &acc1) const {
    acc0 = 0;
    acc1 = 0;
    const uint32_t mult = P.seed_mult;
    const uint32_t fbase = (uint32_t)__cvta_generic_to_shared(filt);
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
      for (int t = 0; t < 16; t += D) {
        const uint32_t x = t ? __funnelshift_r(w[c], w[c + 1], 2 * t) : w[c];
        const uint32_t h = x * mult;
        // byte address = idx * 4 + base as a multiply-add (P.four is opaque), not an ALU-pipe LEA
        const uint32_t addr = __umulhi(h, (uint32_t)BLOOM_WORDS) * P.four + fbase;
        uint32_t word;
        asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(addr));
        uint32_t bit = word << (x & 31);
        if (NH == 2) bit &= word << (__umulhi(h, SEED_MULT2) & 31);
        if ((c * 16 + t) / D < 32)
          acc0 = __funnelshift_l(bit, acc0, 1);
        else
          acc1 = __funnelshift_l(bit, acc1, 1);
      }
    }
    if constexpr (64 / D < 32) acc0 <<= (32 - 64 / D);
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

template <int D, int NH, bool PROF>
__global__ void __launch_bounds__(SCAN_THREADS, 1) k_scan(const ScanParams P) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t *filt = smem;
  uint64_t *cq_all = reinterpret_cast<uint64_t *>(smem + BLOOM_WORDS);
  uint16_t *hl_all = reinterpret_cast<uint16_t *>(cq_all + SCAN_WARPS * CQ_CAP);

  for (int i = threadIdx.x; i < BLOOM_WORDS / 4; i += SCAN_THREADS)
    reinterpret_cast<uint4 *>(filt)[i] = __ldg(reinterpret_cast<const uint4 *>(P.bloom) + i);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ScanWarp<D, NH, PROF> W(P, filt, hl_all + warp * HL_CAP, cq_all + warp * CQ_CAP, lane);

  const uint32_t n_warps = gridDim.x * SCAN_WARPS;
  uint32_t tile = blockIdx.x * SCAN_WARPS + warp;
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
    const uint32_t tile_base = tile * WTILE;
    W.push_hits(acc0, 0, w, tile_base);
    if (D == 1) W.push_hits(acc1, 32, w, tile_base);
    W.flush_tile(w, tile_base);
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
