// dkb_api.cu — the C ABI of include/dkb.h over the sm_100a kernels.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dkb.h"
#include "dkb_build.cuh"
#include "dkb_pack.cuh"

using namespace dkb;
static_assert(DKB_MAX_MULTI == dkb::MAX_SEGMENTS, "ABI and kernel disagree on the batches per launch");

struct dkb_ctx {
  int device = 0;
  int k = 0;
  int n_sms = 0;
  std::string err;

  dkb_tuning user_tuning{0, 0, 0, 0};
  int s = 0, D = 0, NH = 0;  // resolved at table build
  bool gf = false;           // seed filter probed in L2 instead of shared memory
  bool gate = false;         // stage A reads the flags and skips lookups of unusable seeds (modes 3, 4)
  uint32_t bloom_words = BLOOM_WORDS;

  // entries
  size_t n_entries = 0, n_live = 0;
  uint32_t n_variants = 0;
  uint64_t *d_keys = nullptr;
  uint32_t *d_variant = nullptr;
  uint8_t *d_allele = nullptr;
  uint8_t *d_dead = nullptr;
  // key table
  uint4 *d_tslots = nullptr;
  uint32_t table_slots = 0;
  // seeds
  uint4 *d_sslots = nullptr;  // seed table: 64-byte slots (seed + two records), 4 x uint4 each
  uint32_t *d_probe = nullptr;  // shared-memory filter mode: probe array in front of the slots
  uint32_t probe_slots = 0;
  bool canon = false;         // seeds keyed by their strand-canonical form
  uint32_t seed_slots = 0;
  uint32_t n_seeds = 0;
  uint32_t *d_bloom = nullptr, *d_pre = nullptr;
  uint32_t pre_words = 0;  // shared-memory pre-filter of the L2 filter mode
  bool want_pre = false;   // the tuner's choice
  uint64_t bloom_bits_set = 0;
  // counters and results
  uint32_t *d_counts = nullptr;  // [3][n_entries]
  uint32_t *d_hits = nullptr, *d_distinct = nullptr, *d_nkmers = nullptr;
  uint8_t *d_calls = nullptr;
  bool finalised = false;
  // profiling counters
  bool prof = false;
  unsigned long long *d_prof = nullptr;
  // streams / staging for host batches
  cudaStream_t s_scan = nullptr, s_copy = nullptr;
  // per-launch CUDA-event timing of the scan kernel (pooled event pairs)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool, ev_pending;
  double scan_ms_total = 0.0;
  uint64_t scan_launches_timed = 0;
  float last_scan_ms = 0.f;
  struct Stage {
    uint32_t *bases = nullptr, *mask = nullptr;
    size_t bases_cap = 0, mask_cap = 0;  // in words
    // raw reads for the device-side packer (dkb_batch_submit_reads)
    uint8_t *seq = nullptr, *qual = nullptr;
    uint64_t *offsets = nullptr, *nib = nullptr;
    size_t seq_cap = 0, qual_cap = 0, off_cap = 0, nib_cap = 0;  // bytes / elements
    // zero list of dkb_batch_submit_sparse
    uint32_t *zoff = nullptr;
    uint8_t *zbytes = nullptr;
    size_t zoff_cap = 0, zbytes_cap = 0;
    cudaEvent_t copied = nullptr, freed = nullptr;
    bool in_use = false;
  } stage[2];
  int next_stage = 0;
  uint64_t scan_launches = 0, positions_scanned = 0;
  // multi-GPU: the communicator and the overlapped count reduction (dkb_reduce_push)
  void *comm = nullptr;  // ncclComm_t
  int rank = 0, world = 1;
  cudaStream_t s_side = nullptr;
  uint32_t *d_red[2] = {nullptr, nullptr};  // snapshots of the counters being summed
  size_t red_cap = 0;                       // capacity of each, in uint32
  cudaEvent_t ev_snap[2] = {nullptr, nullptr}, ev_red[2] = {nullptr, nullptr};
  int red_next = 0, red_pending = -1, red_last = -1;
};

namespace {

thread_local std::string g_err;

// Dynamic shared-memory limit and carve-out last set for each (device, scan kernel).  The
// attributes belong to the function, not to a context, and the size an L2-filter kernel needs
// changes with the pre-filter every table build chooses - so they are set again whenever the
// size differs from what the function was last given, by any context of this process.
std::mutex g_smem_mu;
std::map<std::pair<int, const void *>, size_t> g_smem_set;

int fail(dkb_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  g_err = msg;
  return code;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? DKB_ENOMEM : DKB_ECUDA,          \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                    \
  } while (0)

uint32_t pow2_at_least(uint64_t x) {
  uint32_t p = 1024;
  while (p < x && p < 0x80000000u) p <<= 1;
  return p;
}

template <typename T>
void dfree(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

void free_table(dkb_ctx *c) {
  dfree(c->d_keys); dfree(c->d_variant); dfree(c->d_allele); dfree(c->d_dead);
  dfree(c->d_tslots);
  dfree(c->d_sslots); dfree(c->d_probe); dfree(c->d_bloom); dfree(c->d_pre);
  c->probe_slots = 0;
  dfree(c->d_counts); dfree(c->d_hits);  // (d_distinct, d_nkmers live in d_hits' block)
  c->d_distinct = c->d_nkmers = nullptr;
  dfree(c->d_calls);
  c->n_entries = c->n_live = 0;
  c->n_variants = 0;
  c->finalised = false;
  c->red_pending = c->red_last = -1;  // snapshots of the old table's counters are void
}

// Words of the L2-resident seed filter: 64 bits per seed, between the shared-memory size
// and 32 MB (it shares L2 with the tables and the stream; DKB_L2_FILTER_MAX_WORDS overrides).
// Measured on the 100 000-candidate WGS shard (6 M seeds): 4 / 8 / 16 / 32 / 48 MB -> 2.89 / 3.65 /
// 3.86 / 3.92 / 3.91 Tbases/s: below 16 MB the filter's false positives (each a DRAM access
// into the 1.5 GB slot table) cost more than the filter's own L2 misses.
uint32_t l2_filter_words(double seeds) {
  double cap = 8.0 * 1024 * 1024;  // 32 MB
  if (const char *e = getenv("DKB_L2_FILTER_MAX_WORDS")) cap = atof(e);
  double w = seeds * 2.0;
  if (w > cap) w = cap;
  if (w < BLOOM_WORDS) w = BLOOM_WORDS;
  return (uint32_t)w;
}
// Shared-memory pre-filter of the L2 filter mode: 145 KB - with the 18 KB of lists and the
// 1 KB the system reserves it fills the 164 KB shared-memory carve-out exactly and leaves
// 92 KB of L1, whose lines track the outstanding L2 loads.  Measured on configs[1]
// (Tbases/s; profiles/README.md): 107 KB (132 KB carve-out) 4.64, 128 KB 4.91, 139 KB 5.04,
// 171 KB (196 KB carve-out, 60 KB of L1) 4.94, 192 KB 2.8.
constexpr uint32_t PRE_WORDS = 37120;

// Words of the pre-filter the build should use for this table (0 = none).  DKB_PREFILTER_WORDS overrides.
uint32_t pre_filter_words(bool want) {
  if (const char *e = getenv("DKB_PREFILTER_WORDS")) {
    long w = atol(e);
    if (w < 0) w = 0;
    if (w > BLOOM_WORDS) w = BLOOM_WORDS;
    return (uint32_t)(w / 4 * 4);
  }
  return want ? PRE_WORDS : 0;
}

// Ladder seeds one SNV-sized haplotype strand needs (k windows) at stride D, seed length s.
double ladder_seeds_per_strand(int k, int s, int D) {
  const int E = k - s, G = ((E + 1) / D) * D, nwin = k;
  double n = 0;
  for (int r = 0; r < D; r++) {
    const int base = E - ((E - r) % D);  // first ladder element of residue r
    n += 1 + (nwin - 1 > base ? (nwin - 1 - base + G - 1) / G : 0);
  }
  return n;
}

// Modelled cost of one 2048-position warp tile, in cycles per scheduler: filter lookups +
// handling of the filter's false positives + stage C for s-mers of unrelated sequence that
// equal a seed by chance.  Constants fitted to measured scan times (profiles/README.md):
//  * a shared-memory lookup ~16 cycles (bank-conflict wavefronts as much as its instructions);
//  * an L2 lookup 119: an SM sustains about one random L2 load per clock, shared by its four
//    schedulers and 32 lanes - but only while ~100 KB of L1 are left to track the loads;
//  * behind the 139 KB pre-filter an L2 lookup costs 22 + 105 x (fraction the pre-filter passes);
//  * a false positive 9 (strides <= 4: verified by its own lane) or 16 (macro tiles: compacted,
//    bases re-read from L2); a chance seed match 25 (one record compare);
//  * a seed table beyond L2 turns every probe into a DRAM access (factor `table`).
double tile_cost(double n_entries, bool hints, int k, int s, int D, int NH, bool gf, bool pre,
                 double *seeds_out) {
  const double n_haps = n_entries / k;  // allele haplotypes (SNV-sized)
  // both strands; ref/alt haplotypes share the seeds that avoid the variant base (x0.7)
  double seeds = n_haps * 2.0 * 0.7 * ladder_seeds_per_strand(k, s, D);
  if (!hints) seeds *= 2.3;  // min-hash rule: ~2/(w+1) density instead of 1/w
  if (canon_for_mode(gf ? 1 : 0)) seeds *= 0.5;  // both strands share their (canonical) seeds
  if (seeds_out) *seeds_out = seeds;
  const double bits = 32.0 * (gf ? l2_filter_words(seeds) : (double)BLOOM_WORDS);
  const double dens = 1.0 - exp(-NH * seeds / bits);
  const double fp = pow(dens, NH) * 1.3 + 1e-4;  // 1.3: per-word load variance
  const double lookups_lane = 64.0 / D, lookups_tile = 2048.0 / D;
  const double hits = lookups_tile * fp;
  const double chance = seeds / pow(4.0, s);  // P(random s-mer is a seed)
  const double table = seeds > 2e6 ? 3.0 : seeds > 5e5 ? 1.3 : 1.0;
  const double chance_cost = 25.0 * table * lookups_tile * chance;
  // (strides 8, 16: lookups gated by the flag stream cost ~8 more and reach the filter 0.58 times
  // as often at the usual 4 % of unusable positions; they are used when that pays - see the build)
  double l2_lookup = D >= 8 ? 8.0 + 0.58 * 119.0 : 119.0;
  if (pre) {
    const double pass = 1.0 - exp(-seeds / (32.0 * pre_filter_words(true)));
    l2_lookup = D >= 8 && pass > 0.55 ? 30.0 + 0.58 * 105.0 * pass : 22.0 + 105.0 * pass;
  }
  if (D >= 8)  // (round-2 refit: the probe array and normal-priority stream loads made the shared-memory mode cheaper)
    return (gf ? 61.0 : 37.0) + lookups_lane * (gf ? l2_lookup : 16.0 + NH) + (gf ? 16.0 : 14.6) * table * hits +
           chance_cost;
  return 100.0 + lookups_lane * (gf ? l2_lookup : 13.0 + 1.5 * NH) + 9.0 * table * (hits < 64 ? hits : 64) +
         (hits > 64 ? 14.0 * table * (hits - 64) : 0.0) + chance_cost;
}

// Resolve (s, D, NH): the caller's choice, else DKB_TUNING="s,D,NH", else the
// cheapest combination under tile_cost().
int resolve_tuning(dkb_ctx *ctx, size_t n_entries, bool hints) {
  dkb_tuning t = ctx->user_tuning;
  if (const char *e = getenv("DKB_TUNING")) {
    int a = 0, b = 0, c = 0, d = 0;
    if (sscanf(e, "%d,%d,%d,%d", &a, &b, &c, &d) >= 1) {
      if (!t.seed_len) t.seed_len = a;
      if (!t.stride) t.stride = b;
      if (!t.bloom_hashes) t.bloom_hashes = c;
      if (!t.filter_mode) t.filter_mode = d;
    }
  }
  const int k = ctx->k;
  if (t.stride != 0 && t.stride != 1 && t.stride != 2 && t.stride != 4 && t.stride != 8 &&
      t.stride != 16)
    return fail(ctx, DKB_EINVAL, "stride must be 1, 2, 4, 8 or 16");
  if (t.bloom_hashes < 0 || t.bloom_hashes > 4)
    return fail(ctx, DKB_EINVAL, "bloom_hashes must be 1..4");
  if (t.seed_len != 0 && (t.seed_len < 8 || t.seed_len > MAX_SEED_LEN))
    return fail(ctx, DKB_EINVAL, "seed_len out of range (8..15 and <= k - stride + 1)");
  if (t.filter_mode < 0 || t.filter_mode > 2)
    return fail(ctx, DKB_EINVAL, "filter_mode must be 0 (auto), 1 (shared memory) or 2 (L2)");
  int best_D = 0, best_NH = 0, best_s = 0, best_gf = 0;
  bool best_pre = false;
  double best = 1e300;
  for (int mode = 0; mode <= 2; mode++) {  // shared memory, L2, L2 behind the pre-filter
    const int gf = mode > 0;
    const bool pre = mode == 2;
    if (pre && pre_filter_words(true) == 0) continue;
    if (mode == 1 && getenv("DKB_PREFILTER_WORDS") && pre_filter_words(true) != 0) continue;
    if (t.filter_mode && t.filter_mode != gf + 1) continue;
    for (int D = 1; D <= 16; D *= 2) {
      if (t.stride && t.stride != D) continue;
      if (gf && D < 2) continue;  // the L2 mode is built for strides 2..16
      for (int s = 8; s <= MAX_SEED_LEN; s++) {
        if (t.seed_len && t.seed_len != s) continue;
        if (s > k - D + 1) continue;
        for (int NH = 1; NH <= 4; NH++) {
          if (gf && NH > 2) continue;
          if (t.bloom_hashes ? t.bloom_hashes != NH : NH > 2) continue;  // 3, 4: manual only
          const double c = tile_cost((double)n_entries, hints, k, s, D, NH, gf != 0, pre, nullptr);
          if (c < best) { best = c; best_D = D; best_NH = NH; best_s = s; best_gf = gf; best_pre = pre; }
        }
      }
    }
  }
  if (best_D == 0)
    return fail(ctx, DKB_EINVAL,
                "no (seed_len, stride, hashes, filter_mode) fits: need seed_len <= k - stride + 1; "
                "L2 mode takes strides 2..16 and 1..2 hashes");
  ctx->s = best_s;
  ctx->D = best_D;
  ctx->NH = best_NH;
  ctx->gf = best_gf != 0;
  ctx->want_pre = best_pre;
  return DKB_OK;
}

// Fold finished scan launches into the running totals; recycle their events.
void collect_timing(dkb_ctx *ctx) {
  size_t keep = 0;
  for (auto &pr : ctx->ev_pending) {
    float ms = 0.f;
    if (cudaEventQuery(pr.second) == cudaSuccess &&
        cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
      ctx->scan_ms_total += ms;
      ctx->scan_launches_timed++;
      ctx->last_scan_ms = ms;
      ctx->ev_pool.push_back(pr);
    } else {
      cudaGetLastError();
      ctx->ev_pending[keep++] = pr;
    }
  }
  ctx->ev_pending.resize(keep);
}

SeedTable seed_table(const dkb_ctx *ctx) {
  SeedTable T;
  T.slots = ctx->d_sslots;
  T.n_slots = ctx->seed_slots;
  T.probe = ctx->d_probe;
  T.n_probe = ctx->probe_slots;
  return T;
}

KeyTable key_table(const dkb_ctx *ctx) {
  KeyTable T;
  T.slots = ctx->d_tslots;
  T.bucket_mask = ctx->table_slots / KBUCKET - 1;
  return T;
}

typedef void (*scan_fn)(const ScanParams);
}  // namespace

// The scan kernel's instantiations are compiled one stride per translation unit
// (dkb_scan_inst.cu with -DDKB_INST_D=...), in parallel; each exports its picker.
namespace dkb {
scan_fn pick_scan_d1(int NH, int fm, bool prof);
scan_fn pick_scan_d2(int NH, int fm, bool prof);
scan_fn pick_scan_d4(int NH, int fm, bool prof);
scan_fn pick_scan_d8(int NH, int fm, bool prof);
scan_fn pick_scan_d16(int NH, int fm, bool prof);
}  // namespace dkb

namespace {
scan_fn pick_scan(int D, int NH, int fm, bool prof) {
  switch (D) {
    case 1: return pick_scan_d1(NH, fm, prof);
    case 2: return pick_scan_d2(NH, fm, prof);
    case 4: return pick_scan_d4(NH, fm, prof);
    case 8: return pick_scan_d8(NH, fm, prof);
    case 16: return pick_scan_d16(NH, fm, prof);
  }
  return nullptr;
}

// One scan launch over n_seg device-resident packed streams (each into its sample's counters).
int launch_scan(dkb_ctx *ctx, int n_seg, const uint32_t *const *d_bases, const uint32_t *const *d_mask,
                const uint64_t *n_positions, const int *samples) {
  ScanParams P;
  memset(&P, 0, sizeof(P));
  // work units: tiles, or macro tiles of 4-8 tiles at strides 8/16
  const uint32_t per_unit = ctx->D >= 8 ? 32u / (64u / ctx->D) : 1u;
  uint64_t units = 0, total_pos = 0;
  P.n_seg = 0;
  for (int i = 0; i < n_seg; i++) {
    if (n_positions[i] == 0) continue;
    ScanSegment &S = P.seg[P.n_seg++];
    S.bases = d_bases[i];
    S.mask = d_mask[i];
    S.counts = ctx->d_counts + (size_t)samples[i] * ctx->n_entries;
    S.n_pos = (uint32_t)n_positions[i];
    S.n_bwords = (uint32_t)dkb_stream_bases_words(n_positions[i]);
    S.n_mwords = (uint32_t)dkb_stream_mask_words(n_positions[i]);
    S.n_tiles = (uint32_t)((n_positions[i] + WTILE - 1) / WTILE);
    S.unit_begin = (uint32_t)units;
    units += (S.n_tiles + per_unit - 1) / per_unit;
    S.unit_end = (uint32_t)units;
    total_pos += n_positions[i];
  }
  if (P.n_seg == 0) return DKB_OK;
  if (units >= 0xFFFF0000ull) return fail(ctx, DKB_EINVAL, "too many positions in one launch");
  for (int i = P.n_seg; i < MAX_SEGMENTS; i++) P.seg[i] = P.seg[P.n_seg - 1];  // never indexed; keep valid
  P.bloom = ctx->d_bloom;
  P.bloom_words = ctx->bloom_words;
  P.pre = ctx->d_pre;
  P.pre_words = ctx->gf ? ctx->pre_words : 0;
  P.st = seed_table(ctx);
  P.kt = key_table(ctx);
  P.seed_mult = SEED_MULT << (32 - 2 * ctx->s);
  P.seed_mask = (1u << (2 * ctx->s)) - 1;
  P.cshift = 32 - 2 * ctx->s;
  P.four = 4;
  for (int i = 0; i < 32; i++) P.pw[i] = 1u << i;
  P.filter_words = BLOOM_WORDS;
  P.k = ctx->k;
  P.s = ctx->s;
  P.prof = ctx->d_prof;
  scan_fn fn = pick_scan(ctx->D, ctx->NH, ctx->gf ? (ctx->pre_words ? 2 : 1) + (ctx->gate ? 2 : 0) : 0, ctx->prof);
  if (!fn) return fail(ctx, DKB_EINVAL, "no scan kernel for this tuning");
  size_t smem_bytes = !ctx->gf ? SCAN_SMEM_BYTES
                      : ctx->pre_words ? SCAN_SMEM_BYTES_PRE + (size_t)ctx->pre_words * 4 : SCAN_SMEM_BYTES_GF;
  if (ctx->D >= 8) smem_bytes += SCAN_TMA_BYTES;  // TMA builds: the per-warp stream ring of the macro path
  {
    std::lock_guard<std::mutex> lk(g_smem_mu);
    auto it = g_smem_set.find({ctx->device, (const void *)fn});
    if (it == g_smem_set.end() || it->second != smem_bytes) {
      CU(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)smem_bytes));
      // L2 filter mode: as much L1 as the block leaves - the smallest carve-out that holds it
      // (+ 1 KB the system reserves); the hint is a percentage of 228 KB that the driver rounds
      // UP to a supported size.  Shared-memory filter mode: everything.
      CU(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                              ctx->gf ? (int)(100 * (smem_bytes + 1024) / (228 * 1024))
                                      : (int)cudaSharedmemCarveoutMaxShared));
      g_smem_set[{ctx->device, (const void *)fn}] = smem_bytes;
    }
  }
  // one CTA per SM; short batches get one CTA per work unit so that they still spread over the SMs
  uint32_t grid = (uint32_t)units;
  if (grid > (uint32_t)ctx->n_sms) grid = ctx->n_sms;
  if (ctx->ev_pending.size() >= 256) collect_timing(ctx);
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  if (!ctx->ev_pool.empty()) {
    ev = ctx->ev_pool.back();
    ctx->ev_pool.pop_back();
  } else {
    CU(cudaEventCreate(&ev.first));
    CU(cudaEventCreate(&ev.second));
  }
  CU(cudaEventRecord(ev.first, ctx->s_scan));
  fn<<<grid, SCAN_THREADS, smem_bytes, ctx->s_scan>>>(P);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ev.second, ctx->s_scan));
  ctx->ev_pending.push_back(ev);
  ctx->scan_launches++;
  ctx->positions_scanned += total_pos;
  ctx->finalised = false;
  return DKB_OK;
}

int launch_scan(dkb_ctx *ctx, const uint32_t *d_bases, const uint32_t *d_mask, uint64_t n_positions,
                int sample) {
  return launch_scan(ctx, 1, &d_bases, &d_mask, &n_positions, &sample);
}

// Grow a staging buffer (elements of T); the caller has made sure no kernel still reads it.
template <typename T>
int grow(dkb_ctx *ctx, T *&p, size_t &cap, size_t need) {
  if (cap >= need) return DKB_OK;
  dfree(p);
  cap = 0;  // a failed allocation must not leave a capacity behind a null pointer
  const size_t want = need + need / 8 + 256;
  T *q = nullptr;
  CU(cudaMalloc(&q, want * sizeof(T)));
  p = q;
  cap = want;
  return DKB_OK;
}

// ---- host memory next to the GPU ---------------------------------------------------------
// NUMA node of the device's PCIe slot (-1 if the system does not say) and that node's CPUs.
int gpu_numa_node(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  for (char *c = bus; *c; c++) *c = (char)tolower(*c);
  std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
  int node = -1;
  if (!(f >> node)) return -1;
  return node;
}

bool node_cpus(int node, cpu_set_t *set) {
  std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  std::string list;
  if (!std::getline(f, list)) return false;
  CPU_ZERO(set);
  int n = 0;
  size_t i = 0;
  while (i < list.size()) {  // "0-31,64-95"
    char *end = nullptr;
    long a = strtol(list.c_str() + i, &end, 10), b = a;
    if (end == list.c_str() + i) break;
    i = end - list.c_str();
    if (i < list.size() && list[i] == '-') {
      b = strtol(list.c_str() + i + 1, &end, 10);
      i = end - list.c_str();
    }
    for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, set); n++; }
    if (i < list.size() && list[i] == ',') i++;
  }
  return n > 0;
}

// Prefer `node` for this thread's page allocations (-1: back to the default policy).  Best
// effort: containers often forbid the call, and then CPU affinity + first touch decide.
void prefer_node(int node) {
#ifdef SYS_set_mempolicy
  if (node < 0) {
    syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
    return;
  }
  unsigned long mask[16] = {0};
  if (node >= (int)(sizeof(mask) * 8)) return;
  mask[node / (8 * sizeof(long))] |= 1ul << (node % (8 * sizeof(long)));
  syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof(mask) * 8);
#else
  (void)node;
#endif
}

// ---- NCCL, bound at run time ---------------------------------------------------------
// libnccl.so.2 is opened with dlopen the first time a communicator is asked for, so a
// single-GPU user needs no NCCL at all; in a process that already holds NCCL (PyTorch's
// bundled copy) the same library instance is used.  Only the calls below are needed; their
// C signatures are stable across NCCL 2.x.
struct NcclId { char internal[128]; };  // ncclUniqueId
struct Nccl {
  void *h = nullptr;
  int (*GetUniqueId)(NcclId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*CommCount)(void *, int *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
  std::string err;
};
constexpr int NCCL_UINT32 = 3, NCCL_SUM = 0;  // ncclUint32, ncclSum

Nccl *nccl() {
  static Nccl N;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {getenv("DKB_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      if ((N.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
      N.err = dlerror();
    }
    if (!N.h) return;
    auto sym = [&](const char *name) {
      void *p = dlsym(N.h, name);
      if (!p) N.err = std::string("missing NCCL symbol ") + name;
      return p;
    };
    N.GetUniqueId = (decltype(N.GetUniqueId))sym("ncclGetUniqueId");
    N.CommInitRank = (decltype(N.CommInitRank))sym("ncclCommInitRank");
    N.CommDestroy = (decltype(N.CommDestroy))sym("ncclCommDestroy");
    N.CommCount = (decltype(N.CommCount))sym("ncclCommCount");
    N.AllReduce = (decltype(N.AllReduce))sym("ncclAllReduce");
    N.GetErrorString = (decltype(N.GetErrorString))sym("ncclGetErrorString");
    N.GetVersion = (decltype(N.GetVersion))sym("ncclGetVersion");
    if (!N.GetUniqueId || !N.CommInitRank || !N.CommDestroy || !N.AllReduce || !N.GetErrorString) {
      dlclose(N.h);
      N.h = nullptr;
    }
  });
  return &N;
}

#define NC(call)                                                                              \
  do {                                                                                        \
    int r_ = (call);                                                                          \
    if (r_ != 0)                                                                              \
      return fail(ctx, DKB_ENCCL, std::string(#call) + ": " + nccl()->GetErrorString(r_));    \
  } while (0)

// Side stream, events and the two snapshot buffers of the overlapped reduction.
int reduce_setup(dkb_ctx *ctx) {
  const size_t need = (ctx->n_entries ? ctx->n_entries : 1) * 3;
  if (!ctx->s_side) {
    CU(cudaStreamCreateWithFlags(&ctx->s_side, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CU(cudaEventCreateWithFlags(&ctx->ev_snap[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ctx->ev_red[i], cudaEventDisableTiming));
    }
  }
  if (ctx->red_cap < need) {
    CU(cudaStreamSynchronize(ctx->s_side));
    CU(cudaStreamSynchronize(ctx->s_scan));
    for (int i = 0; i < 2; i++) dfree(ctx->d_red[i]);
    ctx->red_cap = 0;
    for (int i = 0; i < 2; i++) CU(cudaMalloc(&ctx->d_red[i], need * 4));
    ctx->red_cap = need;
    ctx->red_pending = ctx->red_last = -1;
  }
  return DKB_OK;
}

int finalise_counts(dkb_ctx *ctx, const dkb_thresholds *thr, const uint32_t *d_counts);

// Kernel 3 on snapshot b, on the scan stream, once its sum over ranks is complete.
int finalise_snapshot(dkb_ctx *ctx, const dkb_thresholds *thr, int b) {
  CU(cudaStreamWaitEvent(ctx->s_scan, ctx->ev_red[b], 0));
  int rc = finalise_counts(ctx, thr, ctx->d_red[b]);
  ctx->red_last = b;
  return rc;
}

}  // namespace

extern "C" {

int dkb_abi_version(void) { return DKB_ABI_VERSION; }

const char *dkb_strerror(int code) {
  switch (code) {
    case DKB_OK: return "ok";
    case DKB_EINVAL: return "invalid argument";
    case DKB_ECUDA: return "CUDA error";
    case DKB_ENOMEM: return "out of memory";
    case DKB_ESTATE: return "call out of order";
    case DKB_ENODEV: return "no usable CUDA device (no CPU fallback exists)";
    case DKB_ENCCL: return "NCCL error";
    default: return "unknown error";
  }
}

const char *dkb_last_error(const dkb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int dkb_ctx_create(int device, int k, dkb_ctx **out) {
  dkb_ctx *ctx = nullptr;
  if (!out) return fail(nullptr, DKB_EINVAL, "out is null");
  *out = nullptr;
  if (k < DKB_MIN_K || k > DKB_MAX_K) return fail(nullptr, DKB_EINVAL, "k out of range [8, 31]");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(nullptr, DKB_ENODEV, "no CUDA device; this library has no CPU fallback");
  }
  if (device < 0 || device >= n_dev) return fail(nullptr, DKB_EINVAL, "device index out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return fail(nullptr, DKB_ECUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, DKB_ENODEV,
                std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                    "; kernels are built for sm_100a only");
  ctx = new (std::nothrow) dkb_ctx();
  if (!ctx) return fail(nullptr, DKB_ENOMEM, "host allocation failed");
  ctx->device = device;
  ctx->k = k;
  ctx->n_sms = prop.multiProcessorCount;
  int rc = [&]() -> int {
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&ctx->s_scan, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
    for (auto &st : ctx->stage) {
      CU(cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&st.freed, cudaEventDisableTiming));
    }
    CU(cudaMalloc(&ctx->d_prof, 4 * sizeof(unsigned long long)));
    CU(cudaMemset(ctx->d_prof, 0, 4 * sizeof(unsigned long long)));
    return DKB_OK;
  }();
  if (rc != DKB_OK) {
    g_err = ctx->err;
    dkb_ctx_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return DKB_OK;
}

int dkb_ctx_destroy(dkb_ctx *ctx) {
  if (!ctx) return DKB_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  free_table(ctx);
  dfree(ctx->d_prof);
  for (auto &st : ctx->stage) {
    dfree(st.bases);
    dfree(st.mask);
    dfree(st.seq);
    dfree(st.qual);
    dfree(st.offsets);
    dfree(st.nib);
    dfree(st.zoff);
    dfree(st.zbytes);
    if (st.copied) cudaEventDestroy(st.copied);
    if (st.freed) cudaEventDestroy(st.freed);
  }
  for (auto *v : {&ctx->ev_pool, &ctx->ev_pending})
    for (auto &pr : *v) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  if (ctx->comm && nccl()->h) nccl()->CommDestroy(ctx->comm);
  for (int i = 0; i < 2; i++) {
    dfree(ctx->d_red[i]);
    if (ctx->ev_snap[i]) cudaEventDestroy(ctx->ev_snap[i]);
    if (ctx->ev_red[i]) cudaEventDestroy(ctx->ev_red[i]);
  }
  if (ctx->s_side) cudaStreamDestroy(ctx->s_side);
  if (ctx->s_scan) cudaStreamDestroy(ctx->s_scan);
  if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
  delete ctx;
  return DKB_OK;
}

int dkb_ctx_set_tuning(dkb_ctx *ctx, const dkb_tuning *tuning) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  ctx->user_tuning = tuning ? *tuning : dkb_tuning{0, 0, 0, 0};
  return DKB_OK;
}

int dkb_ctx_get_tuning(const dkb_ctx *ctx, dkb_tuning *out) {
  if (!ctx || !out) return fail(nullptr, DKB_EINVAL, "null argument");
  out->seed_len = ctx->s;
  out->stride = ctx->D;
  out->bloom_hashes = ctx->NH;
  out->filter_mode = ctx->gf ? 2 : 1;
  return DKB_OK;
}

int dkb_table_build(dkb_ctx *ctx, const uint64_t *keys, const uint32_t *variant_ids,
                    const uint8_t *allele_ids, const uint16_t *win_index,
                    const uint16_t *win_count, size_t n, uint32_t n_variants) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (n && (!keys || !variant_ids || !allele_ids)) return fail(ctx, DKB_EINVAL, "null entry array");
  if (n >= 0x40000000ull) return fail(ctx, DKB_EINVAL, "too many entries (max 2^30 - 1)");
  if ((win_index == nullptr) != (win_count == nullptr))
    return fail(ctx, DKB_EINVAL, "win_index and win_count must be given together");
  const uint64_t km = kmer_mask(ctx->k);
  for (size_t i = 0; i < n; i++) {
    if (keys[i] > km) return fail(ctx, DKB_EINVAL, "key wider than 2k bits");
    if (variant_ids[i] >= n_variants) return fail(ctx, DKB_EINVAL, "variant id out of range");
    if (allele_ids[i] >= DKB_N_ALLELES) return fail(ctx, DKB_EINVAL, "allele id must be 0 or 1");
    if (win_index && ((win_index[i] & 0x7FFFu) >= win_count[i]))
      return fail(ctx, DKB_EINVAL, "win_index >= win_count");
  }
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->s_scan));
  CU(cudaStreamSynchronize(ctx->s_copy));
  free_table(ctx);
  int rc = resolve_tuning(ctx, n, win_index != nullptr);
  if (rc != DKB_OK) return rc;

  ctx->n_entries = n;
  ctx->n_variants = n_variants;
  const size_t n1 = n ? n : 1, nv1 = n_variants ? n_variants : 1;
  int slots_per_entry = 8;
  if (const char *e = getenv("DKB_KEY_SLOTS_PER_ENTRY")) slots_per_entry = atoi(e) > 0 ? atoi(e) : 8;
  ctx->table_slots = pow2_at_least((uint64_t)slots_per_entry * n + 2);  // <= 0.25 entries per 2-slot bucket
  uint16_t *d_wi = nullptr, *d_wc = nullptr;
  uint32_t *d_slot_of = nullptr;
  uint32_t *d_set = nullptr, *d_cov = nullptr;
  unsigned int *d_nseeds = nullptr;
  unsigned long long *d_stats = nullptr;
  cudaStream_t st = ctx->s_scan;
  auto cleanup = [&]() {
    dfree(d_wi); dfree(d_wc); dfree(d_slot_of); dfree(d_set); dfree(d_cov); dfree(d_nseeds); dfree(d_stats);
  };
  rc = [&]() -> int {
    CU(cudaMalloc(&ctx->d_keys, n1 * 8));
    CU(cudaMalloc(&ctx->d_variant, n1 * 4));
    CU(cudaMalloc(&ctx->d_allele, n1));
    CU(cudaMalloc(&ctx->d_dead, n1));
    CU(cudaMalloc(&d_slot_of, n1 * 4));
    CU(cudaMalloc(&ctx->d_tslots, (size_t)ctx->table_slots * 16));
    CU(cudaMalloc(&ctx->d_counts, n1 * 3 * 4));
    CU(cudaMalloc(&ctx->d_hits, nv1 * 14 * 4));  // hits, distinct, n_kmers: one block, one memset per finalise
    ctx->d_distinct = ctx->d_hits + nv1 * 6;
    ctx->d_nkmers = ctx->d_hits + nv1 * 12;
    CU(cudaMalloc(&ctx->d_calls, nv1));
    CU(cudaMalloc(&d_nseeds, 4));
    CU(cudaMemsetAsync(ctx->d_tslots, 0xFF, (size_t)ctx->table_slots * 16, st));
    CU(cudaMemsetAsync(ctx->d_counts, 0, n1 * 3 * 4, st));
    CU(cudaMemsetAsync(ctx->d_dead, 0, n1, st));
    CU(cudaMemsetAsync(d_nseeds, 0, 4, st));
    if (n) {
      CU(cudaMemcpyAsync(ctx->d_keys, keys, n * 8, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(ctx->d_variant, variant_ids, n * 4, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(ctx->d_allele, allele_ids, n, cudaMemcpyHostToDevice, st));
      if (win_index) {
        CU(cudaMalloc(&d_wi, n * 2));
        CU(cudaMalloc(&d_wc, n * 2));
        CU(cudaMemcpyAsync(d_wi, win_index, n * 2, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_wc, win_count, n * 2, cudaMemcpyHostToDevice, st));
      }
    }
    // sizing pass: count the distinct seeds in a plain open-addressing set
    const uint32_t set_slots = pow2_at_least(4 * (uint64_t)n * ctx->D + 2);
    CU(cudaMalloc(&d_set, (size_t)set_slots * 4));
    CU(cudaMemsetAsync(d_set, 0xFF, (size_t)set_slots * 4, st));
    BuildParams B;
    B.keys = ctx->d_keys; B.variant = ctx->d_variant; B.allele = ctx->d_allele;
    B.win_index = d_wi; B.win_count = d_wc; B.n = (uint32_t)n;
    B.kt = key_table(ctx); B.slot_of = d_slot_of; B.dead = ctx->d_dead;
    B.k = ctx->k; B.s = ctx->s; B.D = ctx->D;
    // (whether the L2 mode runs behind a pre-filter does not matter here: both are mode > 0)
    ctx->canon = canon_for_mode(ctx->gf ? 1 : 0);
    B.canon = ctx->canon;
    const uint32_t seed_mult = SEED_MULT << (32 - 2 * ctx->s);
    const int TB = 256;
    const uint32_t g1 = (uint32_t)((n + TB - 1) / TB), g2 = (uint32_t)((2 * n + TB - 1) / TB);
    if (n) {
      k_insert_entries<<<g1, TB, 0, st>>>(B);
      k_mark_repeats<<<g1, TB, 0, st>>>(B);
      k_apply_dead<<<g1, TB, 0, st>>>(B);
      k_assign_seeds<<<g2, TB, 0, st>>>(B, ASSIGN_COUNT, d_set, set_slots - 1, d_nseeds, SeedTable{},
                                        nullptr, nullptr, 0, nullptr, 0, seed_mult, ctx->NH);
      CU(cudaGetLastError());
    }
    unsigned int n_seeds = 0;
    CU(cudaMemcpyAsync(&n_seeds, d_nseeds, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->n_seeds = n_seeds;
    // Seed table: 64-byte slots (seed + a 32-byte record per read orientation), a quarter full, any
    // size.  Only occupied slots are ever touched by true hits, so the hot set is 32 B per seed
    // and orientation; an absent seed (a filter false positive) lands on a slot marked
    // ST_MOVED_BIT - a second load - about one time in 8.
    double per_seed = 4.0;  // (2.0: 3 % slower on configs[1] - displaced seeds cost a dependent load per step)
    if (const char *e = getenv("DKB_SEED_SLOTS_PER_SEED")) per_seed = atof(e) >= 1.1 ? atof(e) : per_seed;
    const double want_slots = per_seed * n_seeds + 64;
    if (want_slots >= 4294967295.0) return fail(ctx, DKB_EINVAL, "too many seeds for the seed table");
    ctx->seed_slots = (uint32_t)want_slots;
    CU(cudaMalloc(&ctx->d_sslots, (size_t)ctx->seed_slots * 64));
    if (!ctx->gf && !getenv("DKB_NO_PROBE_ARRAY")) {  // probe array: 8 words per seed (dkb_device.cuh)
      const double want_probe = 8.0 * n_seeds + 64;
      if (want_probe >= 4294967295.0) return fail(ctx, DKB_EINVAL, "too many seeds for the probe array");
      ctx->probe_slots = (uint32_t)want_probe;
      CU(cudaMalloc(&ctx->d_probe, (size_t)ctx->probe_slots * 4));
      CU(cudaMemsetAsync(ctx->d_probe, 0x40, (size_t)ctx->probe_slots * 4, st));  // ST_EMPTY
    }
    CU(cudaMalloc(&d_cov, (size_t)ctx->seed_slots * 24));
    CU(cudaMemsetAsync(d_cov, 0, (size_t)ctx->seed_slots * 24, st));
    ctx->bloom_words = ctx->gf ? l2_filter_words((double)n_seeds) / 4 * 4 : (uint32_t)BLOOM_WORDS;
    CU(cudaMalloc(&ctx->d_bloom, (size_t)ctx->bloom_words * 4));
    CU(cudaMemsetAsync(ctx->d_bloom, 0, (size_t)ctx->bloom_words * 4, st));
    ctx->pre_words = ctx->gf ? pre_filter_words(ctx->want_pre) : 0;
    // Gated lookups (macro path of the L2 modes): worth their instructions when most lookups
    // end in an L2 gather - always without a pre-filter, behind one when it passes more than
    // about half (measured, profiles/README.md: 10 000 candidates, 40 % passed: 7.15 -> 6.45
    // Tbases/s; 14 000, 51 %: 5.78 -> 5.83; 20 000, 64 %: 4.52 -> 5.02; no pre-filter: 3.91 -> 6.23).
    ctx->gate = false;
    if (ctx->gf && ctx->D >= 8) {
      const double pass = ctx->pre_words ? 1.0 - exp(-(double)n_seeds / (32.0 * ctx->pre_words)) : 1.0;
      ctx->gate = pass > 0.55;
      if (const char *e = getenv("DKB_GATE")) ctx->gate = atoi(e) != 0;
    }
    if (ctx->pre_words) {
      CU(cudaMalloc(&ctx->d_pre, (size_t)ctx->pre_words * 4));
      CU(cudaMemsetAsync(ctx->d_pre, 0, (size_t)ctx->pre_words * 4, st));
    }
    const SeedTable T = seed_table(ctx);
    k_init_slots<<<(uint32_t)((4 * (size_t)ctx->seed_slots + TB - 1) / TB), TB, 0, st>>>(T);
    if (n) {
      k_assign_seeds<<<g2, TB, 0, st>>>(B, ASSIGN_INSERT, nullptr, 0, nullptr, T, nullptr,
                                        ctx->d_bloom, ctx->bloom_words, ctx->d_pre, ctx->pre_words,
                                        seed_mult, ctx->NH);
      k_assign_seeds<<<g2, TB, 0, st>>>(B, ASSIGN_RECORD, nullptr, 0, nullptr, T, d_cov,
                                        ctx->d_bloom, ctx->bloom_words, ctx->d_pre, ctx->pre_words,
                                        seed_mult, ctx->NH);
      k_finish_records<<<(uint32_t)((2 * (size_t)ctx->seed_slots + TB - 1) / TB), TB, 0, st>>>(T, d_cov);
    }
    // build statistics, counted on the device (the filter can be 16 MB)
    CU(cudaMalloc(&d_stats, 16));
    CU(cudaMemsetAsync(d_stats, 0, 16, st));
    k_build_stats<<<296, 256, 0, st>>>(ctx->d_bloom, ctx->bloom_words, ctx->d_dead, (uint32_t)n, d_stats);
    CU(cudaGetLastError());
    unsigned long long h_stats[2] = {0, 0};
    CU(cudaMemcpyAsync(h_stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->bloom_bits_set = h_stats[0];
    ctx->n_live = (size_t)h_stats[1];
    // Optional (DKB_L2_PERSIST_MB=n): set n MB of L2 aside for persisting lines and mark the
    // L2 filter as such for everything launched on the scan stream.
    if (const char *e = getenv("DKB_L2_PERSIST_MB")) {
      const size_t want = (size_t)atol(e) << 20;
      cudaStreamAttrValue av;
      memset(&av, 0, sizeof(av));
      if (want && ctx->gf) {
        CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        size_t bytes = (size_t)ctx->bloom_words * 4;
        int max_win = 0;
        CU(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device));
        if (bytes > (size_t)max_win) bytes = (size_t)max_win;
        av.accessPolicyWindow.base_ptr = ctx->d_bloom;
        av.accessPolicyWindow.num_bytes = bytes;
        av.accessPolicyWindow.hitRatio = bytes <= want ? 1.0f : (float)want / (float)bytes;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      }
      CU(cudaStreamSetAttribute(ctx->s_scan, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    return DKB_OK;
  }();
  cleanup();
  if (rc != DKB_OK) free_table(ctx);
  return rc;
}

int dkb_batch_submit_device(dkb_ctx *ctx, const uint32_t *d_bases2, const uint32_t *d_mask1,
                            uint64_t n_positions, int sample) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->d_tslots) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  if (sample < 0 || sample >= DKB_N_SAMPLES) return fail(ctx, DKB_EINVAL, "sample must be 0, 1 or 2");
  if (n_positions == 0) return DKB_OK;
  if (n_positions > 0xFFFFF000ull) return fail(ctx, DKB_EINVAL, "batch too long (max 2^32 - 4096 positions)");
  if (!d_bases2 || !d_mask1) return fail(ctx, DKB_EINVAL, "null stream pointer");
  if ((uintptr_t)d_bases2 & 15) return fail(ctx, DKB_EINVAL, "d_bases2 must be 16-byte aligned");
  if ((uintptr_t)d_mask1 & 3) return fail(ctx, DKB_EINVAL, "d_mask1 must be 4-byte aligned");
  CU(cudaSetDevice(ctx->device));
  return launch_scan(ctx, d_bases2, d_mask1, n_positions, sample);
}

int dkb_batch_submit_device_multi(dkb_ctx *ctx, int n_batches, const uint32_t *const *d_bases2,
                                  const uint32_t *const *d_mask1, const uint64_t *n_positions,
                                  const int *samples) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->d_tslots) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  if (n_batches < 0 || n_batches > DKB_MAX_MULTI) return fail(ctx, DKB_EINVAL, "n_batches out of range");
  if (n_batches && (!d_bases2 || !d_mask1 || !n_positions || !samples)) return fail(ctx, DKB_EINVAL, "null argument");
  for (int i = 0; i < n_batches; i++) {
    if (samples[i] < 0 || samples[i] >= DKB_N_SAMPLES) return fail(ctx, DKB_EINVAL, "sample must be 0, 1 or 2");
    if (n_positions[i] == 0) continue;
    if (n_positions[i] > 0xFFFFF000ull) return fail(ctx, DKB_EINVAL, "batch too long (max 2^32 - 4096 positions)");
    if (!d_bases2[i] || !d_mask1[i]) return fail(ctx, DKB_EINVAL, "null stream pointer");
    if ((uintptr_t)d_bases2[i] & 15) return fail(ctx, DKB_EINVAL, "d_bases2 must be 16-byte aligned");
    if ((uintptr_t)d_mask1[i] & 3) return fail(ctx, DKB_EINVAL, "d_mask1 must be 4-byte aligned");
  }
  CU(cudaSetDevice(ctx->device));
  return launch_scan(ctx, n_batches, d_bases2, d_mask1, n_positions, samples);
}

int dkb_batch_submit(dkb_ctx *ctx, const uint32_t *bases2, const uint32_t *mask1,
                     uint64_t n_positions, int sample) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->d_tslots) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  if (sample < 0 || sample >= DKB_N_SAMPLES) return fail(ctx, DKB_EINVAL, "sample must be 0, 1 or 2");
  if (n_positions == 0) return DKB_OK;
  if (n_positions > 0xFFFFF000ull) return fail(ctx, DKB_EINVAL, "batch too long (max 2^32 - 4096 positions)");
  if (!bases2 || !mask1) return fail(ctx, DKB_EINVAL, "null stream pointer");
  CU(cudaSetDevice(ctx->device));
  const size_t bw = dkb_stream_bases_words(n_positions), mw = dkb_stream_mask_words(n_positions);
  dkb_ctx::Stage &st = ctx->stage[ctx->next_stage];
  ctx->next_stage ^= 1;
  if (st.in_use) CU(cudaStreamWaitEvent(ctx->s_copy, st.freed, 0));
  if (st.bases_cap < bw || st.mask_cap < mw) {
    // the previous scan may still read the old buffers
    if (st.in_use) CU(cudaEventSynchronize(st.freed));
    int rc;
    if ((rc = grow(ctx, st.bases, st.bases_cap, bw)) != DKB_OK) return rc;
    if ((rc = grow(ctx, st.mask, st.mask_cap, mw)) != DKB_OK) return rc;
  }
  CU(cudaMemcpyAsync(st.bases, bases2, bw * 4, cudaMemcpyHostToDevice, ctx->s_copy));
  CU(cudaMemcpyAsync(st.mask, mask1, mw * 4, cudaMemcpyHostToDevice, ctx->s_copy));
  CU(cudaEventRecord(st.copied, ctx->s_copy));
  CU(cudaStreamWaitEvent(ctx->s_scan, st.copied, 0));
  int rc = launch_scan(ctx, st.bases, st.mask, n_positions, sample);
  if (rc != DKB_OK) return rc;
  CU(cudaEventRecord(st.freed, ctx->s_scan));
  st.in_use = true;
  return DKB_OK;
}

int dkb_batch_submit_sparse(dkb_ctx *ctx, const uint32_t *bases2, const uint32_t *zoff,
                            const uint8_t *zbytes, size_t zbytes_used, uint64_t n_positions, int sample) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->d_tslots) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  if (sample < 0 || sample >= DKB_N_SAMPLES) return fail(ctx, DKB_EINVAL, "sample must be 0, 1 or 2");
  if (n_positions == 0) return DKB_OK;
  if (n_positions > 0xFFFFF000ull) return fail(ctx, DKB_EINVAL, "batch too long (max 2^32 - 4096 positions)");
  if (!bases2 || !zoff || (zbytes_used && !zbytes)) return fail(ctx, DKB_EINVAL, "null stream pointer");
  const size_t nb = dkb_zero_list_blocks(n_positions);
  if ((zoff[nb] & 0x7FFFFFFFu) != zbytes_used) return fail(ctx, DKB_EINVAL, "zoff[n_blocks] != zbytes_used");
  // the expander trusts the offsets: they must ascend inside zbytes, a plain-bits block is 256 bytes
  for (size_t b = 0; b < nb; b++) {
    const uint32_t cur = zoff[b] & 0x7FFFFFFFu, nxt = zoff[b + 1] & 0x7FFFFFFFu;
    if (nxt < cur || nxt > zbytes_used || ((zoff[b] & 0x80000000u) && nxt - cur != 256u))
      return fail(ctx, DKB_EINVAL, "malformed zero list (block offsets)");
  }
  CU(cudaSetDevice(ctx->device));
  const size_t bw = dkb_stream_bases_words(n_positions), mw = dkb_stream_mask_words(n_positions);
  dkb_ctx::Stage &st = ctx->stage[ctx->next_stage];
  ctx->next_stage ^= 1;
  const size_t zb1 = zbytes_used ? zbytes_used : 1;
  if (st.in_use) CU(cudaStreamWaitEvent(ctx->s_copy, st.freed, 0));
  int rc;
  if (st.bases_cap < bw || st.mask_cap < mw || st.zoff_cap < nb + 1 || st.zbytes_cap < zb1) {
    if (st.in_use) CU(cudaEventSynchronize(st.freed));  // the previous scan may still read the old buffers
    if ((rc = grow(ctx, st.bases, st.bases_cap, bw)) != DKB_OK) return rc;
    if ((rc = grow(ctx, st.mask, st.mask_cap, mw)) != DKB_OK) return rc;
    if ((rc = grow(ctx, st.zoff, st.zoff_cap, nb + 1)) != DKB_OK) return rc;
    if ((rc = grow(ctx, st.zbytes, st.zbytes_cap, zb1)) != DKB_OK) return rc;
  }
  CU(cudaMemcpyAsync(st.bases, bases2, bw * 4, cudaMemcpyHostToDevice, ctx->s_copy));
  CU(cudaMemcpyAsync(st.zoff, zoff, (nb + 1) * 4, cudaMemcpyHostToDevice, ctx->s_copy));
  if (zbytes_used) CU(cudaMemcpyAsync(st.zbytes, zbytes, zbytes_used, cudaMemcpyHostToDevice, ctx->s_copy));
  CU(cudaEventRecord(st.copied, ctx->s_copy));
  CU(cudaStreamWaitEvent(ctx->s_scan, st.copied, 0));
  k_expand_zero_list<<<(uint32_t)((nb + 7) / 8), 256, 0, ctx->s_scan>>>(st.zoff, st.zbytes, (uint32_t)nb,
                                                                      (uint32_t)mw, st.mask);
  CU(cudaGetLastError());
  rc = launch_scan(ctx, st.bases, st.mask, n_positions, sample);
  if (rc != DKB_OK) return rc;
  CU(cudaEventRecord(st.freed, ctx->s_scan));
  st.in_use = true;
  return DKB_OK;
}

int dkb_batch_submit_reads(dkb_ctx *ctx, const uint8_t *seq, int seq_format, const uint8_t *qual,
                           const uint64_t *offsets, size_t n_reads, int min_baseq, int sample) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->d_tslots) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  if (sample < 0 || sample >= DKB_N_SAMPLES) return fail(ctx, DKB_EINVAL, "sample must be 0, 1 or 2");
  if (seq_format != 0 && seq_format != 1) return fail(ctx, DKB_EINVAL, "seq_format must be 0 (ASCII) or 1 (BAM 4-bit)");
  if (n_reads == 0) return DKB_OK;
  if (!seq || !offsets) return fail(ctx, DKB_EINVAL, "null read buffer");
  if (n_reads >= 0xFFFFFFFFull) return fail(ctx, DKB_EINVAL, "too many reads in one batch");
  for (size_t r = 0; r < n_reads; r++)
    if (offsets[r + 1] < offsets[r]) return fail(ctx, DKB_EINVAL, "offsets must not decrease");
  const uint64_t n_pos = dkb_stream_positions(offsets, n_reads);
  if (n_pos > 0xFFFFF000ull) return fail(ctx, DKB_EINVAL, "batch too long (max 2^32 - 4096 positions)");
  const uint64_t n_bases = offsets[n_reads] - offsets[0];
  CU(cudaSetDevice(ctx->device));
  const size_t bw = dkb_stream_bases_words(n_pos), mw = dkb_stream_mask_words(n_pos);
  // BAM keeps every read's 4-bit codes byte-aligned: (len + 1) / 2 bytes per read
  std::vector<uint64_t> nib;
  uint64_t seq_bytes = n_bases;
  if (seq_format == 1) {
    nib.resize(n_reads);
    uint64_t b = 0;
    for (size_t r = 0; r < n_reads; r++) {
      nib[r] = b;
      b += (offsets[r + 1] - offsets[r] + 1) / 2;
    }
    seq_bytes = b;
  }
  dkb_ctx::Stage &st = ctx->stage[ctx->next_stage];
  ctx->next_stage ^= 1;
  if (st.in_use) CU(cudaEventSynchronize(st.freed));  // buffers may be regrown below
  int rc;
  if ((rc = grow(ctx, st.bases, st.bases_cap, bw)) != DKB_OK) return rc;
  if ((rc = grow(ctx, st.mask, st.mask_cap, mw)) != DKB_OK) return rc;
  if ((rc = grow(ctx, st.seq, st.seq_cap, (size_t)seq_bytes)) != DKB_OK) return rc;
  if (qual && (rc = grow(ctx, st.qual, st.qual_cap, (size_t)n_bases)) != DKB_OK) return rc;
  if ((rc = grow(ctx, st.offsets, st.off_cap, n_reads + 1)) != DKB_OK) return rc;
  if (seq_format == 1 && (rc = grow(ctx, st.nib, st.nib_cap, n_reads)) != DKB_OK) return rc;
  // the reads of this batch start at offsets[0] in the caller's arrays; device copies start at 0
  const uint8_t *seq0 = seq_format == 1 ? seq : seq + offsets[0];
  CU(cudaMemcpyAsync(st.seq, seq0, seq_bytes, cudaMemcpyHostToDevice, ctx->s_copy));
  if (qual) CU(cudaMemcpyAsync(st.qual, qual + offsets[0], n_bases, cudaMemcpyHostToDevice, ctx->s_copy));
  CU(cudaMemcpyAsync(st.offsets, offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->s_copy));
  if (seq_format == 1) {
    CU(cudaMemcpyAsync(st.nib, nib.data(), n_reads * 8, cudaMemcpyHostToDevice, ctx->s_copy));
    CU(cudaStreamSynchronize(ctx->s_copy));  // nib is a local vector
  }
  CU(cudaEventRecord(st.copied, ctx->s_copy));
  CU(cudaStreamWaitEvent(ctx->s_scan, st.copied, 0));
  PackParams Q;
  Q.seq = st.seq;
  Q.qual = qual ? st.qual : nullptr;
  Q.offsets = st.offsets;
  Q.nib_start = st.nib;
  Q.n_reads = (uint32_t)n_reads;
  Q.n_pos = n_pos;
  Q.n_bwords = (uint32_t)bw;
  Q.n_mwords = (uint32_t)mw;
  Q.min_baseq = min_baseq;
  Q.four_bit = seq_format;
  Q.bases2 = st.bases;
  Q.mask1 = st.mask;
  const int TB = 256;
  k_pack<<<(uint32_t)((mw + TB - 1) / TB), TB, 0, ctx->s_scan>>>(Q);
  CU(cudaGetLastError());
  rc = launch_scan(ctx, st.bases, st.mask, n_pos, sample);
  if (rc != DKB_OK) return rc;
  CU(cudaEventRecord(st.freed, ctx->s_scan));
  st.in_use = true;
  return DKB_OK;
}

int dkb_sync(dkb_ctx *ctx) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->s_copy));
  CU(cudaStreamSynchronize(ctx->s_scan));
  return DKB_OK;
}

int dkb_counts_reset(dkb_ctx *ctx) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->d_counts) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemsetAsync(ctx->d_counts, 0, (ctx->n_entries ? ctx->n_entries : 1) * 3 * 4, ctx->s_scan));
  if (ctx->prof) CU(cudaMemsetAsync(ctx->d_prof, 0, 4 * sizeof(unsigned long long), ctx->s_scan));
  ctx->finalised = false;
  return DKB_OK;
}

int dkb_entry_counts_fetch(dkb_ctx *ctx, uint32_t *out) {
  if (!ctx || !out) return fail(ctx, DKB_EINVAL, "null argument");
  if (!ctx->d_counts) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->s_copy));
  CU(cudaMemcpyAsync(out, ctx->d_counts, ctx->n_entries * 3 * 4, cudaMemcpyDeviceToHost, ctx->s_scan));
  CU(cudaStreamSynchronize(ctx->s_scan));
  return DKB_OK;
}

int dkb_entry_counts_device(dkb_ctx *ctx, void **d_ptr, size_t *n_u32) {
  if (!ctx || !d_ptr || !n_u32) return fail(ctx, DKB_EINVAL, "null argument");
  if (!ctx->d_counts) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  *d_ptr = ctx->d_counts;
  *n_u32 = ctx->n_entries * 3;
  return DKB_OK;
}

int dkb_finalise(dkb_ctx *ctx, const dkb_thresholds *thr) {
  if (!ctx || !thr) return fail(ctx, DKB_EINVAL, "null argument");
  if (!ctx->d_counts) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  return dkb_finalise_from(ctx, thr, ctx->d_counts);
}

int dkb_finalise_from(dkb_ctx *ctx, const dkb_thresholds *thr, const uint32_t *d_counts) {
  if (!ctx || !thr || !d_counts) return fail(ctx, DKB_EINVAL, "null argument");
  if (!ctx->d_counts) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  CU(cudaSetDevice(ctx->device));
  return finalise_counts(ctx, thr, d_counts);
}

}  // extern "C"

namespace {
int finalise_counts(dkb_ctx *ctx, const dkb_thresholds *thr, const uint32_t *d_counts) {
  cudaStream_t st = ctx->s_scan;
  const size_t nv1 = ctx->n_variants ? ctx->n_variants : 1;
  CU(cudaMemsetAsync(ctx->d_hits, 0, nv1 * 14 * 4, st));
  const int TB = 256;
  if (ctx->n_entries)
    k_variant_reduce<<<(uint32_t)((ctx->n_entries + TB - 1) / TB), TB, 0, st>>>(
        d_counts, ctx->d_variant, ctx->d_allele, ctx->d_dead, (uint32_t)ctx->n_entries,
        ctx->d_hits, ctx->d_distinct, ctx->d_nkmers);
  Thresholds T{thr->min_child_alt_hits, thr->min_child_alt_distinct, thr->max_parent_alt_hits,
               thr->min_parent_ref_hits};
  if (ctx->n_variants)
    k_calls<<<(ctx->n_variants + TB - 1) / TB, TB, 0, st>>>(ctx->d_hits, ctx->d_distinct,
                                                            ctx->n_variants, T, ctx->d_calls);
  CU(cudaGetLastError());
  ctx->finalised = true;
  return DKB_OK;
}
}  // namespace

extern "C" {

// ---- multi-GPU: one sum over ranks of the per-entry counters ---------------------------
int dkb_comm_unique_id(void *id_out) {
  dkb_ctx *ctx = nullptr;
  if (!id_out) return fail(nullptr, DKB_EINVAL, "id_out is null");
  if (!nccl()->h) return fail(nullptr, DKB_ENCCL, "cannot load libnccl.so.2: " + nccl()->err);
  NcclId id;
  NC(nccl()->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return DKB_OK;
}

int dkb_comm_init(dkb_ctx *ctx, const void *id, int rank, int world) {
  if (!ctx || !id) return fail(ctx, DKB_EINVAL, "null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, DKB_EINVAL, "rank / world out of range");
  if (ctx->comm) return fail(ctx, DKB_ESTATE, "communicator already initialised");
  if (!nccl()->h) return fail(ctx, DKB_ENCCL, "cannot load libnccl.so.2: " + nccl()->err);
  CU(cudaSetDevice(ctx->device));
  NcclId nid;
  memcpy(&nid, id, sizeof(nid));
  void *comm = nullptr;
  NC(nccl()->CommInitRank(&comm, world, nid, rank));
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return DKB_OK;
}

int dkb_comm_destroy(dkb_ctx *ctx) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->comm) return DKB_OK;
  CU(cudaSetDevice(ctx->device));
  if (ctx->s_side) CU(cudaStreamSynchronize(ctx->s_side));
  CU(cudaStreamSynchronize(ctx->s_scan));
  NC(nccl()->CommDestroy(ctx->comm));
  ctx->comm = nullptr;
  ctx->rank = 0;
  ctx->world = 1;
  return DKB_OK;
}

int dkb_comm_info(const dkb_ctx *ctx_c, int *rank, int *world, int *nccl_version) {
  dkb_ctx *ctx = const_cast<dkb_ctx *>(ctx_c);
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (rank) *rank = ctx->rank;
  if (world) *world = ctx->world;
  if (world && ctx->comm && nccl()->CommCount) NC(nccl()->CommCount(ctx->comm, world));
  if (nccl_version) {
    *nccl_version = 0;
    if (ctx->comm && nccl()->GetVersion) nccl()->GetVersion(nccl_version);
  }
  return DKB_OK;
}

int dkb_counts_allreduce(dkb_ctx *ctx) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->d_counts) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  if (!ctx->comm || ctx->world == 1) return DKB_OK;
  CU(cudaSetDevice(ctx->device));
  if (ctx->n_entries)
    NC(nccl()->AllReduce(ctx->d_counts, ctx->d_counts, ctx->n_entries * 3, NCCL_UINT32, NCCL_SUM,
                         ctx->comm, ctx->s_scan));
  ctx->finalised = false;
  return DKB_OK;
}

int dkb_reduce_push(dkb_ctx *ctx, const dkb_thresholds *thr) {
  if (!ctx || !thr) return fail(ctx, DKB_EINVAL, "null argument");
  if (!ctx->d_counts) return fail(ctx, DKB_ESTATE, "dkb_table_build must come first");
  CU(cudaSetDevice(ctx->device));
  int rc = reduce_setup(ctx);
  if (rc != DKB_OK) return rc;
  const int b = ctx->red_next;
  ctx->red_next ^= 1;
  const size_t n = ctx->n_entries * 3;
  // snapshot on the scan stream (behind this batch's scans) ...
  if (n) CU(cudaMemcpyAsync(ctx->d_red[b], ctx->d_counts, n * 4, cudaMemcpyDeviceToDevice, ctx->s_scan));
  CU(cudaEventRecord(ctx->ev_snap[b], ctx->s_scan));
  // ... summed over ranks on the side stream while the scan stream moves on
  CU(cudaStreamWaitEvent(ctx->s_side, ctx->ev_snap[b], 0));
  if (ctx->comm && ctx->world > 1 && n)
    NC(nccl()->AllReduce(ctx->d_red[b], ctx->d_red[b], n, NCCL_UINT32, NCCL_SUM, ctx->comm, ctx->s_side));
  CU(cudaEventRecord(ctx->ev_red[b], ctx->s_side));
  const int prev = ctx->red_pending;
  ctx->red_pending = b;
  if (prev >= 0) return finalise_snapshot(ctx, thr, prev);
  return DKB_OK;
}

int dkb_reduce_flush(dkb_ctx *ctx, const dkb_thresholds *thr) {
  if (!ctx || !thr) return fail(ctx, DKB_EINVAL, "null argument");
  if (ctx->red_pending < 0) return DKB_OK;
  CU(cudaSetDevice(ctx->device));
  const int b = ctx->red_pending;
  ctx->red_pending = -1;
  return finalise_snapshot(ctx, thr, b);
}

int dkb_reduced_counts_fetch(dkb_ctx *ctx, uint32_t *out) {
  if (!ctx || !out) return fail(ctx, DKB_EINVAL, "null argument");
  if (ctx->red_last < 0) return fail(ctx, DKB_ESTATE, "no reduced batch has been finalised yet");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(out, ctx->d_red[ctx->red_last], ctx->n_entries * 3 * 4, cudaMemcpyDeviceToHost,
                     ctx->s_scan));
  CU(cudaStreamSynchronize(ctx->s_scan));
  return DKB_OK;
}

int dkb_results_fetch(dkb_ctx *ctx, uint32_t *hits, uint32_t *distinct, uint32_t *n_kmers,
                      uint8_t *calls) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (!ctx->finalised) return fail(ctx, DKB_ESTATE, "dkb_finalise must come first");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->s_scan;
  const size_t nv = ctx->n_variants;
  if (hits) CU(cudaMemcpyAsync(hits, ctx->d_hits, nv * 6 * 4, cudaMemcpyDeviceToHost, st));
  if (distinct) CU(cudaMemcpyAsync(distinct, ctx->d_distinct, nv * 6 * 4, cudaMemcpyDeviceToHost, st));
  if (n_kmers) CU(cudaMemcpyAsync(n_kmers, ctx->d_nkmers, nv * 2 * 4, cudaMemcpyDeviceToHost, st));
  if (calls) CU(cudaMemcpyAsync(calls, ctx->d_calls, nv, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return DKB_OK;
}

int dkb_stats_get(dkb_ctx *ctx, dkb_stats *out) {
  if (!ctx || !out) return fail(ctx, DKB_EINVAL, "null argument");
  memset(out, 0, sizeof(*out));
  CU(cudaSetDevice(ctx->device));
  out->n_entries = ctx->n_entries;
  out->n_live_entries = ctx->n_live;
  out->table_slots = ctx->table_slots;
  out->n_seeds = ctx->n_seeds;
  out->seed_slots = ctx->seed_slots;
  out->bloom_words = ctx->bloom_words;
  out->bloom_bits_set = ctx->bloom_bits_set;
  out->scan_launches = ctx->scan_launches;
  out->positions_scanned = ctx->positions_scanned;
  CU(cudaStreamSynchronize(ctx->s_scan));
  unsigned long long prof[4];
  CU(cudaMemcpy(prof, ctx->d_prof, sizeof(prof), cudaMemcpyDeviceToHost));
  out->bloom_hits = prof[0];
  out->seed_hits = prof[1];
  out->windows_probed = prof[2];
  out->window_hits = prof[3];
  collect_timing(ctx);
  out->scan_launches_timed = ctx->scan_launches_timed;
  out->scan_ms_total = ctx->scan_ms_total;
  out->last_scan_ms = ctx->last_scan_ms;
  out->prefilter_words = ctx->gf ? ctx->pre_words : 0;
  out->gated_lookups = ctx->gate ? 1u : 0u;
  return DKB_OK;
}

int dkb_profile_counters(dkb_ctx *ctx, int enable) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  if (enable && !ctx->prof) {  // counters start from zero when profiling is switched on
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(ctx->d_prof, 0, 4 * sizeof(unsigned long long), ctx->s_scan));
  }
  ctx->prof = enable != 0;
  return DKB_OK;
}

int dkb_thread_bind_near_gpu(dkb_ctx *ctx, int *numa_node_out) {
  if (!ctx) return fail(nullptr, DKB_EINVAL, "ctx is null");
  const int node = gpu_numa_node(ctx->device);
  if (numa_node_out) *numa_node_out = node;
  cpu_set_t set;
  if (node >= 0 && node_cpus(node, &set)) sched_setaffinity(0, sizeof(set), &set);  // best effort
  return DKB_OK;
}

int dkb_host_alloc(dkb_ctx *ctx, size_t bytes, void **out) {
  if (!ctx || !out) return fail(ctx, DKB_EINVAL, "null argument");
  *out = nullptr;
  CU(cudaSetDevice(ctx->device));
  // allocate (and thereby pin: the pages are populated at once) while this thread runs on,
  // and prefers memory of, the NUMA node the GPU hangs off; then put the thread back
  const int node = gpu_numa_node(ctx->device);
  cpu_set_t old, near;
  const bool have_old = sched_getaffinity(0, sizeof(old), &old) == 0;
  const bool moved = node >= 0 && node_cpus(node, &near) && sched_setaffinity(0, sizeof(near), &near) == 0;
  if (node >= 0) prefer_node(node);
  void *p = nullptr;
  const cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
  if (node >= 0) prefer_node(-1);
  if (moved && have_old) sched_setaffinity(0, sizeof(old), &old);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, DKB_ENOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  }
  *out = p;
  return DKB_OK;
}

int dkb_host_free(dkb_ctx *ctx, void *p) {
  if (!p) return DKB_OK;
  if (ctx) cudaSetDevice(ctx->device);
  if (cudaFreeHost(p) != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, DKB_ECUDA, "cudaFreeHost failed");
  }
  return DKB_OK;
}

int dkb_scan_stream(dkb_ctx *ctx, void **stream_out) {
  if (!ctx || !stream_out) return fail(ctx, DKB_EINVAL, "null argument");
  *stream_out = (void *)ctx->s_scan;
  return DKB_OK;
}

}  // extern "C"
