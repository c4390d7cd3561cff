"""Synthetic trios of the shapes BASELINE.json names: a seeded random
genome, planted SNV / indel candidate DNMs, fixed-length reads at a target depth with
sequencing errors, N bases and a base-quality model.  Data generation only — nothing
here counts k-mers.

Two generators with the same model:
  * `make_trio_host`   — NumPy, decoded reads (ASCII + qualities), for parity tests and
                         the CPU baseline sample;
  * `make_sample_device` — PyTorch on the GPU, emits the packed stream (2-bit bases +
                         1-bit flags + separators) directly in HBM for the bench.
"""
from dataclasses import dataclass, field

import numpy as np

ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP_ASCII = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    COMP_ASCII[_a] = _b


@dataclass
class Variant:
    pos: int      # 0-based position of REF's first base in the genome
    ref: str
    alt: str
    inherited: bool = False  # also carried by the mother (should NOT be called de novo)


@dataclass
class Trio:
    k: int
    genome: np.ndarray  # uint8 ASCII
    variants: list
    reads: dict = field(default_factory=dict)  # sample -> (seq uint8, qual uint8, offsets uint64)

    def variant_tuples(self):
        """(left flank, ref, alt, right flank) per variant, flanks k-1 long (shorter at ends)."""
        g = self.genome
        out = []
        for v in self.variants:
            lo = max(0, v.pos - (self.k - 1))
            end = v.pos + len(v.ref)
            out.append((g[lo:v.pos].tobytes().decode(), v.ref, v.alt,
                        g[end:end + self.k - 1].tobytes().decode()))
        return out


def make_genome(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return ASCII[rng.integers(0, 4, size=n, dtype=np.uint8)]


def plant_variants(genome: np.ndarray, n_variants: int, k: int, seed: int,
                   indel_frac: float = 0.0, inherited_frac: float = 0.1,
                   max_indel: int = 8) -> list:
    """Variants on a regular grid with jitter, at least 2k apart, away from the ends."""
    rng = np.random.default_rng(seed)
    n = len(genome)
    margin = 4 * k
    if n_variants == 0:
        return []
    step = (n - 2 * margin) // n_variants
    if step < 3 * k:
        raise ValueError("genome too short for this many variants")
    out = []
    for i in range(n_variants):
        pos = margin + i * step + int(rng.integers(0, step - 2 * k - max_indel))
        ref_base = chr(genome[pos])
        r = rng.random()
        if r < indel_frac / 2:  # insertion after the anchor base
            ins = "".join("ACGT"[b] for b in rng.integers(0, 4, size=int(rng.integers(1, max_indel + 1))))
            ref, alt = ref_base, ref_base + ins
        elif r < indel_frac:    # deletion of the bases after the anchor
            dl = int(rng.integers(1, max_indel + 1))
            ref, alt = genome[pos:pos + 1 + dl].tobytes().decode(), ref_base
        else:
            alt_base = "ACGT"[(("ACGT".index(ref_base)) + int(rng.integers(1, 4))) % 4]
            ref, alt = ref_base, alt_base
        out.append(Variant(pos, ref, alt, inherited=bool(rng.random() < inherited_frac)))
    return out


def apply_variants(genome: np.ndarray, variants) -> np.ndarray:
    """The alternate haplotype: genome with every given variant applied."""
    parts, cur = [], 0
    for v in sorted(variants, key=lambda x: x.pos):
        parts.append(genome[cur:v.pos])
        parts.append(np.frombuffer(v.alt.encode(), dtype=np.uint8))
        cur = v.pos + len(v.ref)
    parts.append(genome[cur:])
    return np.concatenate(parts)


def sample_reads(haps, n_reads: int, read_len: int, seed: int, err_rate: float = 0.002,
                 n_rate: float = 0.0005, lowq_frac: float = 0.03, rc_frac: float = 0.5,
                 ragged: bool = False):
    """Reads drawn uniformly from the given haplotypes (equal weight).  Returns
    (seq ASCII uint8, qual uint8, offsets uint64[n_reads+1])."""
    rng = np.random.default_rng(seed)
    hap_of = rng.integers(0, len(haps), size=n_reads)
    lens = np.full(n_reads, read_len, dtype=np.int64)
    if ragged:
        lens = rng.integers(1, read_len + 1, size=n_reads)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    seq = np.empty(total, dtype=np.uint8)
    for h, hap in enumerate(haps):
        idx = np.nonzero(hap_of == h)[0]
        if len(idx) == 0:
            continue
        starts = rng.integers(0, len(hap) - read_len + 1, size=len(idx))
        if not ragged:  # fixed length: one gather
            seq.reshape(n_reads, read_len)[idx] = hap[starts[:, None] + np.arange(read_len)[None, :]]
            continue
        for r, s in zip(idx, starts):
            seq[int(offsets[r]):int(offsets[r + 1])] = hap[s:s + int(lens[r])]
    # sequencing errors: substitute with a different base
    err = rng.random(total) < err_rate
    if err.any():
        code = np.searchsorted(ASCII, seq[err])  # ASCII is sorted A<C<G<T
        seq[err] = ASCII[(code + rng.integers(1, 4, size=int(err.sum()))) % 4]
    # reverse-complement a fraction of reads (reads from either strand)
    flip = rng.random(n_reads) < rc_frac
    if not ragged:
        m = seq.reshape(n_reads, read_len)
        m[flip] = COMP_ASCII[m[flip][:, ::-1]]
    else:
        for r in np.nonzero(flip)[0]:
            a, b = int(offsets[r]), int(offsets[r + 1])
            seq[a:b] = COMP_ASCII[seq[a:b][::-1]]
    seq[rng.random(total) < n_rate] = ord("N")
    qual = rng.integers(25, 41, size=total).astype(np.uint8)
    low = rng.random(total) < lowq_frac
    qual[low] = rng.integers(2, 20, size=int(low.sum())).astype(np.uint8)
    return seq, qual, offsets


def make_trio_host(genome_len: int, depth: float, n_variants: int, k: int, seed: int = 1,
                   read_len: int = 150, indel_frac: float = 0.0, ragged: bool = False,
                   **read_kw) -> Trio:
    """Child is heterozygous for every variant; the mother also carries the
    `inherited` ones; the father carries none."""
    genome = make_genome(genome_len, seed)
    variants = plant_variants(genome, n_variants, k, seed + 1, indel_frac=indel_frac)
    child_alt = apply_variants(genome, variants)
    mother_alt = apply_variants(genome, [v for v in variants if v.inherited])
    n_reads = int(genome_len * depth / read_len)
    trio = Trio(k, genome, variants)
    trio.reads[0] = sample_reads([genome, child_alt], n_reads, read_len, seed + 10, ragged=ragged, **read_kw)
    trio.reads[1] = sample_reads([genome, mother_alt], n_reads, read_len, seed + 11, ragged=ragged, **read_kw)
    trio.reads[2] = sample_reads([genome, genome], n_reads, read_len, seed + 12, ragged=ragged, **read_kw)
    return trio


# ---------------------------------------------------------------------------------------
# GPU generator: packed streams built in HBM (bench workload; no host copy of the reads)
# ---------------------------------------------------------------------------------------
def make_sample_device(haps_codes, n_reads: int, read_len: int, seed: int, device,
                       err_rate: float = 0.002, n_rate: float = 0.0005, lowq_frac: float = 0.03,
                       rc_frac: float = 0.5, chunk_reads: int = 1 << 20, return_meta: bool = False):
    """haps_codes: list of uint8 CUDA tensors of base codes 0..3 (one per haplotype).
    Returns (bases2 int32 tensor, mask1 int32 tensor, n_positions, n_bases) in the layout
    of include/dkb.h.  Same read model as `sample_reads`; low-quality and N bases only
    clear the flag (what the host packer would emit)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    stride = read_len + 1  # one separator per read
    n_pos = n_reads * stride
    bw = (n_pos + 63) // 64 * 4
    mw = (n_pos + 127) // 128 * 4
    bases2 = torch.zeros(bw, dtype=torch.int32, device=device)
    mask1 = torch.zeros(mw, dtype=torch.int32, device=device)
    # chunks must cover whole 128-position groups: chunk_reads * stride % 128 == 0
    chunk_reads = max(128, chunk_reads // 128 * 128)
    ar = torch.arange(read_len, device=device)
    sh2 = (2 * torch.arange(16, device=device, dtype=torch.int64))
    sh1 = torch.arange(32, device=device, dtype=torch.int64)
    meta_h, meta_s = [], []
    for r0 in range(0, n_reads, chunk_reads):
        nr = min(chunk_reads, n_reads - r0)
        codes = torch.zeros((nr, stride), dtype=torch.uint8, device=device)
        valid = torch.zeros((nr, stride), dtype=torch.bool, device=device)
        hap_of = torch.randint(0, len(haps_codes), (nr,), generator=g, device=device)
        body = torch.empty((nr, read_len), dtype=torch.uint8, device=device)
        all_starts = torch.zeros((nr,), dtype=torch.int64, device=device)
        for h, hap in enumerate(haps_codes):
            sel = (hap_of == h).nonzero().squeeze(1)
            if sel.numel() == 0:
                continue
            starts = torch.randint(0, hap.numel() - read_len + 1, (sel.numel(),), generator=g,
                                   device=device)
            body[sel] = hap[starts[:, None] + ar[None, :]]
            all_starts[sel] = starts
        if return_meta:
            meta_h.append(hap_of)
            meta_s.append(all_starts)
        err = torch.rand((nr, read_len), generator=g, device=device) < err_rate
        bump = torch.randint(1, 4, (nr, read_len), generator=g, device=device, dtype=torch.uint8)
        body = torch.where(err, (body + bump) & 3, body)
        flip = torch.rand((nr,), generator=g, device=device) < rc_frac
        body = torch.where(flip[:, None], 3 - body.flip(1), body)
        ok = torch.rand((nr, read_len), generator=g, device=device) >= (n_rate + lowq_frac)
        codes[:, :read_len] = torch.where(ok, body, torch.zeros_like(body))
        valid[:, :read_len] = ok
        flat_c = codes.reshape(-1)
        flat_v = valid.reshape(-1)
        p0 = r0 * stride
        pad = (-flat_c.numel()) % 128
        if pad:
            flat_c = torch.cat([flat_c, flat_c.new_zeros(pad)])
            flat_v = torch.cat([flat_v, flat_v.new_zeros(pad)])
        wb = (flat_c.reshape(-1, 16).to(torch.int64) << sh2).sum(1)
        wm = (flat_v.reshape(-1, 32).to(torch.int64) << sh1).sum(1)
        # int64 -> wrap into int32 bit patterns
        wb = torch.where(wb >= 2 ** 31, wb - 2 ** 32, wb).to(torch.int32)
        wm = torch.where(wm >= 2 ** 31, wm - 2 ** 32, wm).to(torch.int32)
        assert p0 % 128 == 0
        bases2[p0 // 16: p0 // 16 + wb.numel()] = wb[: bases2.numel() - p0 // 16]
        mask1[p0 // 32: p0 // 32 + wm.numel()] = wm[: mask1.numel() - p0 // 32]
    if return_meta:
        return bases2, mask1, n_pos, n_reads * read_len, (torch.cat(meta_h), torch.cat(meta_s))
    return bases2, mask1, n_pos, n_reads * read_len
