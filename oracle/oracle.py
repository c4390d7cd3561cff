"""ctypes front-end of oracle/libdnk_oracle.so (built by oracle/Makefile) plus a
pure-Python restatement (`py_*`) of the same spec for tiny cases, used to check
the C oracle itself.  TEST INFRASTRUCTURE ONLY — see oracle/dnk_oracle.c header.
Spec: DESIGN.md §2 (reference src/kmer.rs, src/counter.rs are not mounted)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdnk_oracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def build(force=False):
    src = os.path.join(_HERE, "dnk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libdnk_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_base_code.restype = C.c_int
        L.orc_base_code.argtypes = [C.c_uint8]
        L.orc_revcomp.restype = C.c_uint64
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.orc_canonical.restype = C.c_uint64
        L.orc_canonical.argtypes = [C.c_uint64, C.c_int]
        L.orc_read_kmers.restype = C.c_size_t
        L.orc_read_kmers.argtypes = [u8p, u8p, C.c_size_t, C.c_int, C.c_int, u64p, u32p]
        L.orc_set_build.restype = C.c_void_p
        L.orc_set_build.argtypes = [u64p, u32p, u8p, C.c_size_t]
        L.orc_set_free.argtypes = [C.c_void_p]
        L.orc_set_live.argtypes = [C.c_void_p, u8p]
        L.orc_count_reads.argtypes = [C.c_void_p, u8p, u8p, u64p, C.c_size_t, C.c_int, C.c_int,
                                      u64p, C.c_int]
        L.orc_count_stream.argtypes = [C.c_void_p, u32p, u32p, C.c_uint64, C.c_int, u64p, C.c_int]
        L.orc_variant_stats.argtypes = [C.c_void_p, u32p, u8p, u64p, C.c_size_t, u64p, u64p, u32p]
        L.orc_calls.argtypes = [u64p, u64p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                C.c_uint32, u8p]
        L.orc_allele_kmers.restype = C.c_size_t
        L.orc_allele_kmers.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, u64p, u16p,
                                       u16p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def revcomp(fwd, k):
    return int(lib().orc_revcomp(int(fwd), k))


def canonical(fwd, k):
    return int(lib().orc_canonical(int(fwd), k))


def read_kmers(seq: bytes, qual, k, min_bq=0):
    s = np.frombuffer(seq, dtype=np.uint8)
    q = None if qual is None else np.ascontiguousarray(qual, dtype=np.uint8)
    out = np.zeros(max(len(s), 1), dtype=np.uint64)
    pos = np.zeros(max(len(s), 1), dtype=np.uint32)
    n = lib().orc_read_kmers(_p(s, u8p), _p(q, u8p), len(s), k, min_bq, _p(out, u64p),
                             _p(pos, u32p))
    return out[:n].copy(), pos[:n].copy()


class KmerSet:
    """counter.rs stand-in: key -> owners map with per-entry counts."""

    def __init__(self, keys, variant, allele):
        self.keys = np.ascontiguousarray(keys, dtype=np.uint64)
        self.variant = np.ascontiguousarray(variant, dtype=np.uint32)
        self.allele = np.ascontiguousarray(allele, dtype=np.uint8)
        self.n = len(self.keys)
        self._h = lib().orc_set_build(_p(self.keys, u64p), _p(self.variant, u32p),
                                      _p(self.allele, u8p), self.n)
        if not self._h:
            raise MemoryError("orc_set_build")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_set_free(self._h)
            self._h = None

    def live(self):
        out = np.zeros(max(self.n, 1), dtype=np.uint8)
        lib().orc_set_live(self._h, _p(out, u8p))
        return out[: self.n]

    def count_reads(self, seq, qual, offsets, k, min_bq=0, counts=None, threads=0):
        """seq/qual: uint8 arrays of concatenated reads; offsets: uint64[n_reads+1]."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        qual = None if qual is None else np.ascontiguousarray(qual, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if counts is None:
            counts = np.zeros(self.n, dtype=np.uint64)
        lib().orc_count_reads(self._h, _p(seq, u8p), _p(qual, u8p), _p(offsets, u64p),
                              len(offsets) - 1, k, min_bq, _p(counts, u64p), threads)
        return counts

    def count_stream(self, bases2, mask1, n_positions, k, counts=None, threads=0):
        """Same count from a packed stream (include/dkb.h layout): uint32 arrays."""
        bases2 = np.ascontiguousarray(bases2).view(np.uint32)
        mask1 = np.ascontiguousarray(mask1).view(np.uint32)
        assert len(bases2) * 16 >= n_positions and len(mask1) * 32 >= n_positions
        if counts is None:
            counts = np.zeros(self.n, dtype=np.uint64)
        lib().orc_count_stream(self._h, _p(bases2, u32p), _p(mask1, u32p), int(n_positions), k,
                               _p(counts, u64p), threads)
        return counts

    def variant_stats(self, counts3, n_variants):
        """counts3: uint64 [3, n_entries] -> hits[nv,2,3], distinct[nv,2,3], n_kmers[nv,2]."""
        counts3 = np.ascontiguousarray(counts3, dtype=np.uint64).reshape(3, self.n)
        hits = np.zeros((n_variants, 2, 3), dtype=np.uint64)
        dist = np.zeros((n_variants, 2, 3), dtype=np.uint64)
        nk = np.zeros((n_variants, 2), dtype=np.uint32)
        lib().orc_variant_stats(self._h, _p(self.variant, u32p), _p(self.allele, u8p),
                                _p(counts3, u64p), n_variants, _p(hits, u64p), _p(dist, u64p),
                                _p(nk, u32p))
        return hits, dist, nk


def pack_stream(seq, qual, offsets, min_bq=0):
    """NumPy restatement of the packed read stream of include/dkb.h (bases2, mask1,
    n_positions): 2-bit codes A=0 C=1 G=2 T=3, flag = usable base, one flag-0 separator
    position after every read.  Independent of the library's packers (which are checked
    against it)."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64).astype(np.int64)
    n_reads = len(offsets) - 1
    lens = np.diff(offsets)
    base0 = int(offsets[0]) if n_reads else 0
    n_bases = int(offsets[-1]) - base0 if n_reads else 0
    n_pos = n_bases + n_reads
    code = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate("ACGT"):
        code[ord(ch)] = i
        code[ord(ch.lower())] = i
    c = code[seq[base0:base0 + n_bases]]
    ok = c != 255
    if qual is not None:
        ok &= np.ascontiguousarray(qual, dtype=np.uint8)[base0:base0 + n_bases] >= min_bq
    # stream position of every base: its index plus the number of reads that ended before it
    read_of = np.repeat(np.arange(n_reads, dtype=np.int64), lens)
    pos = np.arange(n_bases, dtype=np.int64) + read_of
    bw = (n_pos + 63) // 64 * 4
    mw = (n_pos + 127) // 128 * 4
    codes = np.zeros(bw * 16, dtype=np.uint64)
    flags = np.zeros(mw * 32, dtype=np.uint64)
    codes[pos] = np.where(ok, c, 0)
    flags[pos] = ok
    bases2 = (codes.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint64))).sum(1).astype(np.uint32)
    mask1 = (flags.reshape(-1, 32) << np.arange(32, dtype=np.uint64)).sum(1).astype(np.uint32)
    return bases2, mask1, n_pos


def expand_zero_list(zoff, zbytes, n_positions):
    """Plain-Python reading of the zero-list form of the flags (include/dkb.h): returns the
    flags of positions [0, n_positions) as a uint8 array.  Checks the host encoder and the
    device expander."""
    flags = np.ones(n_positions, dtype=np.uint8)
    nb = (n_positions + 2047) // 2048
    assert len(zoff) == nb + 1
    for b in range(nb):
        p0 = b * 2048
        o0, o1 = int(zoff[b]), int(zoff[b + 1]) & 0x7FFFFFFF
        raw = bool(o0 >> 31)
        o0 &= 0x7FFFFFFF
        if raw:
            assert o1 - o0 == 256
            bits = np.unpackbits(np.asarray(zbytes[o0:o1], dtype=np.uint8), bitorder="little")
            n = min(2048, n_positions - p0)
            flags[p0:p0 + n] = bits[:n]
            continue
        p = p0
        for v in np.asarray(zbytes[o0:o1]).tolist():
            if v == 255:
                p += 255
            else:
                p += v
                assert p < min(p0 + 2048, n_positions), "zero beyond its block"
                flags[p] = 0
                p += 1
    return flags


def calls(hits, distinct, thr):
    """thr = (min_child_alt_hits, min_child_alt_distinct, max_parent_alt_hits, min_parent_ref_hits)"""
    hits = np.ascontiguousarray(hits, dtype=np.uint64)
    distinct = np.ascontiguousarray(distinct, dtype=np.uint64)
    nv = hits.shape[0]
    out = np.zeros(max(nv, 1), dtype=np.uint8)
    lib().orc_calls(_p(hits, u64p), _p(distinct, u64p), nv, *[int(t) for t in thr], _p(out, u8p))
    return out[:nv]


def allele_kmers(left: str, allele: str, right: str, k: int):
    cap = len(allele) + k + 2
    keys = np.zeros(cap, dtype=np.uint64)
    win = np.zeros(cap, dtype=np.uint16)
    run = C.c_uint16(0)
    n = lib().orc_allele_kmers(left.encode(), allele.encode(), right.encode(), k, _p(keys, u64p),
                               _p(win, u16p), C.byref(run))
    return keys[:n].copy(), win[:n].copy(), int(run.value)


def variant_entries(variants, k, drop_shared=True):
    """variants: list of (left, ref, alt, right) strings.  Returns entry arrays
    (keys, variant, allele, win_index, win_count) under the builder policy of
    include/dkb.h dkb_variant_kmers: per (variant, allele) first window wins;
    keys present in both alleles of a variant dropped when drop_shared."""
    K, V, A, WI, WC = [], [], [], [], []
    for v, (left, ref, alt, right) in enumerate(variants):
        per = []
        for allele in (ref, alt):
            keys, win, run = allele_kmers(left, allele, right, k)
            seen, lst = set(), []
            for key, w in zip(keys.tolist(), win.tolist()):
                if key not in seen:
                    seen.add(key)
                    lst.append((key, w))
            per.append((lst, seen, run))
        for a in (0, 1):
            lst, _, run = per[a]
            other = per[1 - a][1]
            for key, w in lst:
                if drop_shared and key in other:
                    continue
                K.append(key); V.append(v); A.append(a); WI.append(w); WC.append(run)
    return (np.array(K, dtype=np.uint64), np.array(V, dtype=np.uint32), np.array(A, dtype=np.uint8),
            np.array(WI, dtype=np.uint16), np.array(WC, dtype=np.uint16))


# ---------------------------------------------------------------------------
# Pure-Python restatement for tiny cases (checks the C oracle; strings, no bits)
# ---------------------------------------------------------------------------
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def py_encode(s):
    v = 0
    for ch in s:
        v = v * 4 + _CODE[ch]
    return v


def py_canonical_str(s):
    rc = "".join(_COMP[c] for c in reversed(s))
    return min(py_encode(s), py_encode(rc))


def py_read_kmers(seq: str, qual, k, min_bq=0):
    """Canonical keys of every window whose k bases are ACGT (any case) with qual >= min_bq."""
    seq = seq.upper()
    out = []
    for w in range(0, len(seq) - k + 1):
        win = seq[w:w + k]
        if any(c not in _CODE for c in win):
            continue
        if qual is not None and any(q < min_bq for q in qual[w:w + k]):
            continue
        out.append((w, py_canonical_str(win)))
    return out


def py_count(entries, reads, k, min_bq=0):
    """entries: list of (key, variant, allele); reads: list of (seq, qual|None).
    Returns per-entry counts with first-wins de-duplication of repeated triples."""
    first = {}
    owners = {}
    for i, e in enumerate(entries):
        if e in first:
            continue
        first[e] = i
        owners.setdefault(e[0], []).append(i)
    counts = [0] * len(entries)
    for seq, qual in reads:
        for _, key in py_read_kmers(seq, qual, k, min_bq):
            for i in owners.get(key, ()):
                counts[i] += 1
    return counts
