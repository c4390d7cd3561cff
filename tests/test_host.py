"""CPU: the C-ABI library loads and exports every symbol include/dkb.h declares, and its
host-side pieces (k-mer primitives, read packer, variant k-mer builder) match the oracle.
No compute entry point is called (no GPU here)."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_library_exports_every_declared_symbol(dkb):
    from denovo_kmer_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dkb.h")).read()
    declared = set(re.findall(r"\b(dkb_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = C.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().dkb_abi_version() == 1


def test_no_cpu_fallback(dkb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dkb.DkbError) as ei:
        dkb.KmerCounter(31)
    assert ei.value.code == 5  # DKB_ENODEV


def test_product_does_not_touch_oracle():
    pkg = os.path.join(ROOT, "denovo_kmer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower() or f == "__never__", f"{f} mentions the oracle"


@pytest.mark.parametrize("k", [8, 15, 16, 21, 31])
def test_kmer_primitives_match_oracle(dkb, orc, k):
    rnd = random.Random(k)
    for _ in range(300):
        s = "".join(rnd.choice("ACGTacgt") for _ in range(k))
        f = dkb.kmer_encode(s)
        assert f == orc.py_encode(s.upper())
        assert dkb.kmer_revcomp(f, k) == orc.revcomp(f, k)
        assert dkb.kmer_canonical(f, k) == orc.canonical(f, k) == orc.py_canonical_str(s.upper())
    with pytest.raises(dkb.DkbError):
        dkb.kmer_encode("ACGN" * 8, k)


def _py_pack(reads, min_bq):
    """Reference packing in plain Python: list of (seq, qual|None) -> (codes, flags)."""
    codes, flags = [], []
    for s, q in reads:
        for i, ch in enumerate(s):
            c = "ACGT".find(ch.upper())
            ok = c >= 0 and (q is None or q[i] >= min_bq)
            codes.append(c if ok else 0)
            flags.append(1 if ok else 0)
        codes.append(0)
        flags.append(0)
    return codes, flags


def test_pack_reads(dkb):
    rnd = random.Random(4)
    reads = []
    for _ in range(200):
        n = rnd.randint(0, 200)
        reads.append(("".join(rnd.choice("ACGTNacgtRY") for _ in range(n)),
                      [rnd.randint(0, 41) for _ in range(n)]))
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s, _ in reads])
    seq = np.frombuffer("".join(s for s, _ in reads).encode(), dtype=np.uint8)
    qual = np.array([x for _, q in reads for x in q], dtype=np.uint8)
    for min_bq, use_q in ((20, True), (0, True), (20, False)):
        st = dkb.pack_reads(seq, qual if use_q else None, off, min_bq)
        codes, flags = _py_pack([(s, q if use_q else None) for s, q in reads], min_bq)
        assert st.n_positions == len(codes) == int(off[-1]) + len(reads)
        assert st.n_bases == int(off[-1])
        p = np.arange(st.n_positions)
        got_c = (st.bases2[p >> 4] >> (2 * (p & 15)).astype(np.uint32)) & 3
        got_f = (st.mask1[p >> 5] >> (p & 31).astype(np.uint32)) & 1
        assert got_c.tolist() == codes and got_f.tolist() == flags
        bw, mw = dkb.stream_words(st.n_positions)
        assert len(st.bases2) == bw and len(st.mask1) == mw and bw % 4 == 0 and mw % 4 == 0
    empty = dkb.pack_reads(np.zeros(0, np.uint8), None, np.zeros(1, np.uint64), 20)
    assert empty.n_positions == 0


@pytest.mark.parametrize("k,indel_frac,drop", [(31, 0.0, True), (21, 0.6, True), (15, 0.6, False),
                                                (8, 1.0, True)])
def test_variant_kmers_match_oracle(dkb, orc, k, indel_frac, drop):
    from denovo_kmer_b200 import synth
    g = synth.make_genome(30_000, 21)
    g[100:104] = ord("N")  # an N near the start exercises window skipping
    vs = synth.plant_variants(g, 60, k, 22, indel_frac=indel_frac)
    vs.append(synth.Variant(3, "A", "C"))            # left flank shorter than k-1
    vs.append(synth.Variant(len(g) - 3, "A", "C"))   # right flank shorter than k-1
    vs.append(synth.Variant(110, chr(g[110]), "T" if chr(g[110]) != "T" else "G"))  # N in flank
    tup = synth.Trio(k, g, vs).variant_tuples()
    e = dkb.variant_kmers(tup, k, drop_shared=drop)
    keys, var, al, wi, wc = orc.variant_entries(tup, k, drop_shared=drop)
    assert e.keys.tolist() == keys.tolist() and e.variant.tolist() == var.tolist()
    assert e.allele.tolist() == al.tolist()
    assert (e.win_index & 0x7FFF).tolist() == wi.tolist() and e.win_count.tolist() == wc.tolist()
    assert e.n_variants == len(tup) and len(e) > 0
    # bit 15 of win_index: the haplotype window reads as the reverse complement of the key
    left, ref, alt, right = tup[0]
    hap = left[-(k - 1):] + ref + right[: k - 1]
    m = (e.variant == 0) & (e.allele == 0)
    for key, w in zip(e.keys[m].tolist(), e.win_index[m].tolist()):
        lo = max(0, len(left[-(k - 1):]) - k + 1)
        fwd = orc.py_encode(hap[lo + (w & 0x7FFF): lo + (w & 0x7FFF) + k])
        assert (fwd != key) == bool(w & 0x8000)


def test_argument_validation(dkb):
    from denovo_kmer_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.dkb_ctx_create(0, 32, C.byref(h)) == _lib.EINVAL
    assert L.dkb_ctx_create(0, 7, C.byref(h)) == _lib.EINVAL
    assert L.dkb_ctx_create(0, 31, None) == _lib.EINVAL
    assert L.dkb_sync(None) == _lib.EINVAL and L.dkb_ctx_destroy(None) == _lib.OK
    assert L.dkb_strerror(_lib.ENODEV).decode().startswith("no usable CUDA device")
    n = C.c_size_t(0)
    arr = (C.c_char_p * 1)(b"A")
    assert L.dkb_variant_kmers(arr, arr, arr, arr, 1, 40, 1, None, None, None, None, None,
                               C.byref(n)) == _lib.EINVAL


@pytest.mark.parametrize("ragged", [False, True])
def test_pack_reads_threaded_large(dkb, ragged):
    """Streams above 1 M positions are packed by several threads on 128-position cuts;
    compare every position with a NumPy restatement (fixed and ragged read lengths)."""
    from denovo_kmer_b200 import synth
    g = synth.make_genome(300_000, 9)
    seq, qual, off = synth.sample_reads([g], 30_000 if ragged else 12_000, 150, 3, ragged=ragged,
                                        n_rate=0.01, lowq_frac=0.1)
    st = dkb.pack_reads(seq, qual, off, 20)
    assert st.n_positions > (1 << 20)
    lut = np.full(256, 4, np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    lens = np.diff(off).astype(np.int64)
    pos = np.arange(len(seq)) + np.repeat(np.arange(len(off) - 1), lens)
    c = lut[seq]
    ok = (c < 4) & (qual >= 20)
    exp_c = np.zeros(st.n_positions, np.uint8)
    exp_f = np.zeros(st.n_positions, np.uint8)
    exp_c[pos] = np.where(ok, c, 0)
    exp_f[pos] = ok
    p = np.arange(st.n_positions)
    got_c = (st.bases2[p >> 4] >> (2 * (p & 15)).astype(np.uint32)) & 3
    got_f = (st.mask1[p >> 5] >> (p & 31).astype(np.uint32)) & 1
    assert np.array_equal(got_c, exp_c) and np.array_equal(got_f, exp_f)
    assert st.bases2[(st.n_positions + 15) // 16:].sum() == 0
    assert st.mask1[(st.n_positions + 31) // 32:].sum() == 0


@pytest.mark.parametrize("read_len,ragged", [(1, False), (33, False), (64, False), (65, False), (150, True), (251, False)])
def test_pack_reads_isa_paths_agree(dkb, read_len, ragged, monkeypatch):
    """The packer picks the widest path the CPU has (scalar, AVX2 + BMI2: 32 bases per step,
    AVX-512 BW + BMI2: 64 per step); DKB_PACK_ISA lowers it.  All paths must write the same words -
    lower case, non-ACGT bytes, every quality byte value, empty reads, a slice whose offsets do not
    start at 0, thresholds at and beyond the byte range, one thread and several."""
    from denovo_kmer_b200 import synth
    g = synth.make_genome(200_000, 9)
    seq, qual, off = synth.sample_reads([g], (2_500_000 if ragged else 1_200_000) // (read_len + 1) + 100, read_len, 3, ragged=ragged,
                                        n_rate=0.01, lowq_frac=0.1)
    rng = np.random.default_rng(read_len)
    seq, qual = seq.copy(), qual.copy()
    seq[rng.integers(0, len(seq), len(seq) // 50)] |= 0x20
    idx = rng.integers(0, len(seq), len(seq) // 300)
    seq[idx] = rng.integers(0, 256, len(idx)).astype(np.uint8)
    idx = rng.integers(0, len(qual), len(qual) // 100)
    qual[idx] = rng.integers(0, 256, len(idx)).astype(np.uint8)
    off = np.sort(np.concatenate([off, off[rng.integers(0, len(off), len(off) // 20)]]))  # empty reads
    for offs in (off, off[7:-5]):
        for threads in ("1", "3", "16"):
            monkeypatch.setenv("DKB_PACK_THREADS", threads)
            for mq, use_q in ((20, True), (0, True), (255, True), (256, True), (-3, True), (20, False)):
                outs = []
                for isa in ("0", "1", "2"):
                    monkeypatch.setenv("DKB_PACK_ISA", isa)
                    st = dkb.pack_reads(seq, qual if use_q else None, offs, mq)
                    outs.append((st.bases2.copy(), st.mask1.copy(), st.n_positions))
                for o in outs[1:]:
                    assert o[2] == outs[0][2] and np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]), \
                        (read_len, threads, mq, use_q)
    assert outs[0][2] > (1 << 20)  # large enough for the threaded split


def _to_bam4(seq, off):
    """ASCII reads -> BAM 4-bit codes (=ACMGRSVTWYHKDBN, high nibble first), every read on a byte
    boundary; the unused low nibble of an odd read's last byte gets a random code."""
    code = np.full(256, 15, np.uint8)
    for v, ch in enumerate(b"=ACMGRSVTWYHKDBN"):
        code[ch] = code[ord(chr(ch).lower())] = v
    lens = np.diff(off).astype(np.int64)
    starts = np.concatenate([[0], np.cumsum((lens + 1) // 2)])
    out = np.zeros(int(starts[-1]), np.uint8)
    r = np.repeat(np.arange(len(lens)), lens)
    i = np.arange(len(seq)) - np.repeat(off[:-1].astype(np.int64) - int(off[0]), lens)
    byte, hi, c = starts[r] + i // 2, i % 2 == 0, code[seq]
    np.add.at(out, byte[hi], c[hi] << 4)
    np.add.at(out, byte[~hi], c[~hi])
    odd = np.nonzero(lens % 2 == 1)[0]
    out[starts[odd + 1] - 1] |= np.random.default_rng(5).integers(0, 16, len(odd)).astype(np.uint8)
    return out


@pytest.mark.parametrize("read_len,ragged", [(1, False), (33, False), (65, False), (150, True), (251, False)])
def test_pack_reads_four_bit_equals_ascii(dkb, read_len, ragged, monkeypatch):
    """dkb_pack_reads_fmt with BAM 4-bit input (what record.seq().encoded holds: no decode pass in
    the BAM layer) writes the same stream as the ASCII form of the same reads - IUPAC codes and
    '=' are unusable like N, empty reads, odd lengths with a padding nibble, with and without
    qualities, on every code path and thread split."""
    from denovo_kmer_b200 import synth
    g = synth.make_genome(200_000, 9)
    seq, qual, off = synth.sample_reads([g], (2_500_000 if ragged else 1_200_000) // (read_len + 1) + 100, read_len, 3,
                                        ragged=ragged, n_rate=0.01, lowq_frac=0.1)
    rng = np.random.default_rng(read_len)
    seq = seq.copy()
    idx = rng.integers(0, len(seq), len(seq) // 100)
    seq[idx] = np.frombuffer(b"MRSVWYHKDBN=", np.uint8)[rng.integers(0, 12, len(idx))]
    off = np.sort(np.concatenate([off, off[rng.integers(0, len(off), len(off) // 20)]]))  # empty reads
    b4 = _to_bam4(seq, off)
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("DKB_PACK_THREADS", threads)
        for isa in ("0", "1", "2"):
            monkeypatch.setenv("DKB_PACK_ISA", isa)
            for q in (qual, None):
                a = dkb.pack_reads(seq, q, off, 20)
                b = dkb.pack_reads(b4, q, off, 20, four_bit=True)
                assert a.n_positions == b.n_positions and np.array_equal(a.bases2, b.bases2) and \
                    np.array_equal(a.mask1, b.mask1), (read_len, threads, isa, q is None)
    assert a.n_positions > (1 << 20)


def test_pack_reads_refuses_descending_offsets(dkb, monkeypatch):
    """Offsets are checked (by chunks of reads, on the packing threads) before any read is
    touched through them: a descending pair anywhere is DKB_EINVAL, for both input forms."""
    n = 40_000
    off = np.arange(n + 1, dtype=np.uint64) * 40
    seq = np.full(int(off[-1]), ord("A"), np.uint8)
    for threads in ("1", "4"):
        monkeypatch.setenv("DKB_PACK_THREADS", threads)
        assert dkb.pack_reads(seq, None, off, 20).n_positions == int(off[-1]) + n
        for where in (1, n // 2, n - 1):
            bad = off.copy()
            bad[where] = bad[where + 1] + 5
            for four_bit in (False, True):
                with pytest.raises(dkb.DkbError):
                    dkb.pack_reads(seq, None, bad, 20, four_bit=four_bit)


def test_host_code_under_asan(tmp_path):
    """csrc/dkb_host.cpp (no CUDA in it) compiled with AddressSanitizer + UBSan and driven by
    tests/asan_host.cpp: random batches through every packer path and the zero-list coder on
    exact-size heap buffers.  The SIMD paths use masked loads, zero-padded copies and 64-bit
    word stores; this is what shows that none of them reads or writes past a buffer."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / "asan_host")
    cc = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined",
                         "-fno-sanitize-recover=undefined", "-o", exe,
                         os.path.join(root, "denovo_kmer_b200", "csrc", "dkb_host.cpp"),
                         os.path.join(root, "tests", "asan_host.cpp"), "-pthread"], capture_output=True, text=True)
    if cc.returncode != 0 and "asan" in (cc.stderr or "").lower():
        pytest.skip("libasan is not installed")
    assert cc.returncode == 0, cc.stderr
    out = subprocess.run([exe, "120"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("asan ok"), (out.stdout[-500:], out.stderr[-2000:])


def test_header_is_plain_c_and_links(dkb, tmp_path):
    """include/dkb.h must compile as C11 and a C program must link against libdkb.so and
    call the host-side entry points (the boundary is a C ABI, not a C++ one)."""
    import subprocess
    from denovo_kmer_b200 import _lib
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "dkb.h"
int main(void) {
  uint64_t f = 0;
  if (dkb_abi_version() != DKB_ABI_VERSION) return 1;
  if (dkb_kmer_encode("ACGTT", 5, &f) != DKB_OK || f != 111) return 2;
  if (dkb_kmer_canonical(f, 5) != 27 || dkb_kmer_revcomp(f, 5) != 27) return 3;
  const uint8_t seq[] = "ACGTNACGTAC";
  const uint64_t off[3] = {0, 5, 11};
  uint32_t b[4], m[4];
  uint64_t n = 0;
  if (dkb_pack_reads(seq, NULL, off, 2, 0, b, m, &n) != DKB_OK || n != 13) return 4;
  if (m[0] != 0xFCFu) return 5;
  dkb_ctx *ctx = NULL;
  int rc = dkb_ctx_create(0, 40, &ctx);
  if (rc != DKB_EINVAL || ctx != NULL) return 6;
  dkb_thresholds t = {3, 2, 0, 1};
  dkb_tuning tu = {0, 0, 0, 0};
  (void)t; (void)tu;
  printf("%s\n", dkb_strerror(DKB_ENODEV));
  return 0;
}
''')
    exe = tmp_path / "abi"
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe),
                           "-L", libdir, "-ldkb", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "no usable CUDA device" in out.stdout


def _build_cpp_trio(tmp_path):
    import subprocess
    from denovo_kmer_b200 import _lib
    exe = tmp_path / "cpp_trio"
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(HERE, "cpp_trio.cpp"), "-o", str(exe), "-L", libdir, "-ldkb",
                           f"-Wl,-rpath,{libdir}"])
    return subprocess.run([str(exe)], capture_output=True, text=True)


def test_cpp_host_layer_builds_and_refuses_without_gpu(dkb, tmp_path):
    """include/dkb.hpp (RAII C++ layer) compiles, the host-side pieces give the expected
    entry and stream sizes, and without a GPU the context throws DKB_ENODEV."""
    import torch
    out = _build_cpp_trio(tmp_path)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.strip() == ("gpu-ok" if torch.cuda.is_available() else "no-gpu")


@pytest.mark.gpu
def test_cpp_host_layer_on_gpu(dkb, tmp_path):
    out = _build_cpp_trio(tmp_path)
    assert out.returncode == 0 and out.stdout.strip() == "gpu-ok", (out.returncode, out.stdout, out.stderr)


def test_rust_sys_crate_declares_every_symbol():
    """bindings/rust/denovo-kmer-gpu-sys mirrors include/dkb.h (no Rust toolchain here: the
    crate is checked textually - every exported function, every error code, the struct fields)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "dkb.h")).read()
    rs = open(os.path.join(root, "bindings", "rust", "denovo-kmer-gpu-sys", "src", "lib.rs")).read()
    c_funcs = set(re.findall(r"\b(dkb_[a-z_]+)\(", hdr))
    rs_funcs = set(re.findall(r"pub fn (dkb_[a-z_]+)\(", rs))
    assert c_funcs == rs_funcs, (c_funcs ^ rs_funcs)
    for name, val in re.findall(r"(DKB_E[A-Z]+|DKB_OK) = (\d+)", hdr):
        assert re.search(rf"pub const {name}: c_int = {val};", rs), name
    stats = re.search(r"typedef struct dkb_stats \{(.*?)\} dkb_stats;", hdr, re.S).group(1)
    stats = re.sub(r"/\*.*?\*/", "", stats, flags=re.S)
    c_fields = [f.strip() for decl in stats.split(";") for f in decl.split(",")]
    c_fields = [re.sub(r"^(uint64_t|uint32_t|double|float)\s+", "", f) for f in c_fields if f.strip()]
    rs_stats = re.search(r"pub struct DkbStats \{(.*?)\n\}", rs, re.S).group(1)
    rs_fields = re.findall(r"pub (\w+):", rs_stats)
    assert c_fields == rs_fields, (c_fields, rs_fields)


def test_rust_safe_crate_uses_only_declared_symbols():
    """bindings/rust/denovo-kmer-gpu (RAII Counter, Result errors, pinned BatchPacker, the
    INTEGRATION.md loop as examples/trio.rs) cannot be compiled here; check that every
    sys::dkb_* call it makes exists in the -sys crate and passes the declared number of
    arguments, and that every safe method the example uses exists."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sysrs = open(os.path.join(root, "bindings", "rust", "denovo-kmer-gpu-sys", "src", "lib.rs")).read()
    safe = open(os.path.join(root, "bindings", "rust", "denovo-kmer-gpu", "src", "lib.rs")).read()
    example = open(os.path.join(root, "bindings", "rust", "denovo-kmer-gpu", "examples", "trio.rs")).read()
    arity = {}
    for name, args in re.findall(r"pub fn (dkb_[a-z_]+)\((.*?)\)\s*(?:->|;)", sysrs, re.S):
        arity[name] = 0 if not args.strip() else args.count(":")

    def call_args(text, start):
        depth, i, n, seen = 0, start, 0, False
        while True:
            ch = text[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    return n + (1 if seen else 0)
            elif ch == "," and depth == 1:
                n += 1
                seen = False
            elif not ch.isspace() and depth >= 1 and not (depth == 1 and ch == "("):
                seen = True
            i += 1

    used = set()
    for m in re.finditer(r"sys::(dkb_[a-z_]+)\(", safe):
        name = m.group(1)
        used.add(name)
        assert name in arity, f"{name} is not declared in the -sys crate"
        assert call_args(safe, m.end() - 1) == arity[name], name
    assert {"dkb_ctx_create", "dkb_ctx_destroy", "dkb_table_build", "dkb_batch_submit", "dkb_pack_reads_fmt",
            "dkb_host_alloc", "dkb_finalise", "dkb_results_fetch", "dkb_comm_init",
            "dkb_reduce_push"} <= used
    methods = set(re.findall(r"pub fn (\w+)", safe))
    for m in re.findall(r"(?:kc|packer)\.(\w+)\(", example):
        assert m in methods, f"examples/trio.rs calls {m}() which the safe crate does not define"
    assert "impl Drop for Counter" in safe and "Result<" in safe


def test_zero_list_round_trip(dkb, orc):
    """dkb_mask_to_zero_list against the plain-Python reading of the format: random flags of
    several densities, runs of >= 255 usable positions (escape bytes), dense-zero blocks (stored
    as plain bits), lengths that end inside a block / a word, a packed real trio."""
    import numpy as np
    from denovo_kmer_b200 import synth
    rng = np.random.default_rng(11)

    def check(flags):
        n = len(flags)
        mw = (n + 127) // 128 * 4
        bits = np.zeros(mw * 32, dtype=np.uint8)
        bits[:n] = flags
        bits[n:] = rng.integers(0, 2, size=len(bits) - n)  # padding must not matter
        mask1 = np.packbits(bits, bitorder="little").view(np.uint32)
        zoff, zbytes = dkb.mask_to_zero_list(mask1, n)
        assert len(zoff) == (n + 2047) // 2048 + 1 and int(zoff[-1]) == len(zbytes)
        back = orc.expand_zero_list(zoff, zbytes, n)
        assert np.array_equal(back, flags), n
        # the sizing form (zbytes == NULL) reports the same offsets and length
        import ctypes as C
        from denovo_kmer_b200 import _lib
        zoff2 = np.zeros_like(zoff)
        used = C.c_size_t(0)
        rc = _lib.lib().dkb_mask_to_zero_list(mask1.ctypes.data_as(C.POINTER(C.c_uint32)), n,
                                              zoff2.ctypes.data_as(C.POINTER(C.c_uint32)), None, 0, C.byref(used))
        assert rc == 0 and used.value == len(zbytes) and np.array_equal(zoff2, zoff)
        return len(zbytes)

    for n in (1, 31, 2047, 2048, 2049, 5000, 70_001):
        for dens in (0.0, 0.002, 0.04, 0.5, 1.0):
            check((rng.random(n) >= dens).astype(np.uint8))
    f = np.ones(10_000, dtype=np.uint8)
    f[[0, 254, 255, 256, 600, 2047, 2048, 9_999]] = 0      # gaps of 253, 0, 0, 343 (escape), ...
    check(f)
    f = np.ones(6000, dtype=np.uint8)
    f[2048:4096] = 0                                         # a whole block of zeros: plain bits
    assert check(f) == 256
    trio = synth.make_trio_host(30_000, 10, 5, 31, seed=9)
    st = dkb.pack_reads(*trio.reads[0], 20)
    used = check(np.unpackbits(st.mask1.view(np.uint8), bitorder="little")[: st.n_positions])
    assert used < st.n_positions / 8 * 0.5  # less than half of the dense flags' bytes
