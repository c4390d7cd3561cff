#!/usr/bin/env python
"""bench.py — read bases k-mer-probed per second on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic trio shard: counts reset, the
three samples' packed read streams scanned (kernel 2) against the spanning-k-mer table,
(N>1: ONE NCCL sum-allreduce of the per-entry counters, run on a side stream under the next
step's scans), kernel 3 (per-variant reduce + de novo thresholds).  Default workload = BASELINE.json configs[1]: synthetic 30x trio,
chr20-scale (64 Mb), 10k candidate DNMs, k=31, 150 bp reads — per GPU (weak scaling:
every rank holds its own 64 Mb-scale shard of reads, the table is replicated).

  value         whole-job read bases / s, inputs resident in HBM (device-timed)
  e2e           same metric through the C ABI with HOST (pinned) buffers: H2D copies of
                every batch and the D2H read of the results inside the timed region
  roofline      scan kernel only: streamed bytes per launch / mean CUDA-event launch time,
                against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle port (oracle/, all host threads) on a bounded sample of the
                same workload, rank 0 at N=1 only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "read_bases_kmer_probed_per_sec"
UNIT = "bases/s"
READ_LEN = 150
THRESHOLDS = (3, 2, 0, 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=64.0, help="region size per GPU, Mb")
    ap.add_argument("--depth", type=float, default=30.0)
    ap.add_argument("--variants", type=int, default=10000, help="candidates inside this GPU's region")
    ap.add_argument("--table-variants", type=int, default=0,
                    help="total candidates in the replicated table (>= --variants): the extra ones lie "
                         "in other regions of a larger genome, as for one GPU's shard of a WGS job")
    ap.add_argument("--indel-frac", type=float, default=0.0)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tuning", default="", help="seed_len,stride,bloom_hashes (default auto)")
    return ap.parse_args()


def workload_name(a):
    extra = f" of {a.table_variants} in the table" if a.table_variants > a.variants else ""
    return (f"synthetic {a.depth:g}x trio, {a.genome_mb:g} Mb region per GPU, {READ_LEN} bp reads, "
            f"{a.variants} candidate DNMs{extra}, k={a.k}")


# ------------------------------------------------------------------------------------
def build_table_inputs(a):
    """Genome, variants and spanning k-mer entries — identical on every rank."""
    import denovo_kmer_b200 as dkb
    from denovo_kmer_b200 import synth
    glen = int(a.genome_mb * 1e6)
    genome = synth.make_genome(glen, seed=1)
    variants = synth.plant_variants(genome, a.variants, a.k, seed=2, indel_frac=a.indel_frac)
    tuples = synth.Trio(a.k, genome, variants).variant_tuples()
    extra = max(0, a.table_variants - a.variants)
    if extra:  # candidates of other regions: same model, sequence this shard's reads never contain
        other = synth.make_genome(max(int(extra * 400), 1_000_000), seed=7)
        tuples += synth.Trio(a.k, other, synth.plant_variants(other, extra, a.k, seed=8,
                                                              indel_frac=a.indel_frac)).variant_tuples()
    entries = dkb.variant_kmers(tuples, a.k)
    return genome, variants, entries


def cpu_sample(genome, variants, a, region=1_000_000):
    """Bounded CPU sample: BASELINE.json configs[0] shape (1 Mb region, same depth/model)
    cut from the same genome, with the variants that fall inside it."""
    from denovo_kmer_b200 import synth
    region = min(region, len(genome))
    g = genome[:region]
    vs = [v for v in variants if v.pos + 64 < region]
    child_alt = synth.apply_variants(g, vs)
    mother_alt = synth.apply_variants(g, [v for v in vs if v.inherited])
    n_reads = int(region * a.depth / READ_LEN)
    reads = [synth.sample_reads([g, child_alt], n_reads, READ_LEN, 910),
             synth.sample_reads([g, mother_alt], n_reads, READ_LEN, 911),
             synth.sample_reads([g, g], n_reads, READ_LEN, 912)]
    return reads, n_reads * READ_LEN * 3


def run_cpu(entries, reads, n_bases, k, seconds, passes=None):
    """Time the oracle port over the sample with every host thread; returns (bases/s, cores, passes)."""
    import oracle
    ks = oracle.KmerSet(entries.keys, entries.variant, entries.allele)
    cores = os.cpu_count() or 1
    counts = np.zeros((3, len(entries)), dtype=np.uint64)
    t0 = time.perf_counter()
    done = 0
    while True:
        for smp, (seq, qual, off) in enumerate(reads):
            ks.count_reads(seq, qual, off, k, 20, counts=counts[smp], threads=cores)
        done += 1
        el = time.perf_counter() - t0
        if (passes is not None and done >= passes) or (passes is None and el >= seconds):
            break
    return n_bases * done / el, cores, done, el


def run_reference(a):
    """--impl reference: the reference's CPU path (oracle port; the Rust original is neither
    mounted nor buildable here) on the box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    genome, variants, entries = build_table_inputs(a)
    reads, n_bases = cpu_sample(genome, variants, a)
    for _ in range(min(a.warmup, 1)):
        run_cpu(entries, reads, n_bases, a.k, 0, passes=1)
    rate, cores, done, el = run_cpu(entries, reads, n_bases, a.k, 0, passes=max(1, a.steps))
    sample = (f"{done} steps x ({a.depth:g}x trio over a 1 Mb region of the workload genome = "
              f"{n_bases / 1e6:.0f} Mbases) against the full {len(entries)}-entry table")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": el / done * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "reference": "CPU oracle port (oracle/dnk_oracle.c, OpenMP); "
                   "jlanej/denovo_kmer's Rust source is not mounted and no Rust toolchain exists here"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_ours(a):
    import torch
    import torch.distributed as tdist
    import denovo_kmer_b200 as dkb
    from denovo_kmer_b200 import dist, synth

    rank, world, local = dist.env_rank_world()
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG=VERSION/INFO prints to stdout; rank 0's stdout must be the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        tdist.init_process_group("nccl", device_id=dev)

    genome, variants, entries = build_table_inputs(a)
    lut = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    g_codes = torch.from_numpy(lut[genome]).to(dev)
    child_alt = torch.from_numpy(lut[synth.apply_variants(genome, variants)]).to(dev)
    mother_alt = torch.from_numpy(
        lut[synth.apply_variants(genome, [v for v in variants if v.inherited])]).to(dev)
    n_reads = int(len(genome) * a.depth / READ_LEN) // 128 * 128
    haps = [[g_codes, child_alt], [g_codes, mother_alt], [g_codes, g_codes]]
    streams = [synth.make_sample_device(haps[s], n_reads, READ_LEN, 1000 + 10 * rank + s, dev)
               for s in range(3)]
    del g_codes, child_alt, mother_alt
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    bases_per_step = sum(st[3] for st in streams)
    stream_bytes = sum(st[0].numel() * 4 for st in streams)
    mask_bytes = sum(st[1].numel() * 4 for st in streams)

    tuning = tuple(int(x) for x in a.tuning.split(",")) if a.tuning else None
    kc = dkb.KmerCounter(a.k, device=local, tuning=tuning)
    kc.build_table(entries)
    counts_t = dist.counts_tensor(kc)
    ext = torch.cuda.ExternalStream(kc.scan_stream(), device=dev)

    # one launch per <= 2^31 positions, cut after a multiple of 128 reads (word-aligned)
    max_reads = (1 << 31) // (READ_LEN + 1) // 128 * 128

    def submit_resident(s, b2, m1, n_pos):
        per = max_reads * (READ_LEN + 1)
        for p0 in range(0, n_pos, per):
            kc.submit_device(b2.data_ptr() + p0 // 4, m1.data_ptr() + p0 // 8, min(per, n_pos - p0), s)

    # N > 1: every step's counters are summed over ranks (one NCCL allreduce) and finalised
    # from the sum; the allreduce of step i runs on a side stream under the scans of step
    # i + 1 and its finalise is queued behind them (dist.CountPipeline).  The timed region
    # ends after the last step's allreduce and finalise.
    pipe = dist.CountPipeline(kc, THRESHOLDS) if world > 1 else None

    def step(last=True):
        kc.reset_counts()  # (queued behind the previous step's copy on the scan stream)
        for s, (b2, m1, n_pos, _) in enumerate(streams):
            submit_resident(s, b2, m1, n_pos)
        if pipe is None:
            kc.finalise_launch(THRESHOLDS)
        else:
            pipe.finalise(pipe.push())
            if last:
                pipe.flush()

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:  # clocks are sampled from the warm-up steps to the end of the timed region
        sampler.start()
        time.sleep(0.1)
    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    s0 = kc.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(ext)
    for i in range(a.steps):
        step(last=i == a.steps - 1)
    ev1.record(ext)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    s1 = kc.stats()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t_ms, op=tdist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = bases_per_step * world * a.steps / (ms_max * 1e-3)
    hits, distinct, n_kmers, calls = kc.results()
    ref_counts = kc.entry_counts().copy()

    # roofline of the scan kernel: bytes it must stream per launch / mean launch time
    launches = s1["scan_launches_timed"] - s0["scan_launches_timed"]
    scan_ms = (s1["scan_ms_total"] - s0["scan_ms_total"]) / max(launches, 1)
    bytes_per_launch = stream_bytes / max(launches / max(a.steps, 1), 1)
    peak, peak_src = measured_peak_gbs()
    achieved = bytes_per_launch / (scan_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass

    # ---- e2e: host pinned buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not a.no_e2e:
        host = []
        for (b2, m1, n_pos, _) in streams:
            hb = torch.empty(b2.numel(), dtype=torch.int32).pin_memory()
            hm = torch.empty(m1.numel(), dtype=torch.int32).pin_memory()
            hb.copy_(b2)
            hm.copy_(m1)
            host.append((hb, hm, n_pos))
        del streams
        torch.cuda.empty_cache()
        reads_per_batch = 2_097_152  # multiple of 128 reads -> batches start on word boundaries
        pos_per_batch = reads_per_batch * (READ_LEN + 1)

        def e2e_step():
            kc.reset_counts()
            for s, (hb, hm, n_pos) in enumerate(host):
                for p0 in range(0, n_pos, pos_per_batch):
                    n = min(pos_per_batch, n_pos - p0)
                    kc._ck(kc._L.dkb_batch_submit(kc._h, hb.data_ptr() + p0 // 4,
                                                  hm.data_ptr() + p0 // 8, n, s))
            if world > 1:
                with torch.cuda.stream(ext):
                    dist.allreduce_counts(counts_t)
            return kc.finalise(THRESHOLDS)  # kernel 3 + D2H of hits/distinct/n_kmers/calls

        e2e_step()
        barrier()
        ev0.record(ext)
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            res = e2e_step()
        ev1.record(ext)
        barrier()
        wall = time.perf_counter() - t0
        t_e = torch.tensor([max(ev0.elapsed_time(ev1) * 1e-3, wall)], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(t_e, op=tdist.ReduceOp.MAX)
        same = bool(np.array_equal(kc.entry_counts(), ref_counts)) if world == 1 else None
        d2h = sum(x.nbytes for x in res)
        e2e = {"value": bases_per_step * world * a.e2e_steps / float(t_e.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(stream_bytes + mask_bytes), "d2h_bytes_per_step": int(d2h),
               "steps": a.e2e_steps, "counts_equal_device_resident_run": same}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        reads, nb = cpu_sample(genome, variants, a)
        rate, cores, done, el = run_cpu(entries, reads, nb, a.k, a.cpu_seconds)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{done} passes over a {a.depth:g}x trio on a 1 Mb region of the workload genome "
                         f"({nb / 1e6:.0f} Mbases per pass, {el:.1f} s) against the full "
                         f"{len(entries)}-entry table; oracle/dnk_oracle.c with OpenMP"}

    if rank == 0:
        st = kc.stats()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": workload_name(a), "bases_per_step_per_gpu": int(bases_per_step),
                "l2": "inputs larger than L2 (%.2f GB of packed streams per step per GPU)" % (
                    (stream_bytes + mask_bytes) / 1e9),
                "table_entries": int(st["n_entries"]), "seeds": int(st["n_seeds"]),
                "tuning_seedlen_stride_hashes_filtermode": list(kc.tuning()),
                "prefilter_words": int(st.get("prefilter_words", 0)),
                "denovo_calls": int((calls & 1).sum()), "variants": int(len(calls)),
                "collective": "1 NCCL allreduce(sum) of %d uint32 per step" % counts_t.numel() if world > 1 else "none",
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": "dkb::k_scan",
                         "peak_source": peak_src, "launch_ms": scan_ms,
                         "bytes_per_launch": bytes_per_launch,
                         "note": "bytes = 2-bit base stream only (0.2517 B/read base incl. separators); "
                                 "the 1-bit mask stream is read only for verified seed hits"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int((s1["scan_launches"] - s0["scan_launches"]) + 2 * a.steps),
            "clocks": clocks,
        }
        print(json.dumps(out))
    kc.close()
    if world > 1:
        tdist.destroy_process_group()


def main():
    a = parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that chat on fd 1 (torch
    # prints "NCCL version ..." there on the first collective) are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global print

    def print(*args, **kw):  # noqa: A001 - the module's result lines
        if kw.get("file") is not None:
            import builtins
            return builtins.print(*args, **kw)
        os.write(json_fd, (" ".join(str(x) for x in args) + "\n").encode())

    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
