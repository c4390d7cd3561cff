"""Turn gpurun_out/launches_<R>.csv + prof_scan_<R>.ncu-rep into the tracked summaries
under profiles/: <R>_launches.csv (our kernels, per launch), <R>_scan_ncu.txt (raw metrics,
stalls, per-stage instruction split) and scan_traffic.json (DRAM bytes per scan launch,
read by bench.py for roofline.traffic).  Usage: python scripts/make_profile_summary.py r1"""
import collections, csv, io, json, os, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list ----
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{R}.csv"))) if len(r) > 10]
hdr = None
launches = []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    d = dict(zip(hdr, r))
    launches.append((int(d["ID"]), d["Kernel Name"].split("(")[0].replace("void ", ""), d["Grid Size"],
                     d["Block Size"], float(d["Metric Value"].replace(",", ""))))
with open(os.path.join(out_dir, f"{R}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ on: python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs\n")
    f.write("# per-launch times are cold-cache and serialised: compare shares, not absolutes\n")
    f.write("id,kernel,grid,block,duration_ns\n")
    for l in launches:
        f.write(f'{l[0]},"{l[1]}","{l[2]}","{l[3]}",{l[4]:.0f}\n')
agg = collections.defaultdict(lambda: [0, 0.0])
for l in launches:
    agg[l[1]][0] += 1
    agg[l[1]][1] += l[4]
tot = sum(v[1] for v in agg.values())

# ---- full capture ----
rep = os.path.join(ROOT, "gpurun_out", f"prof_scan_{R}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units = rr[0], rr[1]
recs = [dict(zip(h, r)) for r in rr[2:]]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]

def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None

def to_bytes(d, key):
    v = num(d[key]); u = units[h.index(key)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

want += ["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
         "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
         "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
         "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum"]
lines = [f"ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 2 on: python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs",
         "(one launch scans the three samples of the trio: 1.4496 GB of 2-bit stream, 5.76 Gbases)",
         f"kernel: {recs[0].get('Kernel Name')} grid {recs[0].get('Grid Size')} block {recs[0].get('Block Size')}", ""]
for i, d in enumerate(recs):
    lines.append(f"launch {i}:")
    for w in want:
        if w in d:
            lines.append(f"  {w:72s} {d[w]} {units[h.index(w)]}")
    st = [(x, num(d[x])) for x in h if "issue_stalled" in x and x.endswith("per_issue_active.ratio") and num(d[x]) is not None]
    for x, v in sorted(st, key=lambda t: -t[1])[:7]:
        lines.append(f"  stall/issue {x.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
    lines.append("")
dram = [to_bytes(d, "dram__bytes_read.sum") + to_bytes(d, "dram__bytes_write.sum") for d in recs]
traffic = {"dram_bytes_per_launch": sum(dram) / len(dram), "launches": len(dram), "round": R,
           "source": f"profiles/{R}_scan_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, cold L2 per launch; "
                     "a launch = the three samples of a step)"}
wrep = os.path.join(ROOT, "gpurun_out", f"prof_scan_wgs_{R}.ncu-rep")
if os.path.exists(wrep):  # the WGS-shard capture: one launch
    wraw = subprocess.run(["ncu", "-i", wrep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    wr = list(csv.reader(io.StringIO(wraw)))
    wh, wu = wr[0], wr[1]
    wlines = ["ncu --set full --clock-control none -k regex:k_scan -s 3 -c 1 on: python bench.py --steps 2 --warmup 3 --no-e2e "
              "--no-cpu-baseline --genome-mb 128 --variants 4000 --table-variants 100000",
              "(one shard of BASELINE.json configs[2]: 2.8992 GB of 2-bit stream per launch, 11.52 Gbases, 100 000-candidate table)", ""]
    for d in [dict(zip(wh, r)) for r in wr[2:]]:
        for w in want:
            if w in d:
                wlines.append(f"  {w:72s} {d[w]} {wu[wh.index(w)]}")
        st = [(x, num(d[x])) for x in wh if "issue_stalled" in x and x.endswith("per_issue_active.ratio") and num(d[x]) is not None]
        for x, v in sorted(st, key=lambda t: -t[1])[:7]:
            wlines.append(f"  stall/issue {x.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
        def wb(key):
            return num(d[key]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(wu[wh.index(key)], 1)
        traffic["wgs_dram_bytes_per_launch"] = wb("dram__bytes_read.sum") + wb("dram__bytes_write.sum")
    open(os.path.join(out_dir, f"{R}_scan_wgs_ncu.txt"), "w").write("\n".join(wlines) + "\n")
json.dump(traffic, open(os.path.join(out_dir, "scan_traffic.json"), "w"))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = hd = None
a2 = collections.defaultdict(lambda: [0, 0])
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hd = r
        continue
    if hd and len(r) == len(hd):
        d = dict(zip(hd, r))
        try:
            ln = int(d["Line No"]); inst = int(d["Instructions Executed"] or 0); smp = int(d["# Samples"] or 0)
        except ValueError:
            continue
        a2[(cur, ln)][0] += inst
        a2[(cur, ln)][1] += smp
ti = sum(v[0] for v in a2.values()) or 1
ts = sum(v[1] for v in a2.values()) or 1
code = open(os.path.join(ROOT, "denovo_kmer_b200", "csrc", "dkb_scan.cuh")).read().split("\n")
PAT = r"^\s*__device__ __forceinline__ (?:static )?[\w ]+? (\w+)\(|\) (k_scan)\(const ScanParams"
marks = []
for i, l in enumerate(code):
    m = re.search(PAT, l)
    if m:
        marks.append((i + 1, m.group(1) or m.group(2)))
bounds = [m[0] for m in marks] + [10 ** 6]
lines.append(f"per-stage split over the {len(recs)} captured launches (source page, -lineinfo); instructions are warp-level:")
for (ln, name), (a, b) in zip(marks, zip(bounds[:-1], bounds[1:])):
    i = sum(v[0] for (f, l), v in a2.items() if f == "dkb_scan.cuh" and a <= l < b)
    s = sum(v[1] for (f, l), v in a2.items() if f == "dkb_scan.cuh" and a <= l < b)
    lines.append(f"  {name:16s} instructions {100 * i / ti:5.1f}%   stall samples {100 * s / ts:5.1f}%")
i = sum(v[0] for (f, l), v in a2.items() if f != "dkb_scan.cuh")
s = sum(v[1] for (f, l), v in a2.items() if f != "dkb_scan.cuh")
lines.append(f"  {'(intrinsics, helpers)':16s} instructions {100 * i / ti:5.1f}%   stall samples {100 * s / ts:5.1f}%")
lines.append("")
lines.append("launch list shares (profiles/%s_launches.csv):" % R)
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"  {v[0]:3d} x {k:28s} {v[1] / 1e3:10.1f} us  {100 * v[1] / tot:6.2f}%")
open(os.path.join(out_dir, f"{R}_scan_ncu.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-24:]))
