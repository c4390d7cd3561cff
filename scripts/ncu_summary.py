"""Summarise an .ncu-rep: key raw metrics, top stall reasons, per-source-line hot spots."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_sector_hit_rate.pct','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('---',d.get('Kernel Name'), d.get('Grid Size'), d.get('Block Size'))
    for w in want:
        if w in d: print(f"  {w:75s} {d[w]} {units[hdr.index(w)]}")
    stall=[(h,d[h]) for h in hdr if 'issue_stalled' in h and h.endswith('_per_warp_active.pct')]
    for h,v in sorted(stall,key=lambda x:-float(x[1].replace(',','') or 0))[:8]: print('   stall',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_warp_active.pct',''),v)
    break
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
cur=None; hdr=None
agg=collections.defaultdict(lambda:[0,0]); text={}
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=='Line No': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: ln=int(d['Line No']); inst=int(d['Instructions Executed'] or 0); samp=int(d['# Samples'] or 0)
        except: continue
        agg[(cur,ln)][0]+=inst; agg[(cur,ln)][1]+=samp
tot_i=sum(v[0] for v in agg.values()) or 1; tot_s=sum(v[1] for v in agg.values()) or 1
print('total warp-inst',tot_i,'samples',tot_s)
for (f,ln),v in sorted(agg.items(), key=lambda x:-x[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print(f"{f}:{ln:4d} inst {100*v[0]/tot_i:5.1f}%  samples {100*v[1]/tot_s:5.1f}%")
