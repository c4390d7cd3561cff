# Round-2 GPU call B: all GPU tests on the canonical-seed build, bench, canonical on/off over the workload table.
cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "ours rc=$?"; cat gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
export DKB_TUNING=0,0,2,0
for v in 250 1000 2500 5000 10000 20000; do
  for lib in default nocanon canonall; do
    L=ab/libdkb_$lib.so; [ $lib = default ] && L=denovo_kmer_b200/libdkb.so
    DKB_LIBRARY=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs --variants $v > gpurun_out/t_${lib}_$v.json 2> gpurun_out/t_${lib}_$v.err
    python -c "
import json
try:
    d=json.load(open('gpurun_out/t_${lib}_$v.json')); print('$lib', $v, round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['seeds'], d['config']['prefilter_words'])
except Exception as e: print('$lib', $v, 'FAILED', e)"
  done
done
for pw in 43776 51712; do
  DKB_PREFILTER_WORDS=$pw timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs > gpurun_out/t_pw$pw.json 2> gpurun_out/t_pw$pw.err
  python -c "
import json
d=json.load(open('gpurun_out/t_pw$pw.json')); print('prefilter $pw', round(d['value']/1e12,3), round(d['roofline']['frac'],4))"
done
DKB_LIBRARY=ab/libdkb_nocanon.so timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/t_nocanon_wgs.json 2> gpurun_out/t_nocanon_wgs.err
python -c "
import json
d=json.load(open('gpurun_out/t_nocanon_wgs.json')); print('nocanon wgs', round(d['wgs_shard']['value']/1e12,3), round(d['wgs_shard']['roofline']['frac'],4))"
unset DKB_TUNING
for c in "--genome-mb 1 --variants 100" "--k 15" "--k 21" "--k 25" "--depth 100 --variants 50000 --indel-frac 0.5"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $c > gpurun_out/t_shape.json 2> gpurun_out/t_shape.err
  python -c "
import json
d=json.load(open('gpurun_out/t_shape.json')); print('$c', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['seeds'], d['roofline']['launch_ms'])"
done
