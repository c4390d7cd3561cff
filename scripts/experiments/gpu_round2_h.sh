cd /root/repo
mkdir -p gpurun_out
W="--genome-mb 128 --variants 4000 --table-variants 100000"
run() { # name, env, args
  env $2 timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/t_c.json')); print('$1 | $3 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['seeds'], round(d['roofline']['launch_ms'],4))
except Exception as e: print('$1 $3 FAILED', e)"
}
prof() { # name, env, args: DRAM bytes and L2 hit rate of one scan launch
  env $2 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__time_duration.sum --clock-control none -k regex:k_scan -s 3 -c 1 --csv --log-file gpurun_out/h_$1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $3 > /dev/null 2>&1
  python -c "
import csv
rows=[r for r in csv.reader(open('gpurun_out/h_$1.csv')) if len(r)>10]
h=rows[0]; print('$1', [(r[h.index('Metric Name')], r[h.index('Metric Value')], r[h.index('Metric Unit')]) for r in rows[1:]])"
}
for v in base tp1 fp1 tp1fp1; do
  L=ab/libdkb_$v.so; [ $v = base ] && L=denovo_kmer_b200/libdkb.so
  run $v "DKB_LIBRARY=$L DKB_TUNING=0,0,2,0" "$W"
  run $v "DKB_LIBRARY=$L DKB_TUNING=0,0,2,0" ""
done
run f4mb "DKB_L2_FILTER_MAX_WORDS=1048576" "$W"
run f8mb "DKB_L2_FILTER_MAX_WORDS=2097152" "$W"
run f32mb "DKB_L2_FILTER_MAX_WORDS=8388608" "$W"
run f48mb "DKB_L2_FILTER_MAX_WORDS=12000000" "$W"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $W > /dev/null 2>&1 && {
prof base "X=1" "$W"
prof tp1 "DKB_LIBRARY=ab/libdkb_tp1.so DKB_TUNING=0,0,2,0" "$W"
prof fp1 "DKB_LIBRARY=ab/libdkb_fp1.so DKB_TUNING=0,0,2,0" "$W"
prof f4mb "DKB_L2_FILTER_MAX_WORDS=1048576" "$W"
prof f32mb "DKB_L2_FILTER_MAX_WORDS=8388608" "$W"
}
