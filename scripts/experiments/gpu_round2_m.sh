cd /root/repo
mkdir -p gpurun_out
prof() { # name, env, args: DRAM bytes and L2 hit rate of one scan launch
  env $2 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum --clock-control none -k regex:k_scan -s 3 -c 1 --csv --log-file gpurun_out/h_$1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $3 > /dev/null 2>&1
  python -c "
import csv
rows=[r for r in csv.reader(open('gpurun_out/h_$1.csv')) if len(r)>10]
h=rows[0]; print('$1', [(r[h.index('Metric Name')][:44], r[h.index('Metric Value')]) for r in rows[1:]])"
}
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs > /dev/null 2>&1 && {
prof f1mb "DKB_L2_FILTER_MAX_WORDS=300000" ""
prof f2mb "DKB_L2_FILTER_MAX_WORDS=600000" ""
prof f5mb "X=1" ""
prof f32mb_cap_irrelevant "DKB_L2_FILTER_MAX_WORDS=8388608" ""
prof nohint "DKB_LIBRARY=ab/libdkb_fp1.so DKB_TUNING=0,0,2,0" ""
prof nopre "DKB_PREFILTER_WORDS=0" ""
}
