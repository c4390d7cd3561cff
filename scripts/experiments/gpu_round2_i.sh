cd /root/repo
mkdir -p gpurun_out
W="--genome-mb 128 --variants 4000 --table-variants 100000"
prof() { # name, env, args: DRAM bytes and L2 hit rate of one scan launch
  env $2 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum --clock-control none -k regex:k_scan -s 3 -c 1 --csv --log-file gpurun_out/h_$1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $3 > /dev/null 2>&1
  python -c "
import csv
rows=[r for r in csv.reader(open('gpurun_out/h_$1.csv')) if len(r)>10]
h=rows[0]; print('$1', [(r[h.index('Metric Name')][:40], r[h.index('Metric Value')]) for r in rows[1:]])"
}
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $W > /dev/null 2>&1 && {
prof wgs_nohits "DKB_LIBRARY=ab/libdkb_x7.so DKB_TUNING=0,0,2,0" "$W"

prof c1_nohits "DKB_LIBRARY=ab/libdkb_x7.so DKB_TUNING=0,0,2,0" ""

}
