cd /root/repo
for args in "--variants 1000" "--variants 100 --genome-mb 16" "--variants 2500"; do
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$args | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'frac %.3f' % d['roofline']['frac'])"
done
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --variants 1000"
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_scan -s 9 -c 1 --csv --log-file gpurun_out/t30.csv $CMD > /dev/null 2>&1; grep -E "dram__bytes_read|duration|hit_rate" gpurun_out/t30.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
