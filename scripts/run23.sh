cd /root/repo
for args in "--genome-mb 128 --variants 4000 --table-variants 100000" "--genome-mb 128 --variants 4000 --table-variants 100000 --tuning 15,1,2,1" "--genome-mb 128 --variants 4000 --table-variants 100000 --tuning 14,4,2,1" "--genome-mb 128 --variants 4000 --table-variants 100000 --tuning 15,8,2,2" "--genome-mb 128 --variants 4000 --table-variants 100000 --tuning 14,4,2,2"; do
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$args | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'seeds', d['config']['seeds'], 'entries', d['config']['table_entries'], 'calls', d['config']['denovo_calls'])"
done
