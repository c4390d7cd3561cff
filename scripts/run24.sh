cd /root/repo
for cap in 2097152 4194304 8388608; do
for t in 15,16,2,2 15,8,2,2; do
DKB_L2_FILTER_MAX_WORDS=$cap python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --genome-mb 128 --variants 4000 --table-variants 100000 --tuning $t 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('cap $cap $t | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'seeds', d['config']['seeds'])"
done; done
