cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for t in "" 15,8,2 15,8,1 14,8,2 15,16,2 15,16,1; do
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${t:+--tuning $t} 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$t | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'seeds', d['config']['seeds'])"
done
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --variants 100 --genome-mb 16 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('100 variants | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'seeds', d['config']['seeds'])"
