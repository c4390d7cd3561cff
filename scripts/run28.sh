cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/tuning_sweep.py 31 250 16 15,8,2 15,16,2 15,16,1 2>&1 | grep -E "^k=|invalid"
python scripts/tuning_sweep.py 31 2500 16 14,4,2 14,8,2 15,16,2 2>&1 | grep -E "^k=|invalid"
for args in "" "--variants 1000" "--variants 100 --genome-mb 16"; do
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$args | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'frac %.3f' % d['roofline']['frac'])"
done
