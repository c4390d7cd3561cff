cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for args in "" "--variants 5000" "--variants 2500" "--variants 1000" "--variants 100 --genome-mb 1" "--k 25" "--k 21" "--k 15" "--genome-mb 128 --variants 4000 --table-variants 100000"; do
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$args | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'seeds', d['config']['seeds'], 'frac %.3f' % d['roofline']['frac'])"
done
