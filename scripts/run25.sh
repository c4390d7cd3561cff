cd /root/repo
for args in "" "--genome-mb 128 --variants 4000 --table-variants 100000" "--variants 50000 --depth 100 --indel-frac 0.5 --genome-mb 32" "--variants 100000" "--variants 1000" "--k 21" "--k 25" "--k 15" "--variants 100 --genome-mb 1"; do
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$args | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'seeds', d['config']['seeds'], 'frac %.3f' % d['roofline']['frac'])"
done
