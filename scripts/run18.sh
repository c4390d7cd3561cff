cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for args in "" "--k 15" "--k 21" "--k 25" "--variants 50000 --depth 100 --indel-frac 0.5 --genome-mb 32" "--variants 100000"; do
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $args 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$args | value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes'], 'seeds', d['config']['seeds'], 'entries', d['config']['table_entries'], 'calls', d['config']['denovo_calls'])"
done
