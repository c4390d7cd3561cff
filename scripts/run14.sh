cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'])"
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --variants 100 --genome-mb 16 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('100 variants/16Mb: value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], 'frac %.3f' % d['roofline']['frac'])"
