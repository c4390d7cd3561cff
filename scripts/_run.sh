cd /root/repo
mkdir -p gpurun_out
b() { timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 "$@" 2>gpurun_out/err.log | python -c "
import sys,json
s=sys.stdin.read()
try:
    d=json.loads(s); print('$*', 'T/s %.3f'%(d['value']/1e12), 'scan_ms %.4f'%d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['seeds'], 'calls', d['config']['denovo_calls'])
except Exception as e:
    print('$*', 'FAILED', open('gpurun_out/err.log').read()[-300:])
"; }
(
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for v in 250 1000 2500 5000 7000 10000 20000; do b --variants $v; done
b --variants 4000 --genome-mb 128 --table-variants 100000
b --variants 50000 --depth 100 --indel-frac 0.2
for k in 25 21 15; do b --variants 10000 --k $k; done
b --variants 100 --genome-mb 1
) > gpurun_out/run.log 2>&1
cat gpurun_out/run.log
