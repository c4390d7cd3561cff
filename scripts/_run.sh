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
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for v in 2500 5000 10000 20000; do b --variants $v; done
) > gpurun_out/run.log 2>&1
cat gpurun_out/run.log
