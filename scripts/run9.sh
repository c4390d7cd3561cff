cd /root/repo
mkdir -p gpurun_out
for t in 14,4,2 15,2,2 15,2,3 15,1,2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --tuning $t 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$t', 'value %.3f T/s' % (d['value']/1e12), 'ms/step %.3f' % d['ms_per_step'], 'scan_ms %.3f' % d['roofline']['launch_ms'], 'frac %.4f' % d['roofline']['frac'], d['config']['seeds'], d['config']['denovo_calls'])"
done 2>&1 | tee gpurun_out/bench_tunings.log
