cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain12.log 2>&1 && tail -1 gpurun_out/plain12.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'])" && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 9 -c 1 -o gpurun_out/prof_scan_r1b $CMD > gpurun_out/ncu12.log 2>&1
echo rc=$?
