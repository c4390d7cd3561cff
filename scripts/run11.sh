cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain11.log 2>&1 && tail -1 gpurun_out/plain11.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value %.3f T/s' % (d['value']/1e12), 'scan_ms %.3f' % d['roofline']['launch_ms'], d['config']['tuning_seedlen_stride_hashes_filtermode'])" && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_scan -s 9 -c 3 --csv --log-file gpurun_out/traffic11.csv $CMD > gpurun_out/ncu11.log 2>&1
echo rc=$?; grep -E "dram__|hit_rate|duration" gpurun_out/traffic11.csv | cut -d, -f5,13-15 | head -12
