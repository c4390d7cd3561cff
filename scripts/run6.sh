cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/parity6.log
timeout 600 python scripts/quick_scan_bench.py 10000 1e9 31 15,1,1 15,1,2 15,2,1 15,2,2 14,4,1 14,4,2 2>&1 | grep -v "hints\": false" | grep -v "^   prof" | tee gpurun_out/quick6.log
CMD="python scripts/quick_scan_bench.py 10000 1e9 31 14,4,2"
$CMD > gpurun_out/plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 2 -c 1 -o gpurun_out/prof_scan_d4b $CMD > gpurun_out/ncu6.log 2>&1
echo rc=$?
