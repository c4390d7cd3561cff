# k = 15 (configs[4]'s shortest k) under ncu --set full: what bounds the stride-4 kernel; then another soak range.
cd /root/repo
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs --k 15"
$CMD > gpurun_out/plain_k15.log 2>&1; echo "plain rc=$?"; cut -c1-400 gpurun_out/plain_k15.log | tail -1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/prof_scan_k15_r2 $CMD > gpurun_out/ncu_full_k15.log 2>&1
echo "ncu rc=$?"
DKB_FUZZ_BASE=20000 DKB_FUZZ_CASES=500 timeout 150 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
