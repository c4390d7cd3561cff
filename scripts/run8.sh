cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/parity8.log
timeout 600 python scripts/quick_scan_bench.py 10000 1e9 31 15,1,2 15,2,2 14,4,1 14,4,2 2>&1 | grep -v "hints\": false" | grep -v "^   prof" | tee gpurun_out/quick8.log
