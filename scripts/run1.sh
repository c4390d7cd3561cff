cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/parity1.log
timeout 600 python scripts/quick_scan_bench.py 10000 1e9 31 2>&1 | tee gpurun_out/quick1.log
