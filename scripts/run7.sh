cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/quick_scan_bench.py 10000 1e9 31 15,2,2 15,2,3 15,2,4 14,4,2 14,4,3 14,4,4 2>&1 | grep -v "hints\": false"  | tee gpurun_out/quick7.log
