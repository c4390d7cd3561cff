cd /root/repo
mkdir -p gpurun_out
CMD="python scripts/quick_scan_bench.py 10000 1e9 31 15,2,2"
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 2 -c 1 -o gpurun_out/prof_scan_d2 $CMD > gpurun_out/ncu4.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu4.log
