cd /root/repo
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --variants 1000"
$CMD > gpurun_out/plain29.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 9 -c 1 -o gpurun_out/prof_scan_d16b $CMD > gpurun_out/ncu29.log 2>&1
echo rc=$?
