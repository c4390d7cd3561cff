cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --variants 100 --genome-mb 16"
$CMD > gpurun_out/plain15.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 9 -c 1 -o gpurun_out/prof_scan_100v $CMD > gpurun_out/ncu15.log 2>&1
echo rc=$?
