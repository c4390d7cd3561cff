cd /root/repo
CMD="python scripts/tuning_sweep.py 31 250 16"
$CMD > gpurun_out/plain27.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 2 -c 1 -o gpurun_out/prof_scan_d16 $CMD > gpurun_out/ncu27.log 2>&1
echo rc=$?
