# Round profile: launch list of our kernels + full ncu capture of the scan kernel, both on
# the bench command (configs[1], one launch per trio).  Usage (under gpurun): bash scripts/profile_round.sh r2
R=${1:-r2}
cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs"
$CMD > gpurun_out/plain_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_launch_$R.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 2 -o gpurun_out/prof_scan_$R $CMD > gpurun_out/ncu_full_$R.log 2>&1
echo "ncu full rc=$?"
WCMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --genome-mb 128 --variants 4000 --table-variants 100000"
$WCMD > gpurun_out/plain3_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 3 -c 1 -o gpurun_out/prof_scan_wgs_$R $WCMD > gpurun_out/ncu_full_wgs_$R.log 2>&1
echo "ncu wgs rc=$?"
tail -1 gpurun_out/plain_$R.log | cut -c1-300
