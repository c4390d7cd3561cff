cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; cat gpurun_out/bench_ours.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 9 -c 3 -o gpurun_out/prof_scan_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
