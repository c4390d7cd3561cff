# Work-unit sizing for short launches (DKB_UNIT_GROUPS): parity of the macro path at every unit
# size, then A/B on configs[0], two mid sizes and configs[1] (full macro tiles forced vs the launch's choice).
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -3
run() { env $2 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_s.json 2> gpurun_out/t_s.err
  python -c "
import json
d=json.load(open('gpurun_out/t_s.json')); print('$1 | $3 |', round(d['value']/1e12,3), 'Tb/s step_ms', round(d['ms_per_step'],4), 'launch_ms', round(d['roofline']['launch_ms'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'])"; }
for a in "--genome-mb 1 --variants 100" "--genome-mb 4 --variants 400" "--genome-mb 16 --variants 2500" ""; do
  run full "DKB_UNIT_GROUPS=4" "$a"; run auto "" "$a"
done
run g1 "DKB_UNIT_GROUPS=1" "--genome-mb 1 --variants 100"
run g2 "DKB_UNIT_GROUPS=2" "--genome-mb 1 --variants 100"
