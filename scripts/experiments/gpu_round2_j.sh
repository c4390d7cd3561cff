cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
run() { # name, env, args
  env $2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/t_c.json')); print('$1 | $3 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['seeds'], round(d['roofline']['launch_ms'],4), round(d['ms_per_step'],4))
except Exception as e: print('$1 $3 FAILED', e)"
}
run lop3 "X=1" ""
run nh1 "DKB_TUNING=15,16,1,2" ""
run lop3 "X=1" "--variants 20000"
run nh1 "DKB_TUNING=15,16,1,2" "--variants 20000"
run lop3 "X=1" "--variants 5000"
run nh1 "DKB_TUNING=15,16,1,2" "--variants 5000"
run wgs "X=1" "--genome-mb 128 --variants 4000 --table-variants 100000"
run wgs32 "DKB_L2_FILTER_MAX_WORDS=8388608" "--genome-mb 128 --variants 4000 --table-variants 100000"
