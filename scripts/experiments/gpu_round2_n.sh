cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -3
run() { # name, env, args
  env $2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/t_c.json')); print('$1 | $3 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], round(d['roofline']['launch_ms'],4), round(d['ms_per_step'],4))
except Exception as e: print('$1 $3 FAILED', e)"
}
for a in "" "--variants 20000" "--variants 5000" "--variants 2000" "--k 21" "--k 15"; do
  run hashbit "X=1" "$a"
done
