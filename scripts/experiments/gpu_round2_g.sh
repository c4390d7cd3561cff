cd /root/repo
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_configs.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -5
run() { # name, env, args
  env $2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/t_c.json')); print('$1 | $3 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['seeds'], d['config']['prefilter_words'], round(d['roofline']['launch_ms'],4), round(d['ms_per_step'],4))
except Exception as e: print('$1 $3 FAILED', e)"
}
for a in "" "--variants 1000" "--variants 250" "--variants 20000" "--genome-mb 128 --variants 4000 --table-variants 100000" "--depth 100 --variants 50000 --indel-frac 0.5" "--k 21" "--k 15" "--genome-mb 1 --variants 100"; do
  run sync "DKB_LIBRARY=ab/libdkb_sync.so" "$a"
  run h2 "X=1" "$a"
  run h1 "DKB_LIBRARY=ab/libdkb_h1.so DKB_TUNING=0,0,2,0" "$a"
  run h3 "DKB_LIBRARY=ab/libdkb_h3.so DKB_TUNING=0,0,2,0" "$a"
  run h4 "DKB_LIBRARY=ab/libdkb_h4.so DKB_TUNING=0,0,2,0" "$a"
done
