cd /root/repo
mkdir -p gpurun_out
run() { # name, env, args
  env $2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/t_c.json')); print('$1 | $3 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['seeds'], d['config']['prefilter_words'], round(d['roofline']['launch_ms'],4), round(d['ms_per_step'],4))
except Exception as e: print('$1 $3 FAILED', e)"
}
run base "X=1" ""
run nostagec "DKB_LIBRARY=ab/libdkb_x6.so" ""
run perseed4 "DKB_SEED_SLOTS_PER_SEED=4" ""
run perseed8 "DKB_SEED_SLOTS_PER_SEED=8" ""
run key4 "DKB_KEY_SLOTS_PER_ENTRY=4" ""
run key16 "DKB_KEY_SLOTS_PER_ENTRY=16" ""
run l2f32 "DKB_L2_FILTER_MAX_WORDS=600000" ""
run base1000 "DKB_TUNING=15,16,2,1" "--variants 1000"
run base250 "DKB_TUNING=15,16,2,1" "--variants 250"
run wgs "X=1" "--genome-mb 128 --variants 4000 --table-variants 100000"
run wgs_nostagec "DKB_LIBRARY=ab/libdkb_x6.so" "--genome-mb 128 --variants 4000 --table-variants 100000"
run wgs_perseed4 "DKB_SEED_SLOTS_PER_SEED=4" "--genome-mb 128 --variants 4000 --table-variants 100000"
