# Round-2 GPU call C: probe array A/B in the shared-memory filter mode, small-launch shapes.
cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
run() { # name, env, args
  env $2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/t_c.json')); print('$1 | $3 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['seeds'], d['config']['prefilter_words'], round(d['roofline']['launch_ms'],4), round(d['ms_per_step'],4))
except Exception as e: print('$1 $3 FAILED', e)"
}
for v in 250 1000 2000; do
  for t in 15,16,2,1 14,16,2,1; do
    run probe "DKB_TUNING=$t" "--variants $v"
    run noprobe "DKB_TUNING=$t DKB_NO_PROBE_ARRAY=1" "--variants $v"
    run noprobe8 "DKB_TUNING=$t DKB_NO_PROBE_ARRAY=1 DKB_SEED_SLOTS_PER_SEED=8" "--variants $v"
  done
  run auto "X=1" "--variants $v"
done
run ld0 "DKB_TUNING=15,16,2,1 DKB_LIBRARY=ab/libdkb_ld0.so" "--variants 250"
run sep "DKB_TUNING=15,16,2,1" "--variants 250 --separate-launches"
run cfg0 "X=1" "--genome-mb 1 --variants 100"
run cfg0sep "X=1" "--genome-mb 1 --variants 100 --separate-launches"
for t in 15,1,2,1 14,2,2,2 13,2,2,2 12,4,2,2 12,4,1,2 11,4,2,2; do run k15 "DKB_TUNING=$t" "--k 15"; done
for t in 15,4,2,2 14,8,2,2 13,8,2,2 14,8,1,2; do run k21 "DKB_TUNING=$t" "--k 21"; done
for v in 3000 4000 7000; do run auto "X=1" "--variants $v"; run l2pre "DKB_TUNING=15,16,2,2" "--variants $v"; done
run wgs1 "DKB_TUNING=15,16,1,2" "--genome-mb 128 --variants 4000 --table-variants 100000"
run wgs2 "DKB_TUNING=15,16,2,2" "--genome-mb 128 --variants 4000 --table-variants 100000"
run wgs2pre "DKB_TUNING=15,16,2,2 DKB_PREFILTER_WORDS=35584" "--genome-mb 128 --variants 4000 --table-variants 100000"
