# the workload table of profiles/README.md on the final build
cd /root/repo
mkdir -p gpurun_out
run() { # args
  timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-wgs $1 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/t_c.json')); print('$1 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['seeds'], round(d['roofline']['launch_ms'],4), round(d['ms_per_step'],4))
except Exception as e: print('$1 FAILED', e)"
}
for a in "" "--variants 20000" "--variants 7000" "--variants 5000" "--variants 3000" "--variants 2000" "--variants 1000" "--variants 250" "--k 25" "--k 21" "--k 15" "--genome-mb 128 --variants 4000 --table-variants 100000" "--depth 100 --variants 50000 --indel-frac 0.5" "--genome-mb 1 --variants 100"; do run "$a"; done
