cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
run() { env $2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
d=json.load(open('gpurun_out/t_c.json')); print('$1 | $3 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['tuning_seedlen_stride_hashes_filtermode'], d['config']['prefilter_words'], round(d['roofline']['launch_ms'],4))"; }
for a in "" "--variants 20000" "--variants 5000" "--variants 1000" "--variants 250" "--k 21" "--genome-mb 128 --variants 4000 --table-variants 100000" "--depth 100 --variants 50000 --indel-frac 0.5"; do run gated "X=1" "$a"; done
