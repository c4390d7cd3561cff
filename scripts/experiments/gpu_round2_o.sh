cd /root/repo
run() { env $2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-wgs $3 > gpurun_out/t_c.json 2> gpurun_out/t_c.err
  python -c "
import json
d=json.load(open('gpurun_out/t_c.json')); print('$1 | $3 |', round(d['value']/1e12,3), round(d['roofline']['frac'],4), d['config']['prefilter_words'], round(d['roofline']['launch_ms'],4))"; }
run pw35584 "X=1" ""
run pw43776 "DKB_PREFILTER_WORDS=43776" ""
run pw27392 "DKB_PREFILTER_WORDS=27392" ""
run pw43776 "DKB_PREFILTER_WORDS=43776" "--variants 20000"
run pw35584 "X=1" "--variants 20000"
