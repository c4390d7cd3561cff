cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge.py -m gpu -q -x -k "zero_list or one_launch or padding" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-wgs > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "rc=$?"; tail -3 gpurun_out/bench_k.err
python -c "
import json
d=json.load(open('gpurun_out/bench_k.json')); print(d['value']/1e12, d['roofline']['frac']); print(json.dumps(d['e2e'],indent=1)); print(d['checks'], d['clocks'])"
