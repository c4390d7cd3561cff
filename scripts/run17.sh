cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "rc=$?"; tail -2 gpurun_out/bench_8gpu.err; cat gpurun_out/bench_8gpu.json | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('8 GPU value %.3f T/s' % (d['value']/1e12), 'ms/step %.3f' % d['ms_per_step'], 'e2e %.1f G/s' % (d['e2e']['value']/1e9), d['config']['collective'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('4 GPU value %.3f T/s' % (d['value']/1e12), 'ms/step %.3f' % d['ms_per_step'])"
