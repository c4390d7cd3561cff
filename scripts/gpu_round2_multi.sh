# Round-2 multi-GPU call: NCCL test (2 ranks), bench at N GPUs (weak default + strong).  Usage: bash scripts/gpu_round2_multi.sh N
N=${1:-2}
cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Socket|Model name" >> gpurun_out/topo_$N.txt
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo "$d numa $(cat $d/numa_node) class $(cat $d/class)"; fi; done >> gpurun_out/topo_$N.txt
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q 2>&1 | tail -5
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
NCCL_DEBUG=INFO timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "weak rc=$?"
cat gpurun_out/bench_${N}gpu.json | cut -c1-6000
grep -E "nranks|NVLS|Init COMPLETE|comm 0x" gpurun_out/bench_${N}gpu.err | head -8
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-wgs --no-e2e > gpurun_out/bench_${N}gpu_strong.json 2> gpurun_out/bench_${N}gpu_strong.err; echo "strong rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_${N}gpu_strong.json')); print('strong', $N, round(d['value']/1e12,3), d['ms_per_step'], d['checks'])"
