# Round-2 GPU call A: topology, smoke, GPU tests, bench both arms, stream-load A/B.
cd /root/repo
mkdir -p gpurun_out
{
  nvidia-smi topo -m; lscpu | head -25; cat /sys/devices/system/node/node*/cpulist; free -g | head -2
  for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done
  ls /usr/lib/x86_64-linux-gnu | grep -i nccl; python -c "import torch,os; print(torch.__file__)"
} > gpurun_out/topo.txt 2>&1
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "ours rc=$?"; cat gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
for v in base ld0; do
  DKB_LIBRARY=ab/libdkb_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "$v rc=$?"
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$v.json')); print('$v', d['value']/1e12, d['roofline']['frac'], d['wgs_shard']['value']/1e12, d['wgs_shard']['roofline']['frac'])"
done
for v in tma2 tma3; do
  for pw in 26624 35584; do
  DKB_PREFILTER_WORDS=$pw DKB_LIBRARY=ab/libdkb_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-wgs > gpurun_out/bench_${v}_$pw.json 2> gpurun_out/bench_${v}_$pw.err; echo "$v $pw rc=$?"
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_${v}_$pw.json')); print('$v $pw', d['value']/1e12, d['roofline']['frac'])"
  done
done
DKB_PREFILTER_WORDS=26624 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-wgs > gpurun_out/bench_ld1_26624.json 2> gpurun_out/bench_ld1_26624.err
python -c "
import json,sys
d=json.load(open('gpurun_out/bench_ld1_26624.json')); print('ld1 26624', d['value']/1e12, d['roofline']['frac'])"
for mb in 16 48; do
  DKB_L2_PERSIST_MB=$mb timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_p$mb.json 2> gpurun_out/bench_p$mb.err; echo "persist $mb rc=$?"
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_p$mb.json')); print('persist$mb', d['value']/1e12, d['roofline']['frac'], d['wgs_shard']['value']/1e12, d['wgs_shard']['roofline']['frac'])"
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
