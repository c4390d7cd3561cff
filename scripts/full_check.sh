# Full round check on a GPU box: smoke, GPU tests, both bench arms.  Usage: bash scripts/full_check.sh
cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; cat gpurun_out/bench_ours.json
