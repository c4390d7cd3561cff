# Full round check on a GPU box: smoke, GPU tests, both bench arms.  Usage: bash scripts/full_check.sh
cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; cat gpurun_out/bench_ours.json
