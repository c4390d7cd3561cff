cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/tuning_sweep.py 31 2500 16 14,4,2,1 14,4,2,2 15,8,2,2 15,16,2,2 15,16,1,2 2>&1 | grep -E "^k=|invalid"
python scripts/tuning_sweep.py 31 25000 16 15,1,2,1 14,4,1,1 14,4,2,2 15,8,2,2 15,16,2,2 15,16,1,2 14,8,1,2 2>&1 | grep -E "^k=|invalid"
