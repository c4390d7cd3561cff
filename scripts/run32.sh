cd /root/repo
for nv in 1000 2500 5000 10000; do
python scripts/tuning_sweep.py 31 $nv 32 14,4,2 14,8,2 15,8,2 15,16,2 2>&1 | grep -E "^k=|invalid"
done
