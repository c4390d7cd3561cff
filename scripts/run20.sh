cd /root/repo
python scripts/tuning_sweep.py 31 2500 16 15,1,2 15,2,2 14,4,2 13,4,2 15,4,2 14,8,2 2>&1 | grep -v Warn
python scripts/tuning_sweep.py 25 2500 16 13,2,2 15,2,2 13,4,2 15,4,2 14,4,2 12,4,2 11,4,2 2>&1 | grep -v Warn
python scripts/tuning_sweep.py 21 2500 16 14,2,2 13,2,2 14,4,2 13,4,2 12,4,2 13,1,2 2>&1 | grep -v Warn
python scripts/tuning_sweep.py 15 2500 16 13,1,2 12,1,2 12,2,2 12,4,2 11,2,2 12,4,1 2>&1 | grep -v Warn
python scripts/tuning_sweep.py 31 250 16 14,4,2 14,4,1 15,8,2 15,8,1 15,16,1 15,16,2 14,16,1 2>&1 | grep -v Warn
