cd /root/repo
for L in x6; do
echo "== $L"
DKB_LIBRARY=/root/repo/ab/libdkb_$L.so python scripts/quick_scan_bench.py 10000 1.9e9 31 14,4,2 15,4,2 2>&1 | grep -A1 '"hints": true'
done
DKB_LIBRARY=/root/repo/ab/libdkb_x0.so python scripts/quick_scan_bench.py 10000 1.9e9 31 15,4,2 2>&1 | grep -A1 '"hints": true'
