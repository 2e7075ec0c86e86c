cd /root/repo
export FE_PRECISION=tf32
echo "== HEAD lib"; for i in 1 2 3 4; do FE_LIB=$PWD/fastenhancer_b200/_alt/head.so timeout 120 python tools/gpu_diag.py 48k_l 1 5 12 2>&1 | grep -E "DIAG|rror"; done
echo "== working tree lib"; for i in 1 2 3 4; do timeout 120 python tools/gpu_diag.py 48k_l 1 5 12 2>&1 | grep -E "DIAG|rror"; done
echo "== working tree lib, 16k_l / 48k_m"; for v in "16k_l 1 5 12" "48k_m 1 5 12" "16k_m 1 5 12"; do timeout 120 python tools/gpu_diag.py $v 2>&1 | grep -E "DIAG|rror"; done
echo "== racecheck 48k_l"; timeout 500 compute-sanitizer --tool racecheck --print-limit 30 python tools/gpu_diag.py 48k_l 1 1 2 2>&1 | grep -E "Error|hazard|SUMMARY|and " | cut -c1-200 | head -40
