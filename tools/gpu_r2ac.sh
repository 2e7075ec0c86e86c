#!/bin/bash
# fp32 FMA family: chunked float4 attention -- parity + timings
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02ac}; mkdir -p $OUT
echo "=== parity (fp32 family + offline)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fp32 and not fp32x3 or offline" 2>&1 | tail -3 | tee $OUT/pytest_subset.txt
echo "=== timings"
for a in "fp32 48k_l 256 8" "fp32 16k_l 256 16" "fp32 16k_m 256 32" "fp32 16k_b 256 100" "fp32 16k_t 256 100"; do
  set -- $a
  FE_PRECISION=$1 timeout 200 python tools/gpu_diag.py --time $2 $3 $4 2>&1 | grep TIME
done | tee $OUT/timings.txt
timeout 300 python tools/offline_timing.py 16k_t:1:10 16k_b:1:10 16k_m:1:10 48k_l:1:10 2>&1 | grep OFFLINE | cut -c1-150 | tee $OUT/offline_timings.txt
