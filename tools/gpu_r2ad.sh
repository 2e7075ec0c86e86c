#!/bin/bash
# fp32 FMA family, wide configs: retuned RNNFormer register tiles -- parity + timings (previous build: 48k_l 2521, 16k_l 1200, 16k_m 570 us / hop)
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02ad}; mkdir -p $OUT
echo "=== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(16k_m or 16k_l or 48k_m or 48k_l) and fp32 and not fp32x3 or offline" 2>&1 | tail -3 | tee $OUT/pytest_subset.txt
echo "=== timings"
for a in "fp32 48k_l 256 8" "fp32 16k_l 256 16" "fp32 16k_m 256 32" "fp32 48k_m 256 16"; do
  set -- $a
  FE_PRECISION=$1 timeout 200 python tools/gpu_diag.py --time $2 $3 $4 2>&1 | grep TIME
done | tee $OUT/timings.txt
FE_HOP_SLICING=0 FE_PRECISION=fp32 timeout 300 python tools/gpu_diag.py --prof 48k_l 148 4 2>&1 | head -12
