#!/bin/bash
# idle-warp skip in the RNNFormer epilogues: parity subset + timings.  usage: bash tools/gpu_r2m.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02m}; mkdir -p $OUT
echo "=== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "streaming_matches or every_variant or taps" 2>&1 | tail -3 | tee $OUT/pytest_subset.txt
echo "=== timings"
for a in "fp32x3 16k_b 256 200" "fp32x3 16k_t 256 200" "f16 16k_b 256 200" "f16 16k_t 256 200" "f16 16k_s 256 100" "bf16 16k_m 512 64" "f16 48k_l 256 40" "fp32x3 16k_b 1 200" "f16 16k_b 4096 40"; do
  set -- $a
  FE_PRECISION=$1 timeout 120 python tools/gpu_diag.py --time $2 $3 $4 2>&1 | grep TIME
done | tee $OUT/timings.txt
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_fp32x3.txt | head -12
