#!/bin/bash
# asynchronous staging of spilled skip tensors / non-resident GRU state: parity + timings.  usage: bash tools/gpu_r2v.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02v}; mkdir -p $OUT
echo "=== parity (M / L / S configs, every family)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "16k_m or 16k_l or 48k_l or 48k_m or 48k_s or 16k_s or hop_sliced" 2>&1 | tail -3 | tee $OUT/pytest_subset.txt
echo "=== timings"
for a in "bf16 16k_m 512 64" "f16 16k_m 256 64" "f16 16k_l 256 64" "f16 48k_l 256 64" "f16 48k_m 256 64" "f16 48k_s 256 64" "tf32 16k_l 256 32" "f16 16k_s 256 100" "fp32x3 16k_b 256 200" "f16 16k_b 256 200"; do
  set -- $a
  FE_PRECISION=$1 timeout 120 python tools/gpu_diag.py --time $2 $3 $4 2>&1 | grep TIME
done | tee $OUT/timings.txt
FE_HOP_SLICING=0 FE_PRECISION=bf16 timeout 120 python tools/gpu_diag.py --prof 16k_m 512 20 2>&1 | tee $OUT/phase_profile_16k_m_bf16.txt | head -16
