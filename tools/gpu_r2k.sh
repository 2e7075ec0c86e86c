#!/bin/bash
# two streams per CTA for the larger configs: parity + timings.  usage: bash tools/gpu_r2k.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02k}; mkdir -p $OUT
echo "=== parity (every variant)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "every_variant" 2>&1 | tail -3 | tee $OUT/pytest_variants.txt
echo "=== timings"
for a in "f16 16k_m 512 40" "bf16 16k_m 512 40" "tf32 16k_s 256 100" "bf16 16k_s 256 100" "f16 48k_s 256 60" "fp32x3 48k_b 256 100" "tf32 48k_b 256 100"; do
  set -- $a
  for s in 1 2; do FE_PRECISION=$1 timeout 120 python tools/gpu_diag.py --time $2 $3 $4 $s 2>&1 | grep TIME; done
done | tee $OUT/timings.txt
