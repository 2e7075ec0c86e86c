#!/bin/bash
# hop-sliced streaming launches: bit-identity tests + timings.  usage: bash tools/gpu_r2l.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02l}; mkdir -p $OUT
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "hop_sliced or hop_by_hop" 2>&1 | tail -4 | tee $OUT/pytest_sliced.txt
echo "=== timings (FE_HOP_SLICING = 0 / 1)"
for a in "bf16 16k_m 512 64" "f16 16k_m 512 200" "f16 16k_l 256 64" "f16 16k_l 256 200" "f16 48k_l 256 64" "f16 48k_l 256 200" "fp32 48k_l 256 16" "f16 48k_m 256 64" "f16 16k_m 4096 32"; do
  set -- $a
  for on in 0 1; do FE_HOP_SLICING=$on FE_PRECISION=$1 timeout 120 python tools/gpu_diag.py --time $2 $3 $4 2>&1 | grep TIME | sed "s/^/slicing=$on /"; done
done | tee $OUT/timings.txt
