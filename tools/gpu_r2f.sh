#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02f}; mkdir -p $OUT
for rep in 1 2; do
for lib in default pretma; do
  if [ $lib = default ]; then unset FE_LIB; else export FE_LIB=$PWD/fastenhancer_b200/_alt/$lib.so; fi
  for p in f16 fp32x3; do
  echo "$lib $(FE_PRECISION=$p timeout 120 python tools/gpu_diag.py --time 16k_b 256 200 2>&1 | grep -E 'TIME|rror')"
  done
done
done | tee $OUT/timings.txt
export FE_LIB=$PWD/fastenhancer_b200/_alt/pretma.so
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | head -8 | tee $OUT/phase_profile_pretma.txt
