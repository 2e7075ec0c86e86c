#!/bin/bash
# A/B timing of alternative builds on the larger configs; usage: tools/gpu_alt3.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-alt}; mkdir -p $OUT
for rep in 1 2; do
for lib in default $(ls fastenhancer_b200/_alt/*.so 2>/dev/null); do
  if [ $lib = default ]; then unset FE_LIB; else export FE_LIB=$PWD/$lib; fi
  for a in "f16 48k_l 256 64" "f16 16k_l 256 64" "bf16 16k_m 512 64" "f16 16k_b 256 200"; do set -- $a
    echo "$(basename $lib) $(FE_PRECISION=$1 timeout 120 python tools/gpu_diag.py --time $2 $3 $4 2>&1 | grep -E 'TIME|rror')"
  done
done
done | tee $OUT/alt_timings.txt
