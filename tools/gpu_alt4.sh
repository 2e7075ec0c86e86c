#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2; do
for lib in default $(ls fastenhancer_b200/_alt/*.so 2>/dev/null); do
  if [ $lib = default ]; then unset FE_LIB; else export FE_LIB=$PWD/$lib; fi
  for a in "fp32 48k_l 256 8" "fp32 16k_l 256 16"; do set -- $a
    echo "$(basename $lib) $(FE_PRECISION=$1 timeout 120 python tools/gpu_diag.py --time $2 $3 $4 2>&1 | grep -E 'TIME|rror')"
  done
done
done
