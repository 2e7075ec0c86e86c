#!/bin/bash
# A/B timing of alternative builds for both headline families; usage: tools/gpu_alt2.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-alt}
mkdir -p $OUT
for rep in 1 2 3; do
for lib in default $(ls fastenhancer_b200/_alt/*.so 2>/dev/null); do
  if [ $lib = default ]; then unset FE_LIB; else export FE_LIB=$PWD/$lib; fi
  for p in fp32x3 f16; do for a in "16k_b 256 200" "16k_t 256 200" "16k_b 1 200"; do
    echo "$(basename $lib) $(FE_PRECISION=$p timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep -E 'TIME|rror')"
  done; done
done
done | tee $OUT/alt_timings.txt
