#!/bin/bash
# time alternative builds (fastenhancer_b200/_alt/*.so) against the default one; usage: tools/gpu_alt.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-alt}
mkdir -p $OUT
for rep in $(seq 1 ${REPS:-2}); do
for lib in default $(ls fastenhancer_b200/_alt/*.so 2>/dev/null); do
  if [ $lib = default ]; then unset FE_LIB; else export FE_LIB=$PWD/$lib; fi
  for a in "16k_b 256 200" "16k_t 256 200" "16k_b 4096 40"; do
    echo "$(basename $lib) $(timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep -E 'TIME|rror')"
  done
done
done | tee $OUT/alt_timings.txt
