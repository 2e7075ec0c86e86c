#!/bin/bash
# T at the other streams-per-CTA variants (1: batch-1 latency, 4: thousands of streams) with alternative ring builds; usage: tools/gpu_alt_t.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-alt_t}
mkdir -p $OUT
for rep in 1 2; do
for lib in default $(ls fastenhancer_b200/_alt/*.so 2>/dev/null); do
  if [ $lib = default ]; then unset FE_LIB; else export FE_LIB=$PWD/$lib; fi
  for a in "16k_t 1 400" "16k_t 4096 50" "16k_t 148 200 1"; do
    echo "$(basename $lib) $(timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep -E 'TIME|rror')"
  done
done
done | tee $OUT/alt_timings.txt
