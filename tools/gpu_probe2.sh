#!/bin/bash
# tcgen05 shape / operand-source micro-benchmarks + parity + per-phase sub-timer profile
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-p2}
mkdir -p $OUT
for g in ${GROUPS_TC:-2}; do timeout 60 ./tools/tc_bench2 $g 2>&1 | tee -a $OUT/tc_bench2.txt; done
for v in "16k_b 2" "16k_b 1" "16k_t 4" "16k_s 1" "16k_l 1" "48k_m 1"; do
  timeout 120 python tools/gpu_diag.py $v 2>&1 | grep -E "DIAG|Error|error|Traceback" | tail -3
done | tee $OUT/diag.txt
timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b.txt
for a in "16k_t 256 200" "16k_b 256 200" "16k_s 256 100" "16k_m 256 60" "16k_l 256 30" "16k_b 1 200" "16k_b 4096 40"; do
  timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep TIME
done | tee $OUT/timings.txt
