#!/bin/bash
# fp16 conv-section variants: parity, timing against tf32, phase profile, memcheck
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-f16}
mkdir -p $OUT
export FE_PRECISION=f16
for v in "16k_t 4" "16k_s 2" "16k_m 1" "16k_l 1" "48k_b 2" "48k_l 1"; do timeout 120 python tools/gpu_diag.py $v 2>&1 | grep -E "DIAG|rror"; done | tee $OUT/diag.txt
timeout 120 python tools/gpu_diag.py 16k_b 2 2>&1 | tail -24 | tee $OUT/taps_b.txt
for p in f16 tf32; do
  export FE_PRECISION=$p
  for a in "16k_b 256 200" "16k_t 256 200" "16k_s 256 100" "16k_m 256 60" "16k_l 256 30" "48k_l 256 20" "16k_b 4096 40" "16k_t 4096 50" "16k_b 1 200"; do timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep -E "TIME|rror"; done
done | tee $OUT/timings.txt
export FE_PRECISION=f16
timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_f16.txt
timeout 300 compute-sanitizer --tool memcheck python tools/gpu_diag.py 16k_b 2 3 2 2>&1 | tail -3 | tee $OUT/memcheck.txt
