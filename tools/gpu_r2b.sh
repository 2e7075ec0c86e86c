#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02b}; mkdir -p $OUT
echo "=== MN-major probe"; timeout 60 tools/tc_probe_mn 2>&1 | tee $OUT/tc_probe_mn.txt
echo "=== diag fp32x3 (accurate activations)"
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py 16k_b 2 2>&1 | tail -2 | tee $OUT/diag_fp32x3_b.txt
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py --time 16k_b 256 200 2>&1 | grep TIME | tee $OUT/timings.txt
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py --time 16k_b 4096 40 2>&1 | grep TIME | tee -a $OUT/timings.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --time 16k_b 4096 40 2>&1 | grep TIME | tee -a $OUT/timings.txt
echo "=== pytest subset"; timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "network or long or smoke or scripts or concurrently" -s 2>&1 | grep -E "network-dominated|passed|failed|Error" | tee $OUT/pytest_subset.txt
echo "=== phase profile fp32x3"
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | head -12 | tee $OUT/phase_profile_16k_b_fp32x3.txt
