#!/bin/bash
# bring-up of the tensor-core (TF32) variants: parity per variant, sanitizer, timings, phase profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/tc
export FE_PRECISION=tf32
for v in "16k_b 2" "16k_b 1" "16k_t 1" "16k_t 2" "16k_t 4" "16k_s 1" "16k_m 1" "16k_l 1" "48k_t 1" "48k_t 2" "48k_b 1" "48k_s 1" "48k_m 1" "48k_l 1"; do
  timeout 120 python tools/gpu_diag.py $v 2>&1 | grep -E "DIAG|Error|error|Traceback" | tail -3
done
echo "=== full taps 16k_b S=2"; timeout 120 python tools/gpu_diag.py 16k_b 2 2>&1 | tail -24
echo "=== memcheck tf32 16k_b S=2"
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/tc/memcheck.log python tools/gpu_diag.py 16k_b 2 3 2 > gpurun_out/tc/memcheck_run.log 2>&1; tail -3 gpurun_out/tc/memcheck.log
echo "=== timing"
for a in "16k_b 256 200" "16k_b 256 200 1" "16k_t 256 200" "16k_t 256 200 4" "16k_s 256 100" "16k_m 256 60" "16k_l 256 30" "16k_b 4096 40" "16k_b 1 200" "48k_l 256 20"; do
  timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep -E "TIME|rror"
done
FE_PRECISION=fp32 timeout 120 python tools/gpu_diag.py --time 16k_b 256 200 2>&1 | grep -E "TIME|rror"
echo "=== phase profile"
timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tail -28
timeout 120 python tools/gpu_diag.py --prof 16k_l 148 20 2>&1 | tail -28
