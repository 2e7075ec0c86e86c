#!/bin/bash
# first GPU bring-up: parity diagnostics per variant, sanitizer on a tiny case, first timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "from oracle.oracle import build; build()"
for v in "16k_t 1" "16k_t 2" "16k_b 2" "16k_b 1" "16k_t 4" "16k_s 1" "16k_m 1" "16k_l 1" "48k_t 1" "48k_t 2" "48k_b 1" "48k_s 1" "48k_m 1" "48k_l 1"; do
  timeout 120 python tools/gpu_diag.py $v 2>&1 | tail -40
done
echo "=== sanitizer memcheck 16k_t S=2"
timeout 300 compute-sanitizer --tool memcheck python tools/gpu_diag.py 16k_t 2 3 2 2>&1 | tail -15
echo "=== sanitizer racecheck 16k_b S=2"
timeout 400 compute-sanitizer --tool racecheck python tools/gpu_diag.py 16k_b 2 2 2 2>&1 | tail -15
echo "=== timing"
timeout 120 python tools/gpu_diag.py --time 16k_b 256 100
timeout 120 python tools/gpu_diag.py --time 16k_b 256 100 1
timeout 120 python tools/gpu_diag.py --time 16k_t 256 100
timeout 120 python tools/gpu_diag.py --time 16k_b 4096 50
timeout 120 python tools/gpu_diag.py --time 16k_l 148 50
timeout 120 python tools/gpu_diag.py --time 16k_b 1 200
