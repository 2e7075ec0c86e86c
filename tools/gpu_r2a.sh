#!/bin/bash
# Round-2 GPU session A: parity of the new kernel families + first numbers.   usage: bash tools/gpu_r2a.sh TAG
cd "$(dirname "$0")/.."
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/gpu.txt
echo "=== diag fp32x3 / bf16"
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py 16k_b 2 2>&1 | tail -4 | tee $OUT/diag_fp32x3_b.txt
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py 16k_t 2 2>&1 | tail -2 | tee $OUT/diag_fp32x3_t.txt
FE_PRECISION=bf16 timeout 120 python tools/gpu_diag.py 16k_m 1 2>&1 | tail -2 | tee $OUT/diag_bf16_m.txt
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "=== timings"
for p in fp32x3 f16 tf32 fp32; do
  FE_PRECISION=$p timeout 120 python tools/gpu_diag.py --time 16k_b 256 200 2>&1 | grep TIME
done | tee $OUT/timings.txt
FE_PRECISION=bf16 timeout 120 python tools/gpu_diag.py --time 16k_m 512 40 2>&1 | grep TIME | tee -a $OUT/timings.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --time 16k_m 512 40 2>&1 | grep TIME | tee -a $OUT/timings.txt
echo "=== phase profile fp32x3"
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_fp32x3.txt
echo "=== bench (default: config 2)"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
ls -la $OUT
