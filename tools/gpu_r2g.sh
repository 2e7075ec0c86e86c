#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02g}; mkdir -p $OUT
echo "=== diag (mma.sync attention)"
for p in fp32x3 f16 tf32; do FE_PRECISION=$p timeout 120 python tools/gpu_diag.py 16k_b 2 2>&1 | grep -E "blk|DIAG"; done | tee $OUT/diag.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py 16k_s 1 2>&1 | grep -E "DIAG" | tee -a $OUT/diag.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py 16k_l 1 2>&1 | grep -E "DIAG" | tee -a $OUT/diag.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py 48k_l 1 2>&1 | grep -E "DIAG" | tee -a $OUT/diag.txt
echo "=== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=10 -k "every_variant or stage_taps or streaming_matches or network or offline or spec2spec" 2>&1 | tail -4 | tee $OUT/pytest_subset.txt
echo "=== timings"
for lib in default noring; do
  if [ $lib = default ]; then unset FE_LIB; else export FE_LIB=$PWD/fastenhancer_b200/_alt/$lib.so; fi
  for p in fp32x3 f16; do for a in "16k_b 256 200" "16k_t 256 200" "16k_b 4096 40"; do
    echo "$lib $(FE_PRECISION=$p timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep -E 'TIME|rror')"
  done; done
done | tee $OUT/timings.txt
unset FE_LIB
for a in "16k_s 256 100" "16k_m 256 60" "16k_l 256 30" "48k_l 256 20"; do FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep TIME; done | tee -a $OUT/timings.txt
FE_PRECISION=bf16 timeout 120 python tools/gpu_diag.py --time 16k_m 512 40 2>&1 | grep TIME | tee -a $OUT/timings.txt
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | head -14 | tee $OUT/phase_profile_16k_b_fp32x3.txt
