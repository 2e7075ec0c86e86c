#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02c}; mkdir -p $OUT
echo "=== parity: every variant / taps / golden, f16 + fp32x3 + bf16 (tensor-core frequency-axis linears, S=4 B)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=10 -k "every_variant or stage_taps or streaming_matches or network or offline or spec2spec" 2>&1 | tail -6 | tee $OUT/pytest_subset.txt
echo "=== timings"
for lib in default nolintc h2silu; do
  if [ $lib = default ]; then unset FE_LIB; else export FE_LIB=$PWD/fastenhancer_b200/_alt/$lib.so; fi
  for p in fp32x3 f16; do
    for a in "16k_b 256 200" "16k_b 4096 40" "16k_t 256 200"; do
      echo "$lib $(FE_PRECISION=$p timeout 120 python tools/gpu_diag.py --time $a 2>&1 | grep -E 'TIME|rror')"
    done
  done
done | tee $OUT/timings.txt
unset FE_LIB
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --time 16k_b 4096 40 2 2>&1 | grep TIME | tee -a $OUT/timings.txt
FE_PRECISION=bf16 timeout 120 python tools/gpu_diag.py --time 16k_m 512 40 2>&1 | grep TIME | tee -a $OUT/timings.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --time 16k_l 256 30 2>&1 | grep TIME | tee -a $OUT/timings.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --time 48k_l 256 20 2>&1 | grep TIME | tee -a $OUT/timings.txt
echo "=== h2silu parity"
FE_LIB=$PWD/fastenhancer_b200/_alt/h2silu.so FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py 16k_b 2 2>&1 | tail -2 | tee $OUT/diag_h2silu.txt
echo "=== phase profiles"
FE_PRECISION=fp32x3 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_fp32x3.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --prof 16k_b 256 50 2>&1 | tee $OUT/phase_profile_16k_b_f16.txt
FE_PRECISION=f16 timeout 120 python tools/gpu_diag.py --prof 16k_b 4096 20 4 2>&1 | tee $OUT/phase_profile_16k_b_f16_s4.txt
