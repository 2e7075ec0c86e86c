#!/bin/bash
# tensor-core STFT operator: parity tests + timings.  usage: bash tools/gpu_r2j.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02j}; mkdir -p $OUT
echo "=== stft gemm tests"; timeout 300 python -m pytest tests/test_stft_gemm.py -m gpu -q 2>&1 | tail -25 | tee $OUT/pytest_stft_gemm.txt
echo "=== timings"; timeout 300 python tools/stft_timing.py 2>&1 | grep -E "STFT|rror" | tee $OUT/stft_timings.txt
