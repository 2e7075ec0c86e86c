#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02e}; mkdir -p $OUT
echo "=== parity with TMA hop tiles"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=10 -k "every_variant or streaming_matches or hop_by_hop or hop_tiles or host_buffer or full_size or delay or model_classes" 2>&1 | tail -6 | tee $OUT/pytest_subset.txt
echo "=== timings TMA vs plain"
for t in 1 0; do for p in fp32x3 f16; do
  echo "FE_HOP_TMA=$t $(FE_HOP_TMA=$t FE_PRECISION=$p timeout 120 python tools/gpu_diag.py --time 16k_b 256 200 2>&1 | grep -E 'TIME|rror')"
done; done | tee $OUT/timings.txt
echo "=== memcheck (TMA path)"
FE_PRECISION=fp32x3 timeout 300 compute-sanitizer --tool memcheck --log-file $OUT/memcheck.log python tools/gpu_diag.py 16k_b 2 3 3 > $OUT/memcheck_run.log 2>&1; tail -3 $OUT/memcheck.log
