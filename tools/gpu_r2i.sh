#!/bin/bash
# frame-parallel offline schedule: parity tests + timings.  usage: bash tools/gpu_r2i.sh TAG
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02i}; mkdir -p $OUT
echo "=== offline tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "offline" 2>&1 | tail -5 | tee $OUT/pytest_offline.txt
echo "=== timings"; timeout 600 python tools/offline_timing.py 2>&1 | grep -E "OFFLINE|Error|error" | tee $OUT/offline_timings.txt
echo "=== per-launch"; for c in 16k_t:1:10 16k_b:1:10 16k_m:1:10 48k_l:1:10; do FE_TP_TIMING=1 timeout 100 python tools/offline_timing.py $c 2>&1 | grep "per launch" | tail -1 | sed "s/^/$c /" | tee -a $OUT/offline_timings.txt; done
