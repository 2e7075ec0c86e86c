#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02p}; mkdir -p $OUT
echo "=== offline tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "offline" 2>&1 | tail -3
echo "=== bench config 1"; timeout 600 python bench.py --config 1 2>&1 | tail -1 | tee $OUT/bench_config1.json | cut -c1-1500
echo "=== bench config 1 reference"; timeout 600 python bench.py --config 1 --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_config1_reference.json | cut -c1-1200
