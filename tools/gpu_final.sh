#!/bin/bash
# last check of HEAD: full GPU suite, smoke, default bench, config 3
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $OUT/smoke.txt
echo "=== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-200
echo "=== bench config 3"; timeout 900 python bench.py --config 3 --steps 5 2>&1 | tail -1 | tee $OUT/bench_config3.json | cut -c1-200
echo "=== bench config 4 (default fp32 family)"; timeout 900 python bench.py --config 4 --steps 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee $OUT/bench_config4.json | cut -c1-200
