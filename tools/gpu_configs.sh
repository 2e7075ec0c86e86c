#!/bin/bash
# bench.py lines for the other BASELINE shapes on one GPU (their per-GPU shards), then the alternative-build timings.
# usage: bash tools/gpu_configs.sh TAG        (outputs under gpurun_out/TAG/)
cd "$(dirname "$0")/.."
TAG=${1:-cfg}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift; echo "=== $name"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | tee $OUT/bench_$name.json | cut -c1-260; }
run 16k_t_256 --preset 16k_t --streams 256
run 16k_s_256 --preset 16k_s --streams 256
run 16k_m_512 --preset 16k_m --streams 512
run 16k_l_256 --preset 16k_l --streams 256
run 48k_l_256 --preset 48k_l --streams 256
run 16k_b_4096 --preset 16k_b --streams 4096 --seconds 2
run 16k_b_1 --preset 16k_b --streams 1
if ls fastenhancer_b200/_alt/*.so > /dev/null 2>&1; then bash tools/gpu_alt.sh $TAG; fi
