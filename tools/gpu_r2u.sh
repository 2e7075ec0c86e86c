#!/bin/bash
# 2-GPU session on the final code: multi-GPU tests, config 2 with --scatter, config 3 and config 4 shares, torchrun
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02u}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.txt
echo "=== multi-GPU tests"; timeout 600 python -m pytest tests/test_sharding.py -m gpu -q 2>&1 | tail -3 | tee $OUT/pytest_sharding.txt
echo "=== bench --gpus 2 --scatter (config 2)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --scatter --no-extras 2>&1 | tail -1 | tee $OUT/bench_2gpu_scatter.json | cut -c1-300
echo "=== bench --gpus 2 --config 3 (M, 512 streams/GPU, bf16)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | tee $OUT/bench_2gpu_config3.json | cut -c1-300
echo "=== bench --gpus 2 --config 4 --precision f16 (48 kHz L, 256 streams/GPU)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config 4 --precision f16 --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | tee $OUT/bench_2gpu_config4_f16.json | cut -c1-300
