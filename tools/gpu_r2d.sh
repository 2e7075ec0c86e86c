#!/bin/bash
# session D (2 GPUs): full GPU test suite incl. NCCL scatter/gather, checkpoint, front door, reference scripts; 2-GPU bench with --scatter
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02d}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.txt
echo "=== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "=== bench --gpus 2 --scatter (config 2)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --scatter --no-extras 2>&1 | tail -1 | tee $OUT/bench_2gpu_scatter.json
echo "=== bench --gpus 2 --config 3 (M, 512 streams/GPU, bf16)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | tee $OUT/bench_2gpu_config3.json
