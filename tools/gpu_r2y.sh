#!/bin/bash
# BASELINE config 3 as stated: FastEnhancer_M, 4096 streams over 8 B200 (torchrun), bf16 conv / fp32 GRU; and config 2 on 8 GPUs
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02y}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --config 3 --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | tee $OUT/bench_8gpu_config3.json | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras 2>&1 | tail -1 | tee $OUT/bench_8gpu.json | cut -c1-260
