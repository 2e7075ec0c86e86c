#!/bin/bash
# BASELINE config 4 as stated: FastEnhancer_L 48 kHz, 1024 streams over 4 B200 (torchrun), default fp32 family skipped (slow), f16 family + config 2 on 4 GPUs
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r02x}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --config 4 --precision f16 --steps 3 --warmup 3 --no-extras 2>&1 | tail -1 | tee $OUT/bench_4gpu_config4_f16.json | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 --no-extras 2>&1 | tail -1 | tee $OUT/bench_4gpu.json | cut -c1-260
