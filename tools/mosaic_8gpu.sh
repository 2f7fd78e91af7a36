#!/bin/bash
# cfg-5 evidence on N GPUs of one box (gpurun --gpus N): (1) 16384^2 in N strips, bit-compared with the whole image
# on one GPU (SURVEY 8e's correctness check), (2) the 32768^2 gigapixel mosaic in N strips.
N=${1:-8}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tools/mosaic_run.py "${@:2}" 2> gpurun_out/mosaic_err_$1.log | grep '^{' ; tail -3 gpurun_out/mosaic_err_$1.log | cut -c1-300; }
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
run 29531 --size 16384 --octaves 4 --verify --reps 3 --pinned | tee gpurun_out/mosaic_16384_${N}gpu.json
run 29532 --size 32768 --octaves 4 --reps 3 --pinned --phases | tee gpurun_out/mosaic_32768_${N}gpu.json
