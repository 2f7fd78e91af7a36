#!/bin/bash
# 2-GPU sanity: image-sharded bench line and the mosaic path with phase timing (gpurun --gpus 2)
if [ "${1:-all}" = "all" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2))"
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/mosaic_run.py --size 8192 --octaves 4 --reps 3 --phases --pinned 2>/dev/null | grep '^{'
