#!/bin/bash
# usage (under gpurun --gpus 8): bash scripts/gpu_scale8.sh <tag>   -- the driver's scaling run at N=8 (+ N=1 on the same box) and C5 at its real size
TAG=${1:-rX}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --cpu-budget 3 --parity-users 1024 > gpurun_out/scale_${TAG}_n8.json 2> gpurun_out/scale_${TAG}_n8.err
tail -2 gpurun_out/scale_${TAG}_n8.err | cut -c1-300; python scripts/show_bench.py gpurun_out/scale_${TAG}_n8.json
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --cpu-budget 3 --no-extra --parity-users 1024 > gpurun_out/scale_${TAG}_n1.json 2> gpurun_out/scale_${TAG}_n1.err
python scripts/show_bench.py gpurun_out/scale_${TAG}_n1.json
timeout 600 $TR --nproc-per-node 4 bench.py --gpus 4 --steps 20 --warmup 5 --cpu-budget 3 --parity-users 1024 > gpurun_out/scale_${TAG}_n4.json 2> gpurun_out/scale_${TAG}_n4.err
python scripts/show_bench.py gpurun_out/scale_${TAG}_n4.json
# BASELINE configs[4] at its real size: 10M users x 1M items, item-sharded over 8 GPUs (the API arm batches the users)
timeout 1200 $TR --nproc-per-node 8 bench.py --gpus 8 --users 10000000 --steps 2 --warmup 1 --cpu-budget 2 --parity-users 512 --user-batch 2500000 > gpurun_out/c5_${TAG}_n8_10m.json 2> gpurun_out/c5_${TAG}_n8_10m.err
tail -2 gpurun_out/c5_${TAG}_n8_10m.err | cut -c1-300; python scripts/show_bench.py gpurun_out/c5_${TAG}_n8_10m.json
