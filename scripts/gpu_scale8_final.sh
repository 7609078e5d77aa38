#!/bin/bash
# usage (under gpurun --gpus 8): bash scripts/gpu_scale8_final.sh <tag>  -- N=8 with the item axis over all ranks (headline layout) and the grid form
TAG=${1:-rX}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --cpu-budget 3 --parity-users 1024 > gpurun_out/scale_${TAG}_n8.json 2> gpurun_out/scale_${TAG}_n8.err
tail -2 gpurun_out/scale_${TAG}_n8.err | cut -c1-300; python scripts/show_bench.py gpurun_out/scale_${TAG}_n8.json
timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --item-shards 2 --steps 10 --warmup 3 --cpu-budget 2 --parity-users 512 > gpurun_out/scale_${TAG}_n8_grid4x2.json 2> gpurun_out/scale_${TAG}_n8_grid4x2.err
tail -2 gpurun_out/scale_${TAG}_n8_grid4x2.err | cut -c1-300; python scripts/show_bench.py gpurun_out/scale_${TAG}_n8_grid4x2.json
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --cpu-budget 2 --no-extra --parity-users 512 > gpurun_out/scale_${TAG}_n1.json 2> gpurun_out/scale_${TAG}_n1.err
python scripts/show_bench.py gpurun_out/scale_${TAG}_n1.json
