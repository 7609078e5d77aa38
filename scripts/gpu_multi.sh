#!/bin/bash
# usage (under gpurun --gpus N): bash scripts/gpu_multi.sh <tag> <N> [counts]   (counts default "1 N", e.g. "8" or "2 4 8")
TAG=${1:-rX}
N=${2:-2}
COUNTS=${3:-"1 $N"}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x 2>&1 | tail -5
for n in $COUNTS; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --cpu-budget 3 --no-extra --parity-users 1024 > gpurun_out/scale_${TAG}_n1.json 2> gpurun_out/scale_${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --cpu-budget 3 --parity-users 1024 > gpurun_out/scale_${TAG}_n${n}.json 2> gpurun_out/scale_${TAG}_n${n}.err
  fi
  tail -2 gpurun_out/scale_${TAG}_n${n}.err | cut -c1-300
  python scripts/show_bench.py gpurun_out/scale_${TAG}_n${n}.json
done
