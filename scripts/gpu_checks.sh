#!/bin/bash
# usage: bash scripts/gpu_checks.sh <tag> [quick]   -- run on the GPU box through gpurun; outputs under gpurun_out/
TAG=${1:-rX}
MODE=${2:-full}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
if [ "$MODE" = "quick" ]; then
  timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_golden_fixtures.py -m gpu -q -x 2>&1 | tail -8
else
  timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; tail -2 gpurun_out/bench_${TAG}_n1.err; cat gpurun_out/bench_${TAG}_n1.json
timeout 300 python bench.py --workload dense --users 65536 --items 100000 --d 64 --steps 3 --warmup 2 > gpurun_out/bench_${TAG}_dense.json 2>/dev/null; cat gpurun_out/bench_${TAG}_dense.json
timeout 300 python bench.py --workload ranks --users 8192 --items 131072 --steps 3 --warmup 2 > gpurun_out/bench_${TAG}_ranks.json 2>gpurun_out/bench_${TAG}_ranks.err; tail -2 gpurun_out/bench_${TAG}_ranks.err; cat gpurun_out/bench_${TAG}_ranks.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --users 262144 --items 262144 --cpu-budget 1 > /dev/null 2> gpurun_out/ncu_launches.err; tail -1 gpurun_out/ncu_launches.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_kernel -s 1 -c 1 -o gpurun_out/prof_fused_${TAG} -f python bench.py --steps 1 --warmup 1 --cpu-budget 1 > /dev/null 2> gpurun_out/ncu_fused.err; tail -1 gpurun_out/ncu_fused.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_gather_reduce -s 2 -c 1 -o gpurun_out/prof_k1_${TAG} -f python bench.py --steps 1 --warmup 1 --users 1000000 --items 262144 --cpu-budget 1 > /dev/null 2> gpurun_out/ncu_k1.err; tail -1 gpurun_out/ncu_k1.err
ls -la gpurun_out | tail -12
