#!/bin/bash
# usage: bash scripts/gpu_checks.sh <tag> [tests|bench|ab|ncu ...]   -- run on the GPU box through gpurun; outputs under gpurun_out/
TAG=${1:-rX}
shift
WHAT=${@:-tests bench ab ncu}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
B="python bench.py --no-extra --cpu-budget 2 --parity-users 256 --steps 5 --warmup 3"
for w in $WHAT; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
  ;;
bench)
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; tail -3 gpurun_out/bench_${TAG}_n1.err; cat gpurun_out/bench_${TAG}_n1.json
  ;;
ab)
  # one shard of 8 on one GPU (what each rank of an 8-GPU run computes, minus the exchange), with / without the first-tile threshold
  timeout 600 $B --emulate-shards 8 > gpurun_out/bench_${TAG}_shard8.json 2> gpurun_out/bench_${TAG}_shard8.err; tail -2 gpurun_out/bench_${TAG}_shard8.err; cat gpurun_out/bench_${TAG}_shard8.json
  TRK_FILTER_NO_WARMSTART=1 timeout 600 $B --emulate-shards 8 > gpurun_out/bench_${TAG}_shard8_nowarm.json 2>/dev/null; cat gpurun_out/bench_${TAG}_shard8_nowarm.json
  TRK_FILTER_NO_WARMSTART=1 timeout 600 $B > gpurun_out/bench_${TAG}_n1_nowarm.json 2>/dev/null; cat gpurun_out/bench_${TAG}_n1_nowarm.json
  timeout 600 $B --no-clocks > gpurun_out/bench_${TAG}_n1_noclocks.json 2>/dev/null; cat gpurun_out/bench_${TAG}_n1_noclocks.json
  ;;
probe)
  timeout 300 python scripts/filter_probe.py 1000000 125000 2>&1 | grep -v "^\[bench\]" | tee gpurun_out/probe_${TAG}_filter_shard8.txt
  timeout 600 python scripts/filter_probe.py 1000000 1000000 2>&1 | grep -v "^\[bench\]" | tee gpurun_out/probe_${TAG}_filter_full.txt
  timeout 300 python scripts/k1_probe.py 1000000 2>&1 | grep -v "^\[bench\]" | tee gpurun_out/probe_${TAG}_k1.txt
  ;;
scores)
  for sc in ties c3; do
    timeout 900 python bench.py --no-extra --scores $sc --steps 3 --warmup 2 --cpu-budget 3 --parity-users 512 > gpurun_out/bench_${TAG}_scores_${sc}.json 2> gpurun_out/bench_${TAG}_scores_${sc}.err; tail -1 gpurun_out/bench_${TAG}_scores_${sc}.err | cut -c1-200; python scripts/show_bench.py gpurun_out/bench_${TAG}_scores_${sc}.json
  done
  ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --no-extra --steps 2 --warmup 1 --users 262144 --items 262144 --cpu-budget 1 --parity-users 64 > /dev/null 2> gpurun_out/ncu_launches.err; tail -1 gpurun_out/ncu_launches.err
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_kernel -s 1 -c 1 -o gpurun_out/prof_fused_${TAG} -f python bench.py --no-extra --steps 1 --warmup 1 --cpu-budget 1 --parity-users 64 > /dev/null 2> gpurun_out/ncu_fused.err; tail -1 gpurun_out/ncu_fused.err
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_gather_reduce -s 2 -c 1 -o gpurun_out/prof_k1_${TAG} -f python bench.py --no-extra --steps 1 --warmup 1 --users 1000000 --items 262144 --cpu-budget 1 --parity-users 64 > /dev/null 2> gpurun_out/ncu_k1.err; tail -1 gpurun_out/ncu_k1.err
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter_kernel -s 1 -c 1 -o gpurun_out/prof_fused_shard8_${TAG} -f python bench.py --no-extra --steps 1 --warmup 1 --emulate-shards 8 --cpu-budget 1 --parity-users 64 > /dev/null 2> gpurun_out/ncu_fused8.err; tail -1 gpurun_out/ncu_fused8.err
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:wmrb_step_kernel -s 1 -c 1 -o gpurun_out/prof_wmrb_${TAG} -f python bench.py --workload train --users 1000000 --steps 1 --warmup 1 --train-cpu-users 500 > /dev/null 2> gpurun_out/ncu_wmrb.err; tail -1 gpurun_out/ncu_wmrb.err
  ;;
esac
done
ls -la gpurun_out | tail -8
