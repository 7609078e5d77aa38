set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --users 262144 --items 262144 --cpu-budget 1 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launches.err; tail -2 gpurun_out/ncu_launches.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 -o gpurun_out/prof_fused_r1 -f python bench.py --steps 1 --warmup 1 --users 262144 --items 262144 --cpu-budget 1 > /dev/null 2> gpurun_out/ncu_fused.err; tail -2 gpurun_out/ncu_fused.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_gather_reduce -s 2 -c 2 -o gpurun_out/prof_k1_r1 -f python bench.py --steps 1 --warmup 1 --users 1000000 --items 262144 --cpu-budget 1 > /dev/null 2> gpurun_out/ncu_k1.err; tail -2 gpurun_out/ncu_k1.err
ls -la gpurun_out
