# last GPU call of the round: the whole GPU suite + smoke, the default bench line, and the filter A/B against the round's v8 kernel
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_r2_v18_n1.json 2> gpurun_out/bench_r2_v18_n1.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/bench_r2_v18_n1.json | head -3
AB_ROUNDS=1 timeout 100 python scripts/filter_ab.py 1000000 125000 _ab/lib_v8.so 2>&1 | grep -v '^\[bench\]' | tee gpurun_out/probe_r2_v18_final_ab_shard8.txt
AB_ROUNDS=1 timeout 120 python scripts/filter_ab.py 1000000 1000000 _ab/lib_v8.so 2>&1 | grep -v '^\[bench\]' | tee gpurun_out/probe_r2_v18_final_ab_full.txt
