cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
timeout 150 python __graft_entry__.py smoke > $O/r2_smoke_51.log 2>&1; echo "smoke rc=$?"; tail -6 $O/r2_smoke_51.log
timeout 200 python bench.py --steps 5 --warmup 3 > $O/r2_bench_v51_cfg4.json 2> $O/r2_bench_v51.err; echo "bench rc=$?"; tail -c 1500 $O/r2_bench_v51_cfg4.json
