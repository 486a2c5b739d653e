cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
FNSSL_TC_PAIR=1 timeout 300 python tools/lstm_time.py _b256 2>&1 | tee gpurun_out/r2_tc5_time_21.log
FNSSL_TC_PAIR=0 timeout 300 python tools/lstm_time.py _b256 2>&1 | tee -a gpurun_out/r2_tc5_time_21.log
FNSSL_TC_PAIR=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_21_pair.json
FNSSL_TC_PAIR=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_21_solo.json
