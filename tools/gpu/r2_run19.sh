cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
PAIR_DEBUG_ONLY=h256 timeout 600 python tools/tc5_debug.py 2>&1 | tee gpurun_out/r2_tc6_debug_19.log
timeout 300 python tools/lstm_time.py narrow_in256_H256x1_add 2>&1 | grep -v "_B1" | tee gpurun_out/r2_tc6_time_19.log
FNSSL_TC_PAIR256=0 timeout 300 python tools/lstm_time.py narrow_in256_H256x1_add 2>&1 | grep -v "_B1" | tee -a gpurun_out/r2_tc6_time_19.log
