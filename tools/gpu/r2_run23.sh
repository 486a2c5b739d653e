cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
V=$PWD/fn_ssl_b200/variants/libfnssl_b200_split.so
FNSSL_B200_LIB=$V PAIR_DEBUG_ONLY=time_c timeout 600 python tools/tc5_debug.py 2>&1 | grep -v "^$" | tee gpurun_out/r2_tc5_debug_23.log
FNSSL_B200_LIB=$V PAIR_DEBUG_ONLY=freq_c timeout 600 python tools/tc5_debug.py 2>&1 | grep -v "^$" | tee -a gpurun_out/r2_tc5_debug_23.log
timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee gpurun_out/r2_tc5_time_23.log
FNSSL_B200_LIB=$V timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee -a gpurun_out/r2_tc5_time_23.log
