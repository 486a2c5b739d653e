cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
export FNSSL_TC_WAIT_TIMEOUT=1
V=$PWD/fn_ssl_b200/variants/libfnssl_b200_chain.so
FNSSL_B200_LIB=$V PAIR_DEBUG_ONLY=_c timeout 600 python tools/tc5_debug.py 2>&1 | grep -v "^$" | grep -v h256 | tee $O/r2_tc5_debug_26.log
timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee $O/r2_tc5_time_26.log
FNSSL_B200_LIB=$V timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee -a $O/r2_tc5_time_26.log
FNSSL_B200_LIB=$V timeout 300 python tools/tc5_trace.py 2>&1 | tee $O/r2_tc5_trace_26.log
