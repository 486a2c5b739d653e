cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
export FNSSL_TC_WAIT_TIMEOUT=1
PAIR_DEBUG_ONLY=_c timeout 600 python tools/tc5_debug.py 2>&1 | grep -v "^$" | grep -v h256 | tee $O/r2_tc5_debug_30.log
timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee $O/r2_tc5_time_30.log
FNSSL_TC_DEBUG=64 timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee -a $O/r2_tc5_time_30.log
timeout 300 python tools/tc5_trace.py 2>&1 | tee $O/r2_tc5_trace_30.log
