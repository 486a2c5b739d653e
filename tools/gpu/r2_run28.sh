cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
export FNSSL_TC_WAIT_TIMEOUT=1
V=$PWD/fn_ssl_b200/variants/libfnssl_b200_clus.so
PAIR_DEBUG_ONLY=_c timeout 600 python tools/tc5_debug.py 2>&1 | grep -v "^$" | tee $O/r2_tc5_debug_28.log
for lib in "" $V; do
  FNSSL_B200_LIB=$lib timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 H256x1_add_b15 2>&1 | tee -a $O/r2_scope_time_28.log
  FNSSL_B200_LIB=$lib timeout 300 python tools/lstm_time.py in16_H128x2 H128x2_add H256x1_add 2>&1 | grep -v "_b\|_B" | tee -a $O/r2_scope_time_28.log
done
timeout 300 python tools/tc5_trace.py 2>&1 | tee $O/r2_tc5_trace_28.log
