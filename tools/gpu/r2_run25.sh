cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/r2_gputests_v25.log
export FNSSL_TC_WAIT_TIMEOUT=1
timeout 300 python tools/tc5_trace.py 2>&1 | tee $O/r2_tc5_trace_25.log
timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee $O/r2_tc5_time_25.log
FNSSL_TC_DEBUG=128 timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee -a $O/r2_tc5_time_25.log
