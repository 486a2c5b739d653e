cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
export FNSSL_TC_WAIT_TIMEOUT=1
timeout 300 python tools/tc5_trace.py 2>&1 | tee $O/r2_tc5_trace_27_leader.log
FNSSL_TC_DEBUG=16 timeout 300 python tools/tc5_trace.py 2>&1 | tee $O/r2_tc5_trace_27_mate.log
