cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
timeout 300 python tools/tc5_trace.py 2>&1 | tee gpurun_out/r2_tc5_trace_22.log
