cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
timeout 900 python tools/tc5_debug.py 2>&1 | grep -v "^$" | tee gpurun_out/r2_tc5_debug_20.log
FNSSL_TC_PAIR=1 timeout 300 python tools/lstm_time.py H128x2 2>&1 | grep -v "_B1" | tee gpurun_out/r2_tc5_time_20.log
FNSSL_TC_PAIR=0 timeout 300 python tools/lstm_time.py H128x2 2>&1 | grep -v "_B1" | tee -a gpurun_out/r2_tc5_time_20.log
