cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
timeout 600 python tools/tc5_debug.py 2>&1 | tee $O/r2_tc5_debug_10.log
FNSSL_TC_PAIR_MIN=1 timeout 300 python tools/lstm_time.py full_in16_H128x2 full_in256_H128x2_add narrow_in256_H128x2 narrow_in256+16_H128x2_add 2>&1 | grep -v "_B1\|_B2" | tee $O/r2_tc5_time_10.log
