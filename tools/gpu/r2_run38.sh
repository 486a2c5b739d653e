cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
export FNSSL_TC_WAIT_TIMEOUT=1
PAIR_DEBUG_ONLY=m128 timeout 600 python tools/tc5_debug.py 2>&1 | grep -v "^$" | tee $O/r2_tc6_debug_38.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "m128" 2>&1 | tail -4 | tee $O/r2_gputests_38.log
echo "== default dispatch" | tee $O/r2_tc6_time_38.log
timeout 300 python tools/lstm_time.py in16_H128x2 H128x2_add 2>&1 | grep -v "_b256\|_B" | tee -a $O/r2_tc6_time_38.log
echo "== FNSSL_TC_PAIR=0 (lstm_tc4 only)" | tee -a $O/r2_tc6_time_38.log
FNSSL_TC_PAIR=0 timeout 300 python tools/lstm_time.py in16_H128x2 H128x2_add 2>&1 | grep -v "_b256\|_B" | tee -a $O/r2_tc6_time_38.log
