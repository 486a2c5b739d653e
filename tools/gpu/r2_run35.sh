cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
export FNSSL_TC_WAIT_TIMEOUT=1
PAIR_DEBUG_ONLY=h256 timeout 600 python tools/tc5_debug.py 2>&1 | grep -v "^$" | tee $O/r2_tc6_debug_35.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pair_kernel_h256 or at_size or second_source" 2>&1 | tail -4 | tee $O/r2_gputests_35.log
echo "== default dispatch" | tee $O/r2_tc6_time_35.log
timeout 300 python tools/lstm_time.py H256x1_add 2>&1 | tee -a $O/r2_tc6_time_35.log
echo "== FNSSL_TC_PAIR256=0 (lstm_tc4 only)" | tee -a $O/r2_tc6_time_35.log
FNSSL_TC_PAIR256=0 timeout 300 python tools/lstm_time.py H256x1_add 2>&1 | tee -a $O/r2_tc6_time_35.log
