cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
LSTM_TIME_BURST=1 timeout 300 python tools/lstm_time.py in16_H128x2 H128x2_add H256x1_add 2>&1 | grep -v "_b256\|_B" | tee $O/r2_burst_33.log
