cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_46.log 2>&1; tail -5 $O/r2_gputests_46.log
V=$PWD/fn_ssl_b200/variants/libfnssl_b200_quadpush.so
for lib in "" $V; do
  echo "lib=[$lib] lstm_tc4 at B = 16 (pair kernels off)" | tee -a $O/r2_push_time_46.log
  FNSSL_B200_LIB=$lib FNSSL_TC_PAIR=0 FNSSL_TC_PAIR256=0 timeout 300 python tools/lstm_time.py in16_H128x2 H128x2_add in256_H256x1_add 2>&1 | grep -v "_b15\|_b60\|_b256\|_B" | tee -a $O/r2_push_time_46.log
  echo "lib=[$lib] default dispatch" | tee -a $O/r2_push_time_46.log
  FNSSL_B200_LIB=$lib timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 H256x1_add 2>&1 | grep -v "_B" | tee -a $O/r2_push_time_46.log
done
