# round 2, call 3: A/B of epilogue variants (cell state in registers, early H_FREE poll) + MUFU rate probe
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
./tools/micro/mufu_rate.bin > $O/r2_mufu_rate.txt 2>&1; cat $O/r2_mufu_rate.txt
for v in pub1 pub1c pub1c0 prod2c; do
  export FNSSL_B200_LIB=$GRAFT_REPO_ROOT/fn_ssl_b200/variants/libfnssl_b200_$v.so
  echo "=== variant $v"
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lstm_layer_tcgen05 or carried_state or narrow_second or multi_tile or network_matches or end_to_end_4s or stream" > $O/r2_ab2_tests_$v.log 2>&1
  echo "tests rc=$? $(tail -1 $O/r2_ab2_tests_$v.log)"
  timeout 300 python tools/lstm_time.py > $O/r2_ab2_time_$v.log 2>&1; echo "time rc=$?"; cat $O/r2_ab2_time_$v.log
done
export FNSSL_B200_LIB=$GRAFT_REPO_ROOT/fn_ssl_b200/variants/libfnssl_b200_pub1c.so
timeout 300 python tools/tc4_trace.py "full  in256" "full  in16" "narrow in256 H256" > $O/r2_ab2_trace_pub1c.txt 2>&1
head -34 $O/r2_ab2_trace_pub1c.txt
