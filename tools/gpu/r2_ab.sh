# A/B of lstm_tc4 variant libraries (tools/build_variant.py): VARIANTS="a b c" TRACE="x" bash tools/gpu/r2_ab.sh TAG
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out; TAG=${1:-ab}
export FNSSL_TC_WAIT_TIMEOUT=1
for v in $VARIANTS; do
  export FNSSL_B200_LIB=$GRAFT_REPO_ROOT/fn_ssl_b200/variants/libfnssl_b200_$v.so
  echo "=== variant $v"
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lstm_layer_tcgen05 or carried_state or narrow_second or multi_tile or network_matches or end_to_end_4s or stream" > $O/r2_${TAG}_tests_$v.log 2>&1
  echo "tests rc=$? $(tail -1 $O/r2_${TAG}_tests_$v.log)"
  timeout 300 python tools/lstm_time.py $FILTER > $O/r2_${TAG}_time_$v.log 2>&1; echo "time rc=$?"; cat $O/r2_${TAG}_time_$v.log
done
for v in $TRACE; do
  export FNSSL_B200_LIB=$GRAFT_REPO_ROOT/fn_ssl_b200/variants/libfnssl_b200_$v.so
  timeout 300 python tools/tc4_trace.py "full  in256" "full  in16" "narrow in256 H256" > $O/r2_${TAG}_trace_$v.txt 2>&1
  head -34 $O/r2_${TAG}_trace_$v.txt
done
